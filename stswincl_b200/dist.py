"""Data-parallel plumbing for the two hot paths (one process per GPU, torch.distributed).

Both paths shard over clips with no data-path collective: windows never
cross a clip and the reference loss never crosses a sample.  What remains is what the reference
does with ``DistributedSampler`` + DDP (``pixcontrast_18/main_pretrain_swinv5.py:54``,
``contrast/data/__init__.py:25``) and, for the segmentation scripts, with ``nn.DataParallel``
(``seg18/train_swin.py:131-135``): split the clips, all-reduce the gradients.

These helpers are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Indices of the samples rank ``rank`` owns: DistributedSampler(shuffle=False) semantics --
    the index list is padded by wrapping around to a multiple of ``world`` and dealt out
    round-robin (contrast/data/__init__.py:25)."""
    total = (n + world - 1) // world * world
    idx = list(range(n))
    pad = total - n
    if pad:
        idx += (idx * ((pad + n - 1) // n))[:pad]
    return idx[rank:total:world]


def scaled_lr(base_lr: float, batch_size: int, world: int) -> float:
    """lr = batch_size * world_size / 256 * base_lr  (main_pretrain_swinv5.py:38,45)."""
    return batch_size * world / 256.0 * base_lr


def average_gradients(params: Iterable[torch.nn.Parameter], *, bucket_bytes: int = 64 << 20,
                      comm_dtype: Optional[torch.dtype] = None, group=None) -> int:
    """All-reduce (mean) the ``.grad`` of every parameter in flat buckets of ``bucket_bytes``.
    ``comm_dtype=torch.bfloat16`` halves the bytes on the wire (the 96.6 M-parameter Swin head is
    386 MB in fp32).  Returns the number of collectives issued.  DDP does the same overlapped with
    backward; this explicit form is what the CPU tests exercise and what a DataParallel-free
    training loop can call."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    calls, i = 0, 0
    while i < len(grads):
        bucket, size = [], 0
        while i < len(grads) and (not bucket or size + grads[i].numel() * grads[i].element_size() <= bucket_bytes):
            bucket.append(grads[i]); size += grads[i].numel() * grads[i].element_size(); i += 1
        flat = torch.cat([g.reshape(-1) for g in bucket])
        wire = flat.to(comm_dtype) if comm_dtype is not None and comm_dtype != flat.dtype else flat
        dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=group)
        flat = wire.to(flat.dtype) if wire is not flat else flat
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g)); off += g.numel()
        calls += 1
    return calls


class FlatGradients:
    """One flat fp32 buffer holding the gradients of ``params`` (each view 16-byte aligned), for a single
    all-reduce per step instead of DDP's buckets -- the data-parallel exchange of ``bench.py --dp flat``:

        loss.backward()            # may be replayed from a CUDA graph together with ``gather()``
        flat.gather()              # copy every p.grad into the buffer (one multi-tensor copy)
        flat.all_reduce()          # ONE collective (mean over ranks)
        flat.bind()                # p.grad = view of the reduced buffer
        optimizer.step()

    Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        dev = self.params[0].device
        self.buffer = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p, n in zip(self.params, sizes):
            self.views.append(self.buffer[off:off + p.numel()].view_as(p))
            off += n

    def gather(self) -> None:
        grads = [p.grad for p in self.params]
        if any(g is None for g in grads):
            raise RuntimeError("FlatGradients.gather: a parameter has no gradient")
        torch._foreach_copy_(self.views, grads)

    def all_reduce(self, group=None) -> None:
        world = dist.get_world_size(group)
        if world == 1:
            return
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.buffer, op=dist.ReduceOp.AVG, group=group)
        else:                                   # gloo has no AVG
            dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)
            self.buffer.div_(world)

    def bind(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v


def broadcast_parameters(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Replicas start from rank ``src``'s parameters (what DDP's constructor does)."""
    for p in params:
        dist.broadcast(p.data, src=src, group=group)


def gather_key_sets(tensors: Sequence[torch.Tensor], group=None) -> List[torch.Tensor]:
    """All-gather every tensor of ``tensors`` from all ranks; returns, per input tensor, the list of
    per-rank copies flattened into one list (own rank's copy first).  Building block of the
    'negatives span all ranks' extension: the gathered key sets are appended as
    extra sets of ``pixel_contrast_loss`` -- *not* reference behaviour (parity unpinned)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out: List[torch.Tensor] = []
    for t in tensors:
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        out.extend([parts[rank]] + [p for r, p in enumerate(parts) if r != rank])
    return out
