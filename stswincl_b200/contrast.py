"""Drop-in replacements for the reference's pixel-level contrastive loss.

Same names and argument order as ``pixcontrast_18/contrast/models/PixPro_swin_v5.py``
(== ``pixcontrast_cata/...``):

  * ``posMask`` / ``negMask``        :48-69
  * ``regression_loss``              :71-129
  * ``consistency_loss_tail``        the part of ``ConsistencyLoss.forward`` after the encoders (:584-597)

``regression_loss`` never builds the five [N, HW, HW] logit tensors or the ten masks of the
reference: one fused kernel computes the similarity tiles on the tensor cores and reduces them
against the labels in registers (``stswin_pixloss_fwd``); the backward regenerates the 0/1
same-label operand tile by tile (``stswin_pixloss_bwd``).  Gradients flow to ``q`` only, as in
the reference (keys are built under ``no_grad``, :366).  CUDA tensors only -- no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib, ops
from ._lib import StswinError

_BF16 = torch.bfloat16


def _flat_labels(lbl: torch.Tensor) -> torch.Tensor:
    """[N,1,H,W] float / int label map -> [N, HW] uint8 (``.long()`` truncation, :54)."""
    n = lbl.shape[0]
    return lbl.reshape(n, -1).to(torch.uint8).contiguous()


def posMask(pred1: torch.Tensor, pred2: torch.Tensor, class_num: int) -> torch.Tensor:
    """[B,1,H,W] x [B,1,H,W] -> [B,HW,HW] float: 1 where the labels agree (:48-57).
    API-compatibility helper (a label compare); the fused loss does not call it."""
    b = pred1.shape[0]
    a, c = pred1.reshape(b, -1).long(), pred2.reshape(b, -1).long()
    if int(torch.maximum(a.max(), c.max())) >= class_num or int(torch.minimum(a.min(), c.min())) < 0:
        raise RuntimeError("Class values must be smaller than num_classes.")
    return (a[:, :, None] == c[:, None, :]).float()


def negMask(pred1: torch.Tensor, pred2: torch.Tensor, class_num: int) -> torch.Tensor:
    """1 - posMask (:59-69)."""
    return 1 - posMask(pred1, pred2, class_num)


def downsample_labels(mask: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """F.interpolate(mask, size=[H, W], mode='nearest') for label maps (:585-590): index
    floor(dst * src/dst) in float32, which is [::8, ::8] for the reference's 256x448 -> 32x56."""
    hs, ws = mask.shape[-2:]
    ih = torch.floor(torch.arange(H, dtype=torch.float32, device=mask.device) * (torch.tensor(hs, dtype=torch.float32) / H)).long().clamp_(max=hs - 1)
    iw = torch.floor(torch.arange(W, dtype=torch.float32, device=mask.device) * (torch.tensor(ws, dtype=torch.float32) / W)).long().clamp_(max=ws - 1)
    return mask[..., ih[:, None], iw[None, :]]


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


class _Prepared:
    """bf16 (optionally L2-normalised) copy of an embedding map + its per-channel sums."""
    __slots__ = ("xn", "inv_norm", "ksum")

    def __init__(self, x: torch.Tensor, normalize: bool, want_ksum: bool):
        if not x.is_cuda:
            raise StswinError("stswincl_b200.contrast needs CUDA tensors (no CPU path)")
        if x.dtype not in (torch.float32, _BF16):
            x = x.float()
        N, C = x.shape[:2]
        HW = x.numel() // (N * C)
        x = x.contiguous()
        self.xn = torch.empty((N, C, HW), dtype=_BF16, device=x.device)
        self.inv_norm = torch.empty((N, HW), dtype=torch.float32, device=x.device) if normalize else None
        self.ksum = torch.empty((N, C), dtype=torch.float32, device=x.device) if want_ksum else None
        with ops._launch("pix_normalize", float(x.numel() * (x.element_size() + 2)), x):
            st = _lib.load().stswin_pix_normalize(x.data_ptr(), int(x.dtype == torch.float32), self.xn.data_ptr(),
                                                  ops._ptr(self.inv_norm), ops._ptr(self.ksum), N, C, HW, int(normalize),
                                                  ops._stream(x))
        _lib.check(st, "stswin_pix_normalize")


class _PixLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, normalize, lq, lks, *prepared_keys):
        N, C = q.shape[:2]
        HW = q.numel() // (N * C)
        n_sets = len(prepared_keys)
        pq = _Prepared(q.detach(), normalize, want_ksum=False)
        dev = q.device
        stats = torch.empty((N, HW, n_sets, 4), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty((N, HW, n_sets + 1), dtype=torch.float32, device=dev)
        keys = [p.xn for p in prepared_keys]
        with ops._launch("pixloss_fwd", 2.0 * n_sets * N * HW * HW * C, q):
            st = _lib.load().stswin_pixloss_fwd(pq.xn.data_ptr(), _ptr_array(keys), lq.data_ptr(), _ptr_array(lks), n_sets,
                                                N, C, HW, stats.data_ptr(), loss.data_ptr(), coef.data_ptr(), ops._stream(q))
        _lib.check(st, "stswin_pixloss_fwd")
        ctx.keys, ctx.lks, ctx.lq = keys, list(lks), lq
        ctx.ksum = torch.stack([p.ksum for p in prepared_keys], 0).contiguous()
        ctx.coef, ctx.pq, ctx.normalize = coef, pq, normalize
        ctx.q_shape, ctx.q_dtype = q.shape, q.dtype
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        N, C = ctx.q_shape[:2]
        HW = ctx.pq.xn.shape[2]
        n_sets = len(ctx.keys)
        dev = ctx.pq.xn.device
        dq32 = torch.empty((N, HW, C), dtype=torch.float32, device=dev)
        g = d_loss.detach().to(torch.float32).contiguous()
        with ops._launch("pixloss_bwd", 2.0 * n_sets * N * HW * HW * C, dq32):
            st = _lib.load().stswin_pixloss_bwd(_ptr_array(ctx.keys), ctx.lq.data_ptr(), _ptr_array(ctx.lks), ctx.coef.data_ptr(),
                                                ctx.ksum.data_ptr(), g.data_ptr(), n_sets, N, C, HW, dq32.data_ptr(),
                                                ops._stream(dq32))
        _lib.check(st, "stswin_pixloss_bwd")
        dq = ops.transpose(dq32, torch.float32)                       # [N, C, HW]
        if ctx.normalize:                                              # chain rule through x / max(|x|, eps)
            qn = ctx.pq.xn.float()
            dq = ctx.pq.inv_norm[:, None, :] * (dq - qn * (qn * dq).sum(1, keepdim=True))
        return (dq.view(ctx.q_shape).to(ctx.q_dtype), None, None, None) + (None,) * n_sets


def pixel_contrast_loss(q: torch.Tensor, keys: Sequence[torch.Tensor], label_q: torch.Tensor,
                        labels_k: Sequence[torch.Tensor], class_num: int, *, normalize: bool = False,
                        validate_labels: bool = True, _cache: Optional[dict] = None) -> torch.Tensor:
    """General form: any 1..64 key sets (e.g. key sets all-gathered from other ranks).  ``normalize=True`` fuses ``F.normalize(dim=1)`` of q and of
    every key into the loss (otherwise the inputs are taken as already unit-norm, like the
    reference's ``regression_loss``)."""
    if not q.is_cuda:
        raise StswinError("stswincl_b200.contrast needs CUDA tensors (no CPU path)")
    assert len(keys) == len(labels_k) and 1 <= len(keys) <= 64
    if validate_labels:       # F.one_hot raises for labels outside [0, class_num) (:54-55)
        hi = torch.stack([l.max() for l in (label_q, *labels_k)]).max()
        lo = torch.stack([l.min() for l in (label_q, *labels_k)]).min()
        if int(hi) >= class_num or int(lo) < 0:
            raise RuntimeError("Class values must be smaller than num_classes.")
    cache = _cache if _cache is not None else {}
    prepared = []
    for k in keys:
        key = (k.data_ptr(), k._version, normalize)
        if key not in cache:
            cache[key] = _Prepared(k.detach(), normalize, want_ksum=True)
        prepared.append(cache[key])
    lq = _flat_labels(label_q)
    lks = [_flat_labels(l) for l in labels_k]
    return _PixLossFn.apply(q, normalize, lq, lks, *prepared)


def regression_loss(q, k, adj1, adj2, adj3, neg3, label_patch1, label_patch2, label_adj1, label_adj2, label_adj3,
                    label_neg3, class_num):
    """Same signature and value as PixPro_swin_v5.py:71-129 (inputs already L2-normalised)."""
    return pixel_contrast_loss(q, [k, adj1, adj2, adj3, neg3], label_patch1,
                               [label_patch2, label_adj1, label_adj2, label_adj3, label_neg3], class_num)


def consistency_loss_tail(pred_1, pred_2, proj_1_ng, proj_2_ng, proj_adj1_ng, proj_adj2_ng, proj_adj3_ng, proj_neg3_ng,
                          mask_1, mask_2, mask_3, mask_4, mask_5, mask_6, class_num, *, normalize: bool = False,
                          cross_rank_negatives: bool = False, group=None):
    """``ConsistencyLoss.forward`` after ``self.pixpro(...)`` (:584-597): nearest label down-sampling
    to the embedding resolution, then the symmetric sum of two ``regression_loss`` calls; the four
    shared key sets are prepared once.

    ``cross_rank_negatives=True`` (an EXTENSION of the reference, SURVEY D5 / C3 -- parity unpinned,
    checked against the oracle's own generalisation): the four shared key sets and their labels are
    all-gathered over ``group`` and the other ranks' copies are appended as extra key sets, so the
    inter-video negatives span every rank.  With one rank it is the reference loss."""
    H, W = pred_1.shape[-2:]
    m = [downsample_labels(x, H, W) for x in (mask_1, mask_2, mask_3, mask_4, mask_5, mask_6)]
    cache: dict = {}
    shared, shared_l = [proj_adj1_ng, proj_adj2_ng, proj_adj3_ng, proj_neg3_ng], m[2:]
    if cross_rank_negatives:
        import torch.distributed as tdist
        from . import dist as sdist
        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(group) > 1:
            world = tdist.get_world_size(group)
            gk = sdist.gather_key_sets([k.detach().to(_BF16).contiguous() for k in shared], group)
            gl = sdist.gather_key_sets([l.to(torch.uint8).contiguous() for l in shared_l], group)
            # gather_key_sets returns, per input, [own, others...]: keep own sets first, then every other rank's
            own_k, own_l = gk[0::world], gl[0::world]
            oth_k = [t for i in range(len(shared)) for t in gk[i * world + 1:(i + 1) * world]]
            oth_l = [t for i in range(len(shared)) for t in gl[i * world + 1:(i + 1) * world]]
            shared, shared_l = list(own_k) + oth_k, list(own_l) + oth_l
    return (pixel_contrast_loss(pred_1, [proj_2_ng, *shared], m[0], [m[1], *shared_l], class_num, normalize=normalize, _cache=cache)
            + pixel_contrast_loss(pred_2, [proj_1_ng, *shared], m[1], [m[0], *shared_l], class_num, normalize=normalize, _cache=cache))
