"""Drop-in replacements for the reference's pixel-level contrastive loss.

Same names and argument order as ``pixcontrast_18/contrast/models/PixPro_swin_v5.py``
(== ``pixcontrast_cata/...``):

  * ``posMask`` / ``negMask``        :48-69
  * ``regression_loss``              :71-129
  * ``consistency_loss_tail``        the part of ``ConsistencyLoss.forward`` after the encoders (:584-597)
  * ``ConsistencyLoss``              :565-597 (the module; the encoders are the caller's ``pixpro``)

``regression_loss`` never builds the five [N, HW, HW] logit tensors or the ten masks of the
reference.  One step -- one ``regression_loss`` call, or the two symmetric calls of
``ConsistencyLoss.forward`` together -- is six kernel launches without any host synchronisation
(CUDA-graph capturable): label maps are down-sampled / range-checked / counting-sorted on the
device, every embedding map is normalised + cast + stored in label order by one launch, the
similarity tiles run on tcgen05 CTA pairs and are reduced against the labels in registers
(``stswin_pixloss_fwd``); the backward regenerates the 0/1 same-label operand tile by tile
(``stswin_pixloss_bwd``) and applies the chain rule of a fused ``F.normalize`` in its last kernel.
Gradients flow to the query maps only, as in the reference (keys are built under ``no_grad``,
:366).  CUDA tensors only -- no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import StswinError

_BF16 = torch.bfloat16
_MAP_DTYPES = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}
_LABEL_DTYPES = {torch.uint8: 0, torch.float32: 1, torch.int64: 2, torch.int32: 3, torch.bfloat16: 4, torch.float16: 5}
MAX_CLASSES = 254          # labels live in one byte; 254 / 255 are reserved (mixed group / padding)


def posMask(pred1: torch.Tensor, pred2: torch.Tensor, class_num: int) -> torch.Tensor:
    """[B,1,H,W] x [B,1,H,W] -> [B,HW,HW] float: 1 where the labels agree (:48-57).
    API-compatibility helper (a label compare); the fused loss does not call it.  Like the reference's
    ``F.one_hot`` it reads the label range back (one host synchronisation)."""
    b = pred1.shape[0]
    a, c = pred1.reshape(b, -1).long(), pred2.reshape(b, -1).long()
    bad = ((a < 0) | (a >= class_num)).any() | ((c < 0) | (c >= class_num)).any()
    if bool(bad):
        raise RuntimeError("Class values must be smaller than num_classes.")
    return (a[:, :, None] == c[:, None, :]).float()


def negMask(pred1: torch.Tensor, pred2: torch.Tensor, class_num: int) -> torch.Tensor:
    """1 - posMask (:59-69)."""
    return 1 - posMask(pred1, pred2, class_num)


def downsample_labels(mask: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """F.interpolate(mask, size=[H, W], mode='nearest') for label maps (:585-590): index
    floor(dst * src/dst) in float32, which is [::8, ::8] for the reference's 256x448 -> 32x56.
    (The fused loss does this inside ``stswin_pixloss_labels``; this is the stand-alone helper.)"""
    hs, ws = mask.shape[-2:]
    ih = torch.floor(torch.arange(H, dtype=torch.float32, device=mask.device) * (torch.tensor(hs, dtype=torch.float32) / H)).long().clamp_(max=hs - 1)
    iw = torch.floor(torch.arange(W, dtype=torch.float32, device=mask.device) * (torch.tensor(ws, dtype=torch.float32) / W)).long().clamp_(max=ws - 1)
    return mask[..., ih[:, None], iw[None, :]]


def _iarr(vals: Sequence[int]):
    return (ctypes.c_int * len(vals))(*vals)


def _parr(ptrs: Sequence[int]):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class _Slots:
    """Slot tables of one loss step: unique embedding maps (query maps in pixel order, key maps in label order) and
    unique label maps."""

    def __init__(self):
        self.maps: List[torch.Tensor] = []
        self.map_label: List[int] = []          # label slot that orders a key map; -1 for query maps
        self.labels: List[torch.Tensor] = []
        self._mi, self._li = {}, {}

    def label(self, t: torch.Tensor) -> int:
        k = id(t)
        if k not in self._li:
            self._li[k] = len(self.labels)
            self.labels.append(t)
        return self._li[k]

    def map(self, t: torch.Tensor, label_slot: int) -> int:
        k = (id(t), label_slot)
        if k not in self._mi:
            self._mi[k] = len(self.maps)
            self.maps.append(t)
            self.map_label.append(label_slot)
        return self._mi[k]


def _as_map(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise StswinError("stswincl_b200.contrast needs CUDA tensors (no CPU path)")
    x = x.detach()
    if x.dtype not in _MAP_DTYPES:
        x = x.float()
    return x.contiguous()


def _as_label(l: torch.Tensor) -> torch.Tensor:
    if not l.is_cuda:
        raise StswinError("stswincl_b200.contrast needs CUDA tensors (no CPU path)")
    if l.dtype not in _LABEL_DTYPES:
        l = l.float()
    return l.contiguous()


def label_tables(labels: Sequence[torch.Tensor], H: int, W: int, class_num: int):
    """The device-side label pass on its own (``stswin_pixloss_labels``), for tests and inspection: returns
    ``(lab_nat [L,N,HWp] u8, lab_sorted [L,N,HWp] u8, glab [L,N,GLp] u8, perm [L,N,HW] i16, hist [L,N,256] i32,
    ctl [2] i32)`` for label maps [N,1,Hs,Ws] down-sampled to H x W."""
    labels = [_as_label(l) for l in labels]
    dev = labels[0].device
    N = labels[0].shape[0]
    Hs, Ws = labels[0].shape[-2:]
    HW = H * W
    HWp, GLp = (HW + 255) // 256 * 256, ((HW + 255) // 256 * 8 + 15) // 16 * 16
    L = len(labels)
    lab_nat = torch.empty((L, N, HWp), dtype=torch.uint8, device=dev)
    lab_sorted = torch.empty_like(lab_nat)
    glab = torch.empty((L, N, GLp), dtype=torch.uint8, device=dev)
    perm = torch.empty((L, N, HW), dtype=torch.int16, device=dev)
    hist = torch.empty((L, N, 256), dtype=torch.int32, device=dev)
    ctl = torch.empty(2, dtype=torch.int32, device=dev)
    with ops._launch("pix_labels", float(sum(l.numel() for l in labels)), labels[0]):
        st = _lib.load().stswin_pixloss_labels(_parr([l.data_ptr() for l in labels]), _iarr([_LABEL_DTYPES[l.dtype] for l in labels]),
                                               L, 0, N, Hs, Ws, H, W, class_num, lab_nat.data_ptr(), lab_sorted.data_ptr(),
                                               glab.data_ptr(), perm.data_ptr(), hist.data_ptr(), ctl.data_ptr(), None, 0,
                                               ops._stream(labels[0]))
    _lib.check(st, "stswin_pixloss_labels")
    return lab_nat, lab_sorted, glab, perm, hist, ctl


class _PixStepFn(torch.autograd.Function):
    """One loss step over Q queries (each with S key sets).  ``plan`` carries the slot tables; the differentiable
    inputs are the Q query maps."""

    @staticmethod
    def forward(ctx, plan, *queries):
        (slots, qmap, qlab, kmap, klab, S, class_num, normalize, gather, fp32) = plan
        lib = _lib.load()
        q0 = queries[0]
        dev = q0.device
        N, C = q0.shape[:2]
        HW = q0.numel() // (N * C)
        H, W = (q0.shape[2], q0.shape[3]) if q0.dim() == 4 else (1, HW)
        Q = len(queries)
        maps = [_as_map(m) for m in slots.maps]
        labels = [_as_label(l) for l in slots.labels]
        for m in maps:
            if m.numel() != N * C * HW or m.shape[0] != N or m.shape[1] != C:
                raise StswinError(f"embedding maps must all be [{N},{C},...] with {HW} pixels, got {tuple(m.shape)}")
        Hs, Ws = labels[0].shape[-2:]
        for l in labels:
            if tuple(l.shape[-2:]) != (Hs, Ws) or l.numel() != N * Hs * Ws:
                raise StswinError(f"label maps must all be [{N},1,{Hs},{Ws}], got {tuple(l.shape)}")
        n_local, nl_local = len(maps), len(labels)
        world, n_shared = 1, 0
        if gather is not None:
            world, n_shared = gather[1], gather[2]
        n_slots = n_local + world * n_shared if world > 1 else n_local
        if fp32:                                   # second bf16 term of every map behind the first ones
            if world > 1:
                raise StswinError("precision='fp32' and cross_rank_negatives cannot be combined")
            n_slots = 2 * n_local
        nl_slots = nl_local + world * n_shared if world > 1 else nl_local
        HWp, GLp = (HW + 255) // 256 * 256, ((HW + 255) // 256 * 8 + 15) // 16 * 16
        xn = torch.empty((n_slots, N, C, HW), dtype=_BF16, device=dev)
        inv_norm = torch.empty((n_local, N, HW), dtype=torch.float32, device=dev) if normalize else None
        ksum = torch.empty((n_slots, N, C), dtype=torch.float32, device=dev)
        lab_nat = torch.empty((nl_local, N, HWp), dtype=torch.uint8, device=dev)
        lab_sorted = torch.empty((nl_slots, N, HWp), dtype=torch.uint8, device=dev)
        glab = torch.empty((nl_slots, N, GLp), dtype=torch.uint8, device=dev)
        perm = torch.empty((nl_local, N, HW), dtype=torch.int16, device=dev)
        hist = torch.empty((nl_slots, N, 256), dtype=torch.int32, device=dev)
        ctl = torch.empty(2, dtype=torch.int32, device=dev)               # [0] bad-label flag, [1] finalize ticket
        n_terms = 3 if fp32 else 1
        stats = torch.empty((n_terms, Q, N, HW, S, 2, 2), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        loss_q = torch.empty(Q, dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad[1:])
        coef = torch.empty((Q, N, HW, S + 1), dtype=torch.float32, device=dev) if need_grad else None
        partial = torch.empty(Q * ((N * HW + 63) // 64), dtype=torch.float32, device=dev)
        dq32 = torch.empty((Q, N, HW, C), dtype=torch.float32, device=dev) if need_grad else None   # cleared by the forward
        stream = ops._stream(q0)
        with ops._launch("pix_labels", float(sum(l.numel() for l in labels)), q0):
            st = lib.stswin_pixloss_labels(_parr([l.data_ptr() for l in labels]), _iarr([_LABEL_DTYPES[l.dtype] for l in labels]),
                                           nl_local, 0, N, Hs, Ws, H, W, class_num, lab_nat.data_ptr(), lab_sorted.data_ptr(),
                                           glab.data_ptr(), perm.data_ptr(), hist.data_ptr(), ctl.data_ptr(), ksum.data_ptr(), n_local * N * C, stream)
        _lib.check(st, "stswin_pixloss_labels")
        in_bytes = float(sum(m.numel() * m.element_size() for m in maps) + 2 * n_local * N * C * HW)
        with ops._launch("pix_prepare", in_bytes, q0):
            st = lib.stswin_pixloss_prepare(_parr([m.data_ptr() for m in maps]), _iarr([_MAP_DTYPES[m.dtype] for m in maps]),
                                            _iarr(slots.map_label), n_local, 0, N, C, HW, int(normalize), n_local if fp32 else 0, perm.data_ptr(),
                                            xn.data_ptr(), ops._ptr(inv_norm), ksum.data_ptr(), ops._ptr(dq32), dq32.numel() if dq32 is not None else 0, stream)
        _lib.check(st, "stswin_pixloss_prepare")
        if world > 1:
            # C3 (an extension of the reference): the prepared shared key sets of every rank -- embeddings in
            # label order, sorted labels, group labels, histograms, channel sums -- are all-gathered into the slots
            # behind the local ones
            import torch.distributed as tdist
            group, ms, ls = gather[0], gather[3], gather[4]
            for buf, first, n_loc in ((xn, ms, n_local), (ksum, ms, n_local), (lab_sorted, ls, nl_local),
                                      (glab, ls, nl_local), (hist, ls, nl_local)):
                tdist.all_gather_into_tensor(buf[n_loc:].view(world, -1), buf[first:first + n_shared].reshape(-1), group=group)
        flops = 2.0 * Q * S * N * HW * HW * C
        if fp32:          # (q_hi, k_hi), (q_hi, k_lo), (q_lo, k_hi); backward: k_hi, k_lo
            lo = lambda idx: [i + n_local for i in idx]
            qmap_f, kmap_f = qmap + qmap + lo(qmap), kmap + lo(kmap) + kmap
            kmap_b, qmap_lo = kmap + lo(kmap), _iarr(lo(qmap))
        else:
            qmap_f, kmap_f, kmap_b, qmap_lo = qmap, kmap, kmap, None
        with ops._launch("pixloss_fwd", flops * n_terms, q0):
            st = lib.stswin_pixloss_fwd(xn.data_ptr(), n_slots, nl_slots, lab_nat.data_ptr(), lab_sorted.data_ptr(), glab.data_ptr(),
                                        hist.data_ptr(), _iarr(qmap_f), _iarr(qlab), _iarr(kmap_f), _iarr(klab), n_terms, Q, S, N, C, HW,
                                        stats.data_ptr(), loss.data_ptr(), loss_q.data_ptr(), ops._ptr(coef), ctl.data_ptr(), partial.data_ptr(),
                                        ctl.data_ptr() + 4, stream)
        _lib.check(st, "stswin_pixloss_fwd")
        ops.count_extra_launches(n_terms)                 # further terms + finalize kernel inside stswin_pixloss_fwd
        ctx.ws = (xn, lab_nat, lab_sorted, glab, coef, ksum, inv_norm, n_slots, nl_slots, dq32)
        ctx.plan = (qmap, qmap_lo, qlab, kmap_b, klab, 2 if fp32 else 1, Q, S, N, C, HW, flops)
        ctx.q_meta = [(q.shape, q.dtype) for q in queries]
        ctx.mark_non_differentiable(loss_q, ctl)
        ctx.set_materialize_grads(False)              # no zero-filled gradients for the two auxiliary outputs
        return loss, loss_q, ctl

    @staticmethod
    def backward(ctx, d_loss, _d_per_query, _d_flag):
        xn, lab_nat, lab_sorted, glab, coef, ksum, inv_norm, n_slots, nl_slots, dq32 = ctx.ws
        qmap, qmap_lo, qlab, kmap, klab, n_terms, Q, S, N, C, HW, flops = ctx.plan
        dev = xn.device
        out_dtype = ctx.q_meta[0][1] if ctx.q_meta[0][1] in _MAP_DTYPES else torch.float32
        outs = [torch.empty((N, C, HW), dtype=out_dtype, device=dev) for _ in range(Q)]
        clear = 1 if not getattr(ctx, "dq32_used", False) else 0       # a second backward through the same graph clears again
        ctx.dq32_used = True
        g = d_loss.detach().to(torch.float32).contiguous()
        with ops._launch("pixloss_bwd", flops * n_terms, xn):
            st = _lib.load().stswin_pixloss_bwd(xn.data_ptr(), n_slots, nl_slots, lab_nat.data_ptr(), lab_sorted.data_ptr(),
                                                glab.data_ptr(), _iarr(qmap), qmap_lo, _iarr(qlab), _iarr(kmap), _iarr(klab), n_terms,
                                                Q, S, N, C, HW,
                                                coef.data_ptr(), ksum.data_ptr(), g.data_ptr(), dq32.data_ptr(), clear, ops._ptr(inv_norm),
                                                _parr([o.data_ptr() for o in outs]), _MAP_DTYPES[out_dtype], ops._stream(xn))
        _lib.check(st, "stswin_pixloss_bwd")
        ops.count_extra_launches(n_terms)                 # further term + dq finish kernel inside stswin_pixloss_bwd
        grads = []
        for o, (shape, dtype) in zip(outs, ctx.q_meta):
            o = o.view(shape)
            grads.append(o if o.dtype == dtype else o.to(dtype))
        return (None, *grads)


def _loss_step(queries: Sequence[Tuple[torch.Tensor, torch.Tensor]],
               key_sets: Sequence[Sequence[Tuple[torch.Tensor, torch.Tensor]]], class_num: int, *, normalize: bool,
               validate_labels, gather=None, precision: str = "bf16"):
    """queries: [(q map, its label map)]; key_sets[q]: [(key map, its label map)] -- the same number for every query.
    Returns (total loss, per-query losses, int32 control words: [0] != 0 when a label was out of range)."""
    if not 1 <= class_num <= MAX_CLASSES:
        raise StswinError(f"class_num={class_num} unsupported: labels are stored in one byte (1..{MAX_CLASSES} classes)")
    S = len(key_sets[0])
    assert all(len(ks) == S for ks in key_sets) and 1 <= S <= 64 and 1 <= len(queries) <= 2
    slots = _Slots()
    qmap, qlab, kmap, klab = [], [], [], []
    for (q, lq) in queries:
        qlab.append(slots.label(lq))
        qmap.append(slots.map(q, -1))
    shared_first = None
    for ks in key_sets:
        for (k, lk) in ks:
            ls = slots.label(lk)
            klab.append(ls)
            kmap.append(slots.map(k, ls))
    if gather is not None:
        # gather = (group, world, shared maps [(k, lk)]): their slots must be contiguous (they are registered in
        # order by the loops above) -- other ranks' copies land behind the local slots
        group, world, shared = gather
        ms = [slots.map(k, slots.label(lk)) for (k, lk) in shared]
        ls = [slots.label(lk) for (_, lk) in shared]
        assert ms == list(range(ms[0], ms[0] + len(ms))) and ls == list(range(ls[0], ls[0] + len(ls)))
        n_local, nl_local, n_sh = len(slots.maps), len(slots.labels), len(shared)
        import torch.distributed as tdist
        rank = tdist.get_rank(group)
        Q = len(queries)
        kmap2, klab2 = [], []
        for qi in range(Q):
            kmap2 += kmap[qi * S:(qi + 1) * S]
            klab2 += klab[qi * S:(qi + 1) * S]
            for r in range(world):
                if r != rank:
                    kmap2 += [n_local + r * n_sh + j for j in range(n_sh)]
                    klab2 += [nl_local + r * n_sh + j for j in range(n_sh)]
        S = S + (world - 1) * n_sh
        kmap, klab = kmap2, klab2
        gather = (group, world, n_sh, ms[0], ls[0])
    if precision not in ("bf16", "fp32", "auto"):
        raise ValueError("precision must be 'bf16', 'fp32' or 'auto'")
    fp32 = precision == "fp32" or (precision == "auto" and queries[0][0].dtype == torch.float32 and not torch.is_autocast_enabled())
    plan = (slots, qmap, qlab, kmap, klab, S, class_num, bool(normalize), gather, fp32)
    total, per_query, flag = _PixStepFn.apply(plan, *[q for q, _ in queries])
    if validate_labels is True or validate_labels == "host":
        if int(flag[0]) != 0:                             # F.one_hot raises for labels outside [0, class_num) (:54-55)
            raise RuntimeError("Class values must be smaller than num_classes.")
    return total, per_query, flag


def pixel_contrast_loss(q: torch.Tensor, keys: Sequence[torch.Tensor], label_q: torch.Tensor,
                        labels_k: Sequence[torch.Tensor], class_num: int, *, normalize: bool = False,
                        validate_labels="device", precision: str = "bf16") -> torch.Tensor:
    """General form: any 1..64 key sets (e.g. key sets all-gathered from other ranks).  ``normalize=True``
    fuses ``F.normalize(dim=1)`` of q and of every key into the loss (otherwise the inputs are taken as already
    unit-norm, like the reference's ``regression_loss``).

    ``validate_labels``: the label range check of ``F.one_hot`` (:54-55) runs on the device.  ``"device"`` (default):
    no host synchronisation -- an out-of-range label makes the loss NaN; ``True`` / ``"host"``: additionally read the
    flag back and raise ``RuntimeError`` like the reference (one synchronisation, not CUDA-graph capturable)."""
    if not q.is_cuda:
        raise StswinError("stswincl_b200.contrast needs CUDA tensors (no CPU path)")
    assert len(keys) == len(labels_k) and 1 <= len(keys) <= 64
    total, _, _ = _loss_step([(q, label_q)], [list(zip(keys, labels_k))], class_num, normalize=normalize,
                             validate_labels=validate_labels, precision=precision)
    return total


def regression_loss(q, k, adj1, adj2, adj3, neg3, label_patch1, label_patch2, label_adj1, label_adj2, label_adj3,
                    label_neg3, class_num, validate_labels="device", precision: str = "bf16"):
    """Same signature and value as PixPro_swin_v5.py:71-129 (inputs already L2-normalised).  ``precision="fp32"``: the
    similarities as three bf16 product terms (hi*hi + hi*lo + lo*hi), loss and gradient <= 1e-3 against the fp32 run."""
    return pixel_contrast_loss(q, [k, adj1, adj2, adj3, neg3], label_patch1,
                               [label_patch2, label_adj1, label_adj2, label_adj3, label_neg3], class_num,
                               validate_labels=validate_labels, precision=precision)


def consistency_loss_tail(pred_1, pred_2, proj_1_ng, proj_2_ng, proj_adj1_ng, proj_adj2_ng, proj_adj3_ng, proj_neg3_ng,
                          mask_1, mask_2, mask_3, mask_4, mask_5, mask_6, class_num, *, normalize: bool = False,
                          cross_rank_negatives: bool = False, group=None, validate_labels="device", precision: str = "bf16"):
    """``ConsistencyLoss.forward`` after ``self.pixpro(...)`` (:584-597): nearest label down-sampling to the embedding
    resolution, then the symmetric sum of two ``regression_loss`` calls -- here ONE fused step: both calls share the
    label pass, the normalise / cast pass, one similarity launch and one backward launch; the four shared key sets are
    prepared once.

    ``cross_rank_negatives=True`` (an EXTENSION of the reference, SURVEY D5 / C3 -- parity unpinned, checked against the
    oracle's own generalisation): the four shared key sets are all-gathered over ``group`` in their prepared form
    (bf16, label order) and the other ranks' copies are appended as extra key sets, so the inter-video negatives span
    every rank.  With one rank it is the reference loss."""
    shared = [(proj_adj1_ng, mask_3), (proj_adj2_ng, mask_4), (proj_adj3_ng, mask_5), (proj_neg3_ng, mask_6)]
    gather = None
    if cross_rank_negatives:
        import torch.distributed as tdist
        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(group) > 1:
            gather = (group, tdist.get_world_size(group), shared)
    total, _, _ = _loss_step([(pred_1, mask_1), (pred_2, mask_2)],
                             [[(proj_2_ng, mask_2), *shared], [(proj_1_ng, mask_1), *shared]], class_num,
                             normalize=normalize, validate_labels=validate_labels, gather=gather, precision=precision)
    return total


class ConsistencyLoss(nn.Module):
    """Drop-in for ``ConsistencyLoss`` (PixPro_swin_v5.py:565-597): ``forward(im_1..im_6, mask_1..mask_6) -> loss``.

    The reference builds its ``PixPro`` (ResNet-18 + Swin head + ASPP + projection heads, twice) from hard-coded
    checkpoint paths inside the constructor; those encoders are the caller's side of the boundary, so the
    module takes the built ``pixpro`` -- any module with the reference's
    ``forward(seq_1..seq_6) -> (pred_1, pred_2, proj_1_ng, proj_2_ng, proj_adj1_ng, proj_adj2_ng, proj_adj3_ng,
    proj_neg3_ng)`` contract, e.g. the reference class itself with ``encoder_2`` / ``encoder_k_2`` replaced by
    ``stswincl_b200.swin.SwinTransformerLayerv5``, or ``stswincl_b200.pixpro.PixPro``.
    ``args`` needs ``data`` ('endo18' -> 12 classes, 'cata' -> ``num_class_table[args.tag]``) and
    ``pixpro_pos_ratio`` exactly as at :568-575."""

    num_class_table = {'1': 9, '2': 18, '3': 26}      # PixPro_swin_v5.py:14

    def __init__(self, args, pixpro: Optional[nn.Module] = None, *, cross_rank_negatives: bool = False,
                 validate_labels="device"):
        super().__init__()
        self.pixpro_pos_ratio = getattr(args, "pixpro_pos_ratio", None)
        if pixpro is None:
            raise ValueError("ConsistencyLoss needs the built `pixpro` module (the reference class, or stswincl_b200.pixpro.PixPro "
                             "around the caller's encoders): ResNet / ASPP / projection heads are not part of this library")
        self.pixpro = pixpro
        self.fuse_normalize = bool(getattr(pixpro, "fuse_normalize", False))
        if args.data == 'endo18':
            self.class_num = 12
        elif args.data == 'cata':
            self.class_num = int(self.num_class_table[args.tag])
        else:
            raise ValueError(f"unknown args.data {args.data!r} (the reference knows 'endo18' and 'cata')")
        self.cross_rank_negatives = cross_rank_negatives
        self.validate_labels = validate_labels

    def forward(self, im_1, im_2, im_3, im_4, im_5, im_6, mask_1, mask_2, mask_3, mask_4, mask_5, mask_6):
        (pred_1, pred_2, proj_1_ng, proj_2_ng, proj_adj1_ng, proj_adj2_ng, proj_adj3_ng,
         proj_neg3_ng) = self.pixpro(im_1, im_2, im_3, im_4, im_5, im_6)
        return consistency_loss_tail(pred_1, pred_2, proj_1_ng, proj_2_ng, proj_adj1_ng, proj_adj2_ng, proj_adj3_ng,
                                     proj_neg3_ng, mask_1, mask_2, mask_3, mask_4, mask_5, mask_6, self.class_num,
                                     normalize=self.fuse_normalize, cross_rank_negatives=self.cross_rank_negatives,
                                     validate_labels=self.validate_labels)
