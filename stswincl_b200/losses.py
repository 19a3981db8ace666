"""Drop-in replacement for the reference's OHEM segmentation loss.

Same class name, constructor and forward contract as ``seg18/utils/losses.py:16-40``
(``OhemCELoss2D(n_min, thresh=0.7, ignore_index=-1)``; ``forward(pred[B,K,H,W], target[B,H,W]) -> 0-dim``).
The reference sorts all ``B*H*W`` per-pixel losses every step only to read ``loss[n_min]``; here one
pass computes the per-pixel cross-entropy together with the count / sum of the losses above the
threshold, and the rarely needed "mean of the n_min largest" branch finds the n_min-th value with a
three-pass radix select -- all on the device, no host synchronisation (``stswin_ohem_ce_fwd``).
The backward recomputes the softmax of the selected pixels only (``stswin_ohem_ce_bwd``).
CUDA tensors only -- no CPU path.
"""
from __future__ import annotations

import math

import torch

from . import _lib, ops
from ._lib import StswinError


class _OhemCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, n_min, thresh, ignore_index):
        B, K = logits.shape[0], logits.shape[1]
        HW = logits[0, 0].numel()
        lib = _lib.load()
        dev = logits.device
        loss_px = torch.empty(B * HW, dtype=torch.float32, device=dev)
        ws = torch.empty(int(lib.stswin_ohem_ws_bytes()), dtype=torch.uint8, device=dev)
        out = torch.empty(4 + 1, dtype=torch.float32, device=dev)       # sel[4] | loss
        sel, loss = out[:4], out[4]
        is_f32 = int(logits.dtype == torch.float32)
        esz = logits.element_size()
        with ops._launch("ohem_ce_fwd", float(B * HW * (K * esz + 8 + 4)), logits):     # logits + labels in, losses out
            st = lib.stswin_ohem_ce_fwd(logits.data_ptr(), is_f32, labels.data_ptr(), B, K, HW, ignore_index, thresh, n_min,
                                        loss_px.data_ptr(), ws.data_ptr(), loss.data_ptr(), sel.data_ptr(),
                                        ops._stream(logits))
        _lib.check(st, "stswin_ohem_ce_fwd")
        ctx.save_for_backward(logits, labels, loss_px, sel)
        ctx.ignore_index = ignore_index
        return loss.clone()

    @staticmethod
    def backward(ctx, d_loss):
        logits, labels, loss_px, sel = ctx.saved_tensors
        B, K = logits.shape[0], logits.shape[1]
        HW = logits[0, 0].numel()
        d_loss = d_loss.to(torch.float32).contiguous()
        d_logits = torch.empty_like(logits)
        esz = logits.element_size()
        with ops._launch("ohem_ce_bwd", float(B * HW * (K * esz + 8 + 4)), logits):     # at least: gradient out, labels + losses in
            st = _lib.load().stswin_ohem_ce_bwd(logits.data_ptr(), int(logits.dtype == torch.float32), labels.data_ptr(), B, K,
                                                HW, ctx.ignore_index, loss_px.data_ptr(), sel.data_ptr(), d_loss.data_ptr(),
                                                d_logits.data_ptr(), ops._stream(logits))
        _lib.check(st, "stswin_ohem_ce_bwd")
        return d_logits, None, None, None, None


class OhemCELoss2D(torch.nn.CrossEntropyLoss):
    """2D cross-entropy with online hard example mining (``seg18/utils/losses.py:16-40``)."""

    def __init__(self, n_min, thresh=0.7, ignore_index=-1):
        super().__init__(None, None, ignore_index, reduction='none')
        self.thresh = -math.log(thresh)
        self.n_min = n_min
        self.ignore_index = ignore_index

    def forward(self, pred, target):
        return self.OhemCELoss(pred, target)

    def OhemCELoss(self, logits, labels):
        if not logits.is_cuda:
            raise StswinError("OhemCELoss2D: logits must be a CUDA tensor (stswincl_b200 has no CPU path)")
        if logits.dim() < 2 or labels.shape != logits.shape[:1] + logits.shape[2:]:
            raise ValueError(f"Expected target of shape {tuple(logits.shape[:1] + logits.shape[2:])}, got {tuple(labels.shape)}")
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        n_pix = labels.numel()
        if not 0 <= self.n_min < n_pix:
            raise IndexError(f"index {self.n_min} is out of bounds for dimension 0 with size {n_pix}")
        if self.n_min == 0:
            raise StswinError("OhemCELoss2D: n_min must be at least 1 (the reference returns the mean of an empty slice)")
        return _OhemCEFn.apply(logits.contiguous(), labels.to(torch.int64).contiguous(), int(self.n_min), float(self.thresh),
                               int(self.ignore_index))
