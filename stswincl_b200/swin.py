"""Drop-in replacements for the reference's spatio-temporal Swin head.

Same constructor signatures, same ``state_dict`` keys and shapes, same forward
contracts as ``seg18/net/Ours/swin_512.py`` (== ``segcata/net/Ours/swin_tem_cata.py``
== ``pixcontrast_*/contrast/models/Ours/swin_tem.py``):

  * ``WindowAttention``          swin_512.py:73-141
  * ``Mlp``                      swin_512.py:7-23
  * ``SwinTransformerBlock``     swin_512.py:143-237  (post-norm, mask fill -100)
  * ``PatchMerging``             swin_512.py:239-277
  * ``SwinTransformerLayerv5``   swin_512.py:280-327

so ``model.swin = stswincl_b200.swin.SwinTransformerLayerv5(...)`` followed by
``load_state_dict(reference_state, strict=True)`` is the whole integration.

All arithmetic runs in the sm_100a kernels behind ``include/stswin_b200.h`` through
``torch.autograd.Function`` wrappers: bf16 storage, fp32 accumulation / statistics.
Parameters stay fp32 (the reference's master copy under AMP) and are rounded to bf16
per call.  There is no CPU path and no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import StswinError
from .optim import bf16_weight as _w16

_BF16 = torch.bfloat16
# bf16 mode: keep GELU'(fc1 output) for the backward as one byte per element (step 0.005 over [-0.14, 1.14]; the bf16 rounding
# of a value near 1 is 0.002-0.004) instead of bf16 -- the fc1 + GELU layer is HBM-write bound.  False
# restores the bf16 form.
GELU_GRAD_Q8 = True


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _relative_position_index(ws_h: int, ws_w: int) -> torch.Tensor:
    """Closed form of swin_512.py:88-99."""
    n = torch.arange(ws_h * ws_w)
    h, w = n // ws_w, n % ws_w
    dh = h[:, None] - h[None, :] + ws_h - 1
    dw = w[:, None] - w[None, :] + ws_w - 1
    return dh * (2 * ws_w - 1) + dw


def _shift_attn_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """Closed form of swin_512.py:171-192 -- kept as a buffer for state_dict parity only; the
    kernels rebuild the mask from (H, W, ws, shift)."""
    def band(p, extent):
        return (p >= extent - ws).long() + (p >= extent - shift).long()
    ids = 3 * band(torch.arange(H), H)[:, None] + band(torch.arange(W), W)[None, :]
    per_win = ids.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    differ = per_win[:, None, :] != per_win[:, :, None]
    return torch.where(differ, torch.tensor(-100.0), torch.tensor(0.0))


def _wgrad_splits(n_out: int, k_in: int, tokens: int) -> int:
    tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    return max(1, min((tokens + 63) // 64, round(2 * 148 / tiles)))


def _linear_wgrad(dy2d: torch.Tensor, x2d: torch.Tensor, shape, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW[N_out, K_in] = dy^T x, reduced over tokens on the tensor cores (fp32, split-K).
    ``out`` must be zero-filled (the split-K epilogue accumulates)."""
    dw = out if out is not None else torch.zeros(shape, dtype=torch.float32, device=dy2d.device)
    ops.gemm(dy2d, x2d, a_mn_major=True, b_mn_major=True, mode=ops.EPI_F32_REDUCE, out=dw,
             k_splits=_wgrad_splits(shape[0], shape[1], dy2d.shape[0]))
    return dw


def _zeros_like_many(shapes, device):
    """One zero-filled fp32 allocation carved into the given shapes (one memset instead of many)."""
    sizes = [int(math.prod(s)) for s in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 3) // 4 * 4            # keep every view 16-byte aligned (TMA reduce target)
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    return [flat[o:o + n].view(s) for o, n, s in zip(offs, sizes, shapes)]


class _AttentionFn(torch.autograd.Function):
    """qkv Linear -> window attention core -> proj Linear on tokens in natural order.

    x [Bp, T, H*W, C] bf16.  Gradients for x, table, qkv.{weight,bias}, proj.{weight,bias}."""

    @staticmethod
    def forward(ctx, x, table, w_qkv, b_qkv, w_proj, b_proj, geom, mask=None):
        H, W, nH, ws, shift, qk_scale = geom
        Bp, T, L, C = x.shape
        x2 = x.reshape(-1, C)
        wq, wp = _w16(w_qkv), _w16(w_proj)
        qkv = ops.gemm(x2, wq, bias=b_qkv)
        attn, lse2 = ops.winattn_fwd(qkv.view(Bp, T, L, 3 * C), table, H, W, nH, ws, shift, qk_scale=qk_scale, mask=mask)
        out = ops.gemm(attn.view(-1, C), wp, bias=b_proj)
        ctx.save_for_backward(x2, qkv, attn, lse2, table, wq, wp)
        ctx.geom = geom
        ctx.mask = mask
        ctx.has_qkv_bias = b_qkv is not None
        return out.view(Bp, T, L, C)

    @staticmethod
    def backward(ctx, d_out):
        x2, qkv, attn, lse2, table, wq, wp = ctx.saved_tensors
        H, W, nH, ws, shift, qk_scale = ctx.geom
        C = x2.shape[1]
        Bp_T_L = d_out.shape[:3]
        d2 = d_out.contiguous().view(-1, C)
        d_table, d_bqkv, d_bproj, d_wproj_, d_wqkv_ = _zeros_like_many([tuple(table.shape), (3 * C,), (C,), tuple(wp.shape), tuple(wq.shape)], d2.device)
        ops.colsum(d2, d_bproj)                           # proj bias gradient: one read of d_out
        if not ctx.has_qkv_bias:
            d_bqkv = None
        d_attn = ops.gemm(d2, wp, b_mn_major=True)
        d_wproj = _linear_wgrad(d2, attn.view(-1, C), wp.shape, d_wproj_)
        d_qkv = ops.winattn_bwd(qkv.view(*Bp_T_L, 3 * C), table, lse2, d_attn.view(*Bp_T_L, C), H, W, nH, ws, shift,
                                d_table, d_bqkv, qk_scale=qk_scale, mask=ctx.mask)
        dq2 = d_qkv.view(-1, 3 * C)
        d_x = ops.gemm(dq2, wq, b_mn_major=True)
        d_wqkv = _linear_wgrad(dq2, x2, wq.shape, d_wqkv_)
        return d_x.view(*Bp_T_L, C), d_table, d_wqkv, d_bqkv, d_wproj, d_bproj, None, None


class _BlockFn(torch.autograd.Function):
    """One post-norm SwinTransformerBlock (swin_512.py:196-237) on bf16 tokens [Bp, T, H*W, C]:

        y   = x + proj(attn_core(qkv(x)))           residual fused into the proj GEMM epilogue
        z   = y + fc2(gelu(fc1(norm2(y))))          GELU / residual fused into the GEMM epilogues
        out = norm1(z)
    """

    @staticmethod
    def forward(ctx, x, table, w_qkv, b_qkv, w_proj, b_proj, g1, be1, g2, be2, w_fc1, b_fc1, w_fc2, b_fc2, geom):
        H, W, nH, ws, shift, qk_scale, eps, infer = geom
        geom = geom[:7]
        Bp, T, L, C = x.shape
        x2 = x.reshape(-1, C)
        wq, wp, w1, w2 = _w16(w_qkv), _w16(w_proj), _w16(w_fc1), _w16(w_fc2)
        qkv = ops.gemm(x2, wq, bias=b_qkv)
        attn, lse2 = ops.winattn_fwd(qkv.view(Bp, T, L, 3 * C), table, H, W, nH, ws, shift, qk_scale=qk_scale)
        y = ops.gemm(attn.view(-1, C), wp, bias=b_proj, aux=x2, mode=ops.EPI_BIAS_RES)
        yn, mean2, rstd2 = ops.layernorm_fwd(y, g2, be2, eps)
        if infer or not any(ctx.needs_input_grad):
            # inference (no_grad key encoders of the pre-training model, evaluation): no gelu' output,
            # nothing kept for a backward
            h = ops.gemm(yn, w1, bias=b_fc1, mode=ops.EPI_BIAS_GELU_FWD)
            z = ops.gemm(h, w2, bias=b_fc2, aux=y, mode=ops.EPI_BIAS_RES)
            out, _, _ = ops.layernorm_fwd(z, g1, be1, eps)
            return out.view(Bp, T, L, C)
        # gelu'(fc1 out) for the backward, one byte per element (a bounded multiplier; fc1 + GELU is HBM-write bound)
        q8 = GELU_GRAD_Q8 and w1.shape[0] % 16 == 0
        dgelu = torch.empty((x2.shape[0], w1.shape[0]), dtype=torch.uint8 if q8 else _BF16, device=x.device)
        h = ops.gemm(yn, w1, bias=b_fc1, mode=ops.EPI_BIAS_GELU_Q8 if q8 else ops.EPI_BIAS_GELU, out2=dgelu)
        z = ops.gemm(h, w2, bias=b_fc2, aux=y, mode=ops.EPI_BIAS_RES)
        out, mean1, rstd1 = ops.layernorm_fwd(z, g1, be1, eps)
        ctx.save_for_backward(x2, qkv, attn, lse2, y, yn, mean2, rstd2, dgelu, h, z, mean1, rstd1,
                              table, wq, wp, w1, w2, g1, g2)
        ctx.geom = geom
        ctx.has_qkv_bias = b_qkv is not None
        return out.view(Bp, T, L, C)

    @staticmethod
    def backward(ctx, d_out):
        (x2, qkv, attn, lse2, y, yn, mean2, rstd2, dgelu, h, z, mean1, rstd1,
         table, wq, wp, w1, w2, g1, g2) = ctx.saved_tensors
        H, W, nH, ws, shift, qk_scale, eps = ctx.geom
        C = x2.shape[1]
        shp = d_out.shape[:3]
        dev = x2.device
        d2 = d_out.contiguous().view(-1, C)
        (d_g1, d_be1, d_bfc2, d_bfc1, d_g2, d_be2, d_bproj, d_table, d_bqkv, d_wfc2_, d_wfc1_, d_wproj_, d_wqkv_) = \
            _zeros_like_many([(C,), (C,), (C,), (w1.shape[0],), (C,), (C,), (C,), tuple(table.shape), (3 * C,),
                              tuple(w2.shape), tuple(w1.shape), tuple(wp.shape), tuple(wq.shape)], dev)
        if not ctx.has_qkv_bias:
            d_bqkv = None
        # out = norm1(z)
        dz = ops.layernorm_bwd(d2, z, mean1, rstd1, g1, d_g1, d_be1, dx_colsum=d_bfc2)
        # z = y + h W2^T + b2 ; h = gelu(u), and the forward stored gelu'(u)
        du = ops.gemm(dz, w2, b_mn_major=True, mode=ops.EPI_MUL_AUX_Q8 if dgelu.dtype == torch.uint8 else ops.EPI_MUL_AUX,
                      aux=dgelu, colsum=d_bfc1)
        d_wfc2 = _linear_wgrad(dz, h, w2.shape, d_wfc2_)
        # u = yn W1^T + b1 ; yn = norm2(y)
        dyn = ops.gemm(du, w1, b_mn_major=True)
        d_wfc1 = _linear_wgrad(du, yn, w1.shape, d_wfc1_)
        dy = ops.layernorm_bwd(dyn, y, mean2, rstd2, g2, d_g2, d_be2, dres=dz, dx_colsum=d_bproj)
        # y = x + attn Wp^T + bp
        d_attn = ops.gemm(dy, wp, b_mn_major=True)
        d_wproj = _linear_wgrad(dy, attn.view(-1, C), wp.shape, d_wproj_)
        d_qkv = ops.winattn_bwd(qkv.view(*shp, 3 * C), table, lse2, d_attn.view(*shp, C), H, W, nH, ws, shift,
                                d_table, d_bqkv, qk_scale=qk_scale)
        dq2 = d_qkv.view(-1, 3 * C)
        d_x = None
        if ctx.needs_input_grad[0]:          # the first block of a head fed with features that carry no gradient skips this GEMM
            d_x = ops.gemm(dq2, wq, b_mn_major=True, mode=ops.EPI_BIAS_RES, aux=dy).view(*shp, C)      # + residual path
        d_wqkv = _linear_wgrad(dq2, x2, wq.shape, d_wqkv_)
        return (d_x, d_table, d_wqkv, d_bqkv, d_wproj, d_bproj, d_g1, d_be1, d_g2, d_be2,
                d_wfc1, d_bfc1, d_wfc2, d_bfc2, None)


class _PatchMergeFn(torch.autograd.Function):
    """2x2 gather + LayerNorm(4C) in one pass, then the bias-free reduction GEMM (swin_512.py:255-277)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, w_red, geom):
        H, W, eps = geom
        B, T, L, C = x.shape
        xn, mean, rstd = ops.layernorm_fwd(x.reshape(B * T, L, C), gamma, beta, eps, patch_merge_hw=(H, W))
        wr = _w16(w_red)
        out = ops.gemm(xn, wr)
        ctx.save_for_backward(x, xn, mean, rstd, gamma, wr)
        ctx.geom = geom
        return out.view(B, T, L // 4, wr.shape[0])

    @staticmethod
    def backward(ctx, d_out):
        x, xn, mean, rstd, gamma, wr = ctx.saved_tensors
        H, W, eps = ctx.geom
        B, T, L, C = x.shape
        d2 = d_out.contiguous().view(-1, wr.shape[0])
        dxn = ops.gemm(d2, wr, b_mn_major=True)
        d_w = _linear_wgrad(d2, xn, wr.shape)
        d_gamma = torch.zeros(4 * C, dtype=torch.float32, device=x.device)
        d_beta = torch.zeros_like(d_gamma)
        dx = ops.layernorm_bwd(dxn, x.reshape(B * T, L, C), mean, rstd, gamma, d_gamma, d_beta, patch_merge_hw=(H, W))
        return dx.view(B, T, L, C), d_gamma, d_beta, d_w, None


class _TransposeFn(torch.autograd.Function):
    """[batch, R, Cc] -> [batch, Cc, R] with dtype change; the backward is the inverse move."""

    @staticmethod
    def forward(ctx, x, out_dtype):
        ctx.in_dtype = x.dtype
        return ops.transpose(x.contiguous(), out_dtype)

    @staticmethod
    def backward(ctx, g):
        return ops.transpose(g.contiguous(), ctx.in_dtype), None


class _FnCtx:
    """Minimal stand-in for an autograd context, used to run ``_BlockFn`` inside another Function."""

    def __init__(self, n_inputs: int):
        self.needs_input_grad = (True,) * n_inputs
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


class _MidLayerFn(torch.autograd.Function):
    """The middle layer of a stage (swin_512.py:302-307): ``cat([x[:, :1], blk1(blk0(x[:, 1:3])), x[:, 3:]])`` as ONE
    autograd node.  Forward: frames 1..2 are copied out, run through the two blocks, and the output is assembled from
    the block result and frames 0 / 3 of x (no ``cat``).  Backward: the incoming gradient is only read; the gradient of
    x is a fresh buffer filled from three sources (frames 0 / 3 of d_out, the blocks' input gradient for frames 1..2),
    so there is no zero-fill and no gradient accumulation for the pass-through frames."""

    N_BLOCK_ARGS = 13

    @staticmethod
    def forward(ctx, x, geom0, geom1, *params):
        n = _MidLayerFn.N_BLOCK_ARGS
        B, _, L, C = x.shape
        xm = torch.empty((B, 2, L, C), dtype=x.dtype, device=x.device)
        ops.copy_frames(xm, x[:, 1:3])
        c0, c1 = _FnCtx(n + 2), _FnCtx(n + 2)
        blk = _BlockFn32 if x.dtype == torch.float32 else _BlockFn
        ctx.blk = blk
        y = blk.forward(c1, blk.forward(c0, xm, *params[:n], geom0), *params[n:], geom1)
        out = torch.empty_like(x)
        ops.copy_frames(out[:, 0], x[:, 0])
        ops.copy_frames(out[:, 3], x[:, 3])
        ops.copy_frames(out[:, 1:3], y)
        ctx.blocks = (c0, c1)
        return out

    @staticmethod
    def backward(ctx, d_out):
        c0, c1 = ctx.blocks
        d_out = d_out.contiguous()
        B, _, L, C = d_out.shape
        d_y = torch.empty((B, 2, L, C), dtype=d_out.dtype, device=d_out.device)
        ops.copy_frames(d_y, d_out[:, 1:3])
        g1 = ctx.blk.backward(c1, d_y)
        g0 = ctx.blk.backward(c0, g1[0])
        d_x = torch.empty_like(d_out)
        ops.copy_frames(d_x[:, 0], d_out[:, 0])
        ops.copy_frames(d_x[:, 3], d_out[:, 3])
        ops.copy_frames(d_x[:, 1:3], g0[0].contiguous())
        return (d_x, None, None, *g0[1:-1], *g1[1:-1])


class _MlpFn(torch.autograd.Function):
    """Stand-alone Mlp.forward (swin_512.py:17-23) on bf16 rows [..., C_in]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x2 = x.reshape(-1, x.shape[-1])
        w1b, w2b = _w16(w1), _w16(w2)
        q8 = GELU_GRAD_Q8 and w1b.shape[0] % 16 == 0
        dgelu = torch.empty((x2.shape[0], w1b.shape[0]), dtype=torch.uint8 if q8 else _BF16, device=x.device)
        h = ops.gemm(x2, w1b, bias=b1, mode=ops.EPI_BIAS_GELU_Q8 if q8 else ops.EPI_BIAS_GELU, out2=dgelu)
        y = ops.gemm(h, w2b, bias=b2)
        ctx.save_for_backward(x2, h, dgelu, w1b, w2b)
        return y.view(*x.shape[:-1], w2b.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, h, dgelu, w1b, w2b = ctx.saved_tensors
        d2 = dy.contiguous().view(-1, w2b.shape[0])
        d_b2, d_b1, d_w2_, d_w1_ = _zeros_like_many([(w2b.shape[0],), (w1b.shape[0],), tuple(w2b.shape), tuple(w1b.shape)], d2.device)
        ops.colsum(d2, d_b2)
        du = ops.gemm(d2, w2b, b_mn_major=True, mode=ops.EPI_MUL_AUX_Q8 if dgelu.dtype == torch.uint8 else ops.EPI_MUL_AUX,
                      aux=dgelu, colsum=d_b1)
        d_w2 = _linear_wgrad(d2, h, w2b.shape, d_w2_)
        dx = ops.gemm(du, w1b, b_mn_major=True)
        d_w1 = _linear_wgrad(du, x2, w1b.shape, d_w1_)
        return dx.view(*dy.shape[:-1], w1b.shape[1]), d_w1, d_b1, d_w2, d_b2


# ---------------------------------------------------------------------------------------------------------------
# fp32-accurate mode (ops32): the reference's no-AMP run at <= 1e-3
# ---------------------------------------------------------------------------------------------------------------
_PRECISION = "bf16"


def set_precision(mode: str) -> str:
    """Arithmetic of the drop-in modules: ``"bf16"`` (default: bf16 storage, fp32 accumulation -- the counterpart of the
    reference under ``amp.autocast``), ``"fp32"`` (activations stay fp32, dense layers as split-bf16 tcgen05 GEMMs, the
    attention core as fp32 SIMT kernels: <= 1e-3 against the reference's fp32 run) or ``"auto"`` (fp32 for fp32 inputs
    outside autocast, bf16 otherwise).  A module's own ``precision`` attribute, when set, wins.  Returns the old mode."""
    global _PRECISION
    if mode not in ("bf16", "fp32", "auto"):
        raise ValueError("precision must be 'bf16', 'fp32' or 'auto'")
    old, _PRECISION = _PRECISION, mode
    return old


def _use_fp32(module, x: torch.Tensor) -> bool:
    mode = getattr(module, "precision", None) or _PRECISION
    if mode == "auto":
        return x.dtype == torch.float32 and not torch.is_autocast_enabled()
    return mode == "fp32"


def _zeros_f32(shape, dev):
    return torch.zeros(shape, dtype=torch.float32, device=dev)


class _AttentionFn32(torch.autograd.Function):
    """``_AttentionFn`` in the fp32-accurate mode: x [Bp, T, H*W, C] fp32."""

    @staticmethod
    def forward(ctx, x, table, w_qkv, b_qkv, w_proj, b_proj, geom, mask=None):
        from . import ops32 as o32
        H, W, nH, ws, shift, qk_scale = geom
        Bp, T, L, C = x.shape
        x2 = x.reshape(-1, C)
        qkv = o32.linear(x2, w_qkv, b_qkv)
        attn, lse = o32.winattn_fwd(qkv.view(Bp, T, L, 3 * C), table, H, W, nH, ws, shift, qk_scale, mask)
        out = o32.linear(attn.view(-1, C), w_proj, b_proj)
        ctx.save_for_backward(x2, qkv, attn, lse, table, w_qkv, w_proj)
        ctx.geom, ctx.mask, ctx.has_qkv_bias = geom, mask, b_qkv is not None
        return out.view(Bp, T, L, C)

    @staticmethod
    def backward(ctx, d_out):
        from . import ops32 as o32
        x2, qkv, attn, lse, table, w_qkv, w_proj = ctx.saved_tensors
        H, W, nH, ws, shift, qk_scale = ctx.geom
        C = x2.shape[1]
        shp = d_out.shape[:3]
        d2 = d_out.contiguous().view(-1, C)
        dev = d2.device
        d_bproj = o32.colsum(d2, _zeros_f32((C,), dev))
        d_attn = o32.linear_dgrad(d2, w_proj)
        d_wproj = o32.linear_wgrad(d2, attn.view(-1, C))
        d_table = _zeros_f32(tuple(table.shape), dev)
        d_qkv = o32.winattn_bwd(qkv.view(*shp, 3 * C), table, attn, lse, d_attn.view(*shp, C), H, W, nH, ws, shift, d_table,
                                qk_scale, ctx.mask)
        dq2 = d_qkv.view(-1, 3 * C)
        d_bqkv = o32.colsum(dq2, _zeros_f32((3 * C,), dev)) if ctx.has_qkv_bias else None
        d_x = o32.linear_dgrad(dq2, w_qkv)
        d_wqkv = o32.linear_wgrad(dq2, x2)
        return d_x.view(*shp, C), d_table, d_wqkv, d_bqkv, d_wproj, d_bproj, None, None


class _BlockFn32(torch.autograd.Function):
    """``_BlockFn`` (one post-norm SwinTransformerBlock, swin_512.py:196-237) in the fp32-accurate mode."""

    @staticmethod
    def forward(ctx, x, table, w_qkv, b_qkv, w_proj, b_proj, g1, be1, g2, be2, w_fc1, b_fc1, w_fc2, b_fc2, geom):
        from . import ops32 as o32
        H, W, nH, ws, shift, qk_scale, eps, infer = geom
        Bp, T, L, C = x.shape
        x2 = x.reshape(-1, C)
        qkv = o32.linear(x2, w_qkv, b_qkv)
        attn, lse = o32.winattn_fwd(qkv.view(Bp, T, L, 3 * C), table, H, W, nH, ws, shift, qk_scale)
        y = o32.linear(attn.view(-1, C), w_proj, b_proj, res=x2)
        yn, mean2, rstd2 = o32.layernorm_fwd(y, g2, be2, eps)
        u = o32.linear(yn, w_fc1, b_fc1)
        z = o32.linear(u, w_fc2, b_fc2, res=y, gelu_input=True)
        out, mean1, rstd1 = o32.layernorm_fwd(z, g1, be1, eps)
        if not infer:
            ctx.save_for_backward(x2, qkv, attn, lse, y, yn, mean2, rstd2, u, z, mean1, rstd1, table, w_qkv, w_proj, w_fc1, w_fc2, g1, g2)
            ctx.geom, ctx.has_qkv_bias = geom[:7], b_qkv is not None
        return out.view(Bp, T, L, C)

    @staticmethod
    def backward(ctx, d_out):
        from . import ops32 as o32
        (x2, qkv, attn, lse, y, yn, mean2, rstd2, u, z, mean1, rstd1, table, w_qkv, w_proj, w_fc1, w_fc2, g1, g2) = ctx.saved_tensors
        H, W, nH, ws, shift, qk_scale, eps = ctx.geom
        C = x2.shape[1]
        shp = d_out.shape[:3]
        dev = x2.device
        d2 = d_out.contiguous().view(-1, C)
        d_g1, d_be1, d_g2, d_be2 = (_zeros_f32((C,), dev) for _ in range(4))
        dz = o32.layernorm_bwd(d2, z, mean1, rstd1, g1, d_g1, d_be1)
        d_bfc2 = o32.colsum(dz, _zeros_f32((C,), dev))
        dh = o32.linear_dgrad(dz, w_fc2)
        d_wfc2 = o32.linear_wgrad(dz, u, gelu_input=True)
        du = o32.mul_gelu_grad(dh, u)
        d_bfc1 = o32.colsum(du, _zeros_f32((w_fc1.shape[0],), dev))
        dyn = o32.linear_dgrad(du, w_fc1)
        d_wfc1 = o32.linear_wgrad(du, yn)
        dy = o32.layernorm_bwd(dyn, y, mean2, rstd2, g2, d_g2, d_be2, dres=dz)
        d_bproj = o32.colsum(dy, _zeros_f32((C,), dev))
        d_attn = o32.linear_dgrad(dy, w_proj)
        d_wproj = o32.linear_wgrad(dy, attn.view(-1, C))
        d_table = _zeros_f32(tuple(table.shape), dev)
        d_qkv = o32.winattn_bwd(qkv.view(*shp, 3 * C), table, attn, lse, d_attn.view(*shp, C), H, W, nH, ws, shift, d_table, qk_scale)
        dq2 = d_qkv.view(-1, 3 * C)
        d_bqkv = o32.colsum(dq2, _zeros_f32((3 * C,), dev)) if ctx.has_qkv_bias else None
        d_x = o32.linear_dgrad(dq2, w_qkv, res=dy)
        d_wqkv = o32.linear_wgrad(dq2, x2)
        return (d_x.view(*shp, C), d_table, d_wqkv, d_bqkv, d_wproj, d_bproj, d_g1, d_be1, d_g2, d_be2,
                d_wfc1, d_bfc1, d_wfc2, d_bfc2, None)


class _PatchMergeFn32(torch.autograd.Function):
    """``_PatchMergeFn`` in the fp32-accurate mode."""

    @staticmethod
    def forward(ctx, x, gamma, beta, w_red, geom):
        from . import ops32 as o32
        H, W, eps = geom
        B, T, L, C = x.shape
        xn, mean, rstd = o32.layernorm_fwd(x.reshape(B * T, L, C), gamma, beta, eps, patch_merge_hw=(H, W))
        out = o32.linear(xn, w_red)
        ctx.save_for_backward(x, xn, mean, rstd, gamma, w_red)
        ctx.geom = geom
        return out.view(B, T, L // 4, w_red.shape[0])

    @staticmethod
    def backward(ctx, d_out):
        from . import ops32 as o32
        x, xn, mean, rstd, gamma, w_red = ctx.saved_tensors
        H, W, eps = ctx.geom
        B, T, L, C = x.shape
        d2 = d_out.contiguous().view(-1, w_red.shape[0])
        dxn = o32.linear_dgrad(d2, w_red)
        d_w = o32.linear_wgrad(d2, xn)
        d_gamma, d_beta = _zeros_f32((4 * C,), x.device), _zeros_f32((4 * C,), x.device)
        dx = o32.layernorm_bwd(dxn, x.reshape(B * T, L, C), mean, rstd, gamma, d_gamma, d_beta, patch_merge_hw=(H, W))
        return dx.view(B, T, L, C), d_gamma, d_beta, d_w, None


class _MlpFn32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        from . import ops32 as o32
        x2 = x.reshape(-1, x.shape[-1])
        u = o32.linear(x2, w1, b1)
        y = o32.linear(u, w2, b2, gelu_input=True)
        ctx.save_for_backward(x2, u, w1, w2)
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        from . import ops32 as o32
        x2, u, w1, w2 = ctx.saved_tensors
        d2 = dy.contiguous().view(-1, w2.shape[0])
        d_b2 = o32.colsum(d2, _zeros_f32((w2.shape[0],), d2.device))
        du = o32.mul_gelu_grad(o32.linear_dgrad(d2, w2), u)
        d_w2 = o32.linear_wgrad(d2, u, gelu_input=True)
        d_b1 = o32.colsum(du, _zeros_f32((w1.shape[0],), d2.device))
        return o32.linear_dgrad(du, w1).view(*dy.shape[:-1], w1.shape[1]), o32.linear_wgrad(du, x2), d_b1, d_w2, d_b2


def _as_f32(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise StswinError("stswincl_b200 modules need CUDA tensors (no CPU path)")
    return x if x.dtype == torch.float32 else x.float()


def _as_tokens_bf16(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise StswinError("stswincl_b200 modules need CUDA tensors (no CPU path)")
    return x if x.dtype == _BF16 else x.to(_BF16)


class Mlp(nn.Module):
    """Parameter container with the reference's names (swin_512.py:7-23); the math runs fused
    inside the block (fc1 + GELU epilogue, fc2 + residual epilogue)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("only nn.GELU (erf) is implemented, as used by every reference config")
        if drop != 0.:
            raise NotImplementedError("dropout > 0 is not used by any reference config and is not implemented")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        """fc1 -> GELU (erf) -> fc2 on the last dim (swin_512.py:17-23), stand-alone: two GEMMs with the bias / GELU
        epilogues fused.  Inside ``SwinTransformerBlock`` the same GEMMs run with the residual fused as well."""
        in_dtype = x.dtype
        if _use_fp32(self, x):
            y = _MlpFn32.apply(_as_f32(x).contiguous(), self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
            return y if in_dtype == torch.float32 else y.to(in_dtype)
        y = _MlpFn.apply(_as_tokens_bf16(x).contiguous(), self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        return y if in_dtype == _BF16 else y.to(in_dtype)


class WindowAttention(nn.Module):
    """swin_512.py:73-141.  ``forward(x_v [B_, T, N, C], mask=None) -> [B_, T, N, C]``.

    Called stand-alone, every window is attended independently (it is treated as a ws x ws image
    with one un-shifted window); an explicit ``mask`` [nW, N, N] is added to the logits of window
    ``b_ % nW`` exactly as at :127-131.  ``SwinTransformerBlock`` does not go through this path: its
    kernels rebuild the shift mask from the geometry."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.window_size = to_2tuple(window_size)
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("dropout > 0 is not used by any reference config and is not implemented")
        if self.window_size[0] != self.window_size[1]:
            raise NotImplementedError("square windows only (the reference always passes to_2tuple(ws))")
        ws = self.window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(ws, ws))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def _qk_scale_arg(self) -> float:
        default = (self.dim // self.num_heads) ** -0.5
        return 0.0 if abs(self.scale - default) < 1e-12 else float(self.scale)

    def forward(self, x_v, mask=None):
        B_, T, N, C = x_v.shape
        if mask is not None:
            assert mask.shape[1:] == (N, N) and B_ % mask.shape[0] == 0, "mask must be [nW, N, N] with B_ a multiple of nW"
            mask = mask.detach().to(device=x_v.device, dtype=torch.float32).contiguous()
        ws = self.window_size[0]
        assert N == ws * ws and C == self.dim, "input feature has wrong size"
        in_dtype = x_v.dtype
        geom = (ws, ws, self.num_heads, ws, 0, self._qk_scale_arg())
        args = (self.relative_position_bias_table, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, geom, mask)
        if _use_fp32(self, x_v):
            out = _AttentionFn32.apply(_as_f32(x_v).contiguous(), *args)
            return out if in_dtype == torch.float32 else out.to(in_dtype)
        out = _AttentionFn.apply(_as_tokens_bf16(x_v).contiguous(), *args)
        return out if in_dtype == _BF16 else out.to(in_dtype)


class SwinTransformerBlock(nn.Module):
    """swin_512.py:143-237.  ``forward(x_v [B, T, L, C]) -> [B, T, L, C]`` (same dtype as the input).

    The reference asserts T == 2; T == 1 is also accepted here as long as
    T * window_size**2 is 16, 32, 64 or 128."""

    def __init__(self, dim, input_resolution, num_heads, window_size=8, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.input_resolution = tuple(input_resolution)
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("only nn.LayerNorm is implemented")
        if drop_path != 0.:
            raise NotImplementedError("drop_path > 0 is not used by any reference config and is not implemented")
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if self.shift_size > 0:
            H, W = self.input_resolution
            attn_mask = _shift_attn_mask(H, W, self.window_size, self.shift_size)
        else:
            attn_mask = None
        self.register_buffer("attn_mask", attn_mask)

    def _geom(self):
        H, W = self.input_resolution
        return (H, W, self.num_heads, self.window_size, self.shift_size, self.attn._qk_scale_arg(), self.norm1.eps)

    def _block_params(self):
        a, m = self.attn, self.mlp
        return (a.relative_position_bias_table, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)

    def forward_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """bf16 in, bf16 out; no dtype round trip (used by the layer container)."""
        # grad mode is decided here: inside autograd.Function.forward it always reads "disabled", and
        # ctx.needs_input_grad ignores torch.no_grad()
        infer = not torch.is_grad_enabled()
        fn = _BlockFn32 if x.dtype == torch.float32 else _BlockFn          # fp32 tokens: the fp32-accurate mode
        return fn.apply(x, *self._block_params(), self._geom() + (infer,))

    def forward(self, x_v):
        H, W = self.input_resolution
        B, T, L, C = x_v.shape
        assert L == H * W, "input feature has wrong size"
        in_dtype = x_v.dtype
        tokens = _as_f32(x_v) if _use_fp32(self, x_v) else _as_tokens_bf16(x_v)
        out = self.forward_tokens(tokens.contiguous())
        return out if in_dtype == out.dtype else out.to(in_dtype)


class PatchMerging(nn.Module):
    """swin_512.py:239-277.  ``forward(x [B, T, L, C]) -> [B, T, L/4, 2C]``."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = tuple(input_resolution)
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward_tokens(self, x: torch.Tensor) -> torch.Tensor:
        H, W = self.input_resolution
        fn = _PatchMergeFn32 if x.dtype == torch.float32 else _PatchMergeFn
        return fn.apply(x, self.norm.weight, self.norm.bias, self.reduction.weight, (H, W, self.norm.eps))

    def forward(self, x):
        H, W = self.input_resolution
        B, T, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        in_dtype = x.dtype
        tokens = _as_f32(x) if _use_fp32(self, x) else _as_tokens_bf16(x)
        out = self.forward_tokens(tokens.contiguous())
        return out if in_dtype == out.dtype else out.to(in_dtype)


class SwinTransformerLayerv5(nn.Module):
    """swin_512.py:280-327.  ``forward(x [B, 4, C, H, W]) -> (x3 [B,4,C,H,W], x6 [B,4,2C,H/2,W/2])``.

    Differences in mechanism, not in result: tokens stay bf16 and token-major between blocks;
    the two frame pairs of layers 0 / 2 (which are independent, :302-307) run as one batch of 2B;
    the NCHW <-> token-major moves are single transposing kernels."""

    def __init__(self, dim=512, input_resolution=(64, 80), num_heads=4):
        super().__init__()
        self.dim = dim
        self.input_resolution = tuple(input_resolution)
        self.num_heads = num_heads
        self.num_layers = 3
        self.pairs = [[slice(0, 2), slice(2, 4)], [slice(1, 3)], [slice(0, 2), slice(2, 4)]]  # t=4
        H, W = self.input_resolution
        self.layers = nn.ModuleList()
        for _ in range(self.num_layers):
            self.layers.append(nn.Sequential(
                SwinTransformerBlock(dim, (H, W), num_heads),
                SwinTransformerBlock(dim, (H, W), num_heads, shift_size=4)))
        for _ in range(self.num_layers):
            self.layers.append(nn.Sequential(
                SwinTransformerBlock(dim * 2, (H // 2, W // 2), num_heads, window_size=4),
                SwinTransformerBlock(dim * 2, (H // 2, W // 2), num_heads, window_size=4, shift_size=2)))
        self.downsample = PatchMerging((H, W), dim)

    def _run_layer(self, x: torch.Tensor, layer_idx: int, both_pairs: bool) -> torch.Tensor:
        """x [B, 4, L, C] bf16 -> same.  ``both_pairs``: frames (0,1) and (2,3); else frames (1,2),
        with frames 0 and 3 passing through (swin_512.py:302-307)."""
        B, T, L, C = x.shape
        blk0, blk1 = self.layers[layer_idx][0], self.layers[layer_idx][1]
        if both_pairs:
            y = blk1.forward_tokens(blk0.forward_tokens(x.view(B * 2, 2, L, C)))
            return y.view(B, 4, L, C)
        if not torch.is_grad_enabled():
            xm = torch.empty((B, 2, L, C), dtype=x.dtype, device=x.device)
            ops.copy_frames(xm, x[:, 1:3])
            out = torch.empty_like(x)
            ops.copy_frames(out[:, 0], x[:, 0])
            ops.copy_frames(out[:, 3], x[:, 3])
            ops.copy_frames(out[:, 1:3], blk1.forward_tokens(blk0.forward_tokens(xm)))
            return out
        return _MidLayerFn.apply(x, blk0._geom() + (False,), blk1._geom() + (False,), *blk0._block_params(), *blk1._block_params())

    def _check_input(self, x_v):
        B, T, C, H, W = x_v.shape
        assert T == 4, "input feature has wrong size"
        assert (H, W) == self.input_resolution and C == self.dim, "input feature has wrong size"
        if not x_v.is_cuda:
            raise StswinError("stswincl_b200 modules need CUDA tensors (no CPU path)")
        # outputs keep the input dtype like the reference's (:326); under autocast its last op, LayerNorm, returns fp32
        keep = x_v.dtype if not torch.is_autocast_enabled() else torch.float32
        out_dtype = keep if keep in (torch.float32, _BF16) else torch.float32
        return keep, out_dtype

    def _to_tokens(self, x_v):
        B, T, C, H, W = x_v.shape
        x_in = x_v if x_v.dtype in (torch.float32, _BF16) else x_v.float()
        tok_dtype = torch.float32 if _use_fp32(self, x_v) else _BF16        # fp32 tokens select the fp32-accurate blocks
        return _TransposeFn.apply(x_in.reshape(B * T, C, H * W), tok_dtype).view(B, T, H * W, C)

    def _from_tokens(self, t, stage: int, out_dtype, keep):
        B, T = t.shape[:2]
        H, W = self.input_resolution
        C = self.dim
        if stage == 2:
            H, W, C = H // 2, W // 2, 2 * C
        out = _TransposeFn.apply(t.reshape(B * T, H * W, C), out_dtype).view(B, T, C, H, W)
        return out if keep == out_dtype else out.to(keep)       # fp16 / fp64 callers: cast at the boundary

    def stages(self, like: torch.Tensor):
        """The forward as a chain of four stages, each ``(fn, parameters)`` with ``fn`` mapping a tuple of tensors to
        the next tuple; composing them on ``(x,)`` gives ``forward(x)``.  ``dist.SegmentedStep`` cuts the autograd
        graph at the stage boundaries, so that the gradients of a stage are complete -- and can travel -- while the
        backward of the stages before it still runs.  ``like``: an input tensor (fixes the output dtype)."""
        keep, out_dtype = self._check_input(like)
        L = self.layers

        def s0(x):
            return (self._run_layer(self._to_tokens(x), 0, True),)

        def s1(t):
            return (self._run_layer(self._run_layer(t, 1, False), 2, True),)

        def s2(t):
            return (self._from_tokens(t, 1, out_dtype, keep), self._run_layer(self.downsample.forward_tokens(t.contiguous()), 3, True))

        def s3(out1, u):
            return (out1, self._from_tokens(self._run_layer(self._run_layer(u, 4, False), 5, True), 2, out_dtype, keep))

        return [(s0, list(L[0].parameters())), (s1, list(L[1].parameters()) + list(L[2].parameters())),
                (s2, list(self.downsample.parameters()) + list(L[3].parameters())),
                (s3, list(L[4].parameters()) + list(L[5].parameters()))]

    def forward(self, x_v):
        state = (x_v,)
        for fn, _ in self.stages(x_v):
            state = fn(*state)
        return state
