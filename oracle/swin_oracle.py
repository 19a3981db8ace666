"""Floating-point oracle for HP-1 (spatio-temporal shifted-window attention).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  A functional torch restatement
(CPU, fp32 or fp64, differentiable through autograd) of

  * WindowAttention.forward        seg18/net/Ours/swin_512.py:109-141
  * SwinTransformerBlock.forward   seg18/net/Ours/swin_512.py:196-237
  * Mlp.forward                    seg18/net/Ours/swin_512.py:17-23
  * PatchMerging.forward           seg18/net/Ours/swin_512.py:255-277
  * SwinTransformerLayerv5.forward seg18/net/Ours/swin_512.py:302-327

It is written around the gather/scatter index of SURVEY.md Appx A.3 instead of
roll/partition/reverse copies, takes parameters as a flat ``{state_dict key:
tensor}`` mapping (same key names as the reference modules), and has no module
state.  Pinned against the reference by ``tests/golden/swin_*.npz``.
"""
from __future__ import annotations

import math
from typing import Mapping, Optional

import torch
import torch.nn.functional as F

from . import index_oracle as ix

Params = Mapping[str, torch.Tensor]


def _sub(params: Params, prefix: str) -> dict:
    plen = len(prefix)
    return {k[plen:]: v for k, v in params.items() if k.startswith(prefix)}


def expanded_bias(table: torch.Tensor, ws: int, T: int) -> torch.Tensor:
    """[nH, T*N, T*N]: table gathered by the relative index and tiled over the
    frame pair (swin_512.py:122-124)."""
    idx = torch.from_numpy(ix.relative_position_index(ws))
    N = ws * ws
    b = table[idx.reshape(-1)].reshape(N, N, -1).permute(2, 0, 1)
    return b.repeat(1, T, T)


def attention_core(qkv: torch.Tensor, table: torch.Tensor, ws: int, num_heads: int,
                   mask: Optional[torch.Tensor] = None, T: int = 2,
                   qk_scale: Optional[float] = None) -> torch.Tensor:
    """Per-window multi-head attention on projected tokens (swin_512.py:117-138).

    qkv [B_, L, 3C] with L = T*N, channels ordered [which][head][hd] (:116) -> [B_, L, C].
    ``mask`` is the [nW, N, N] {0,-100} buffer or None; bias and mask are tiled T x T (:124,128).
    """
    B_, L, C3 = qkv.shape
    C = C3 // 3
    hd = C // num_heads
    scale = qk_scale if qk_scale is not None else hd ** -0.5
    qkv = qkv.reshape(B_, L, 3, num_heads, hd)
    q = qkv[:, :, 0] * scale          # scale after the bias add (:119)
    k = qkv[:, :, 1]
    v = qkv[:, :, 2]
    s = torch.einsum("blhd,bmhd->bhlm", q, k)
    s = s + expanded_bias(table, ws, T).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        m = mask.to(s.dtype).repeat(1, T, T)                       # (:128)
        s = (s.reshape(B_ // nW, nW, num_heads, L, L) + m[None, :, None]).reshape(B_, num_heads, L, L)
    pattn = torch.softmax(s, dim=-1)
    return torch.einsum("bhlm,bmhd->blhd", pattn, v).reshape(B_, L, C)


def window_attention(x_win: torch.Tensor, p: Params, ws: int, num_heads: int,
                     mask: Optional[torch.Tensor] = None,
                     qk_scale: Optional[float] = None) -> torch.Tensor:
    """x_win [B_, T, N, C] -> [B_, T, N, C]   (WindowAttention.forward, swin_512.py:109-141).

    Token order inside a window is t-major then row-major spatial."""
    B_, T, N, C = x_win.shape
    qkv = F.linear(x_win.reshape(B_, T * N, C), p["qkv.weight"], p.get("qkv.bias"))
    o = attention_core(qkv, p["relative_position_bias_table"], ws, num_heads, mask, T, qk_scale)
    o = F.linear(o, p["proj.weight"], p["proj.bias"])
    return o.reshape(B_, T, N, C)


def attention_core_unrolled(qkv: torch.Tensor, table: torch.Tensor, input_resolution, ws: int, shift: int,
                            num_heads: int) -> torch.Tensor:
    """The same core on tokens in their natural (un-rolled, un-partitioned) order:
    qkv [B, T, H*W, 3C] -> [B, T, H*W, C].  Gather by the window index of SURVEY Appx A.3
    (= roll + window_partition, swin_512.py:210-218), attend, scatter back to the same
    coordinates (= window_reverse + inverse roll, :224-231).  This is the exact contract of the
    CUDA kernel ``stswin_winattn_fwd``."""
    H, W = input_resolution
    B, T, L, C3 = qkv.shape
    N = ws * ws
    nW = (H // ws) * (W // ws)
    gidx = torch.from_numpy(ix.window_gather_index(H, W, ws, shift)).reshape(-1)   # [nW*N]
    mask_np = ix.shift_attn_mask(H, W, ws, shift)
    mask = torch.from_numpy(mask_np) if mask_np is not None else None
    qw = qkv[:, :, gidx, :].reshape(B, T, nW, N, C3).permute(0, 2, 1, 3, 4).reshape(B * nW, T * N, C3)
    ow = attention_core(qw, table, ws, num_heads, mask, T)
    ow = ow.reshape(B, nW, T, N, C3 // 3).permute(0, 2, 1, 3, 4).reshape(B, T, nW * N, C3 // 3)
    out = torch.empty_like(ow)
    out[:, :, gidx, :] = ow
    return out


def mlp(x: torch.Tensor, p: Params) -> torch.Tensor:
    """fc1 -> exact (erf) GELU -> fc2   (swin_512.py:17-23, nn.GELU default)."""
    h = F.linear(x, p["fc1.weight"], p["fc1.bias"])
    h = 0.5 * h * (1.0 + torch.erf(h * (1.0 / math.sqrt(2.0))))
    return F.linear(h, p["fc2.weight"], p["fc2.bias"])


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def swin_block(x: torch.Tensor, p: Params, input_resolution, num_heads: int,
               window_size: int = 8, shift_size: int = 0) -> torch.Tensor:
    """x [B, T, L, C] -> [B, T, L, C]   (swin_512.py:196-237).

    Post-norm variant: y = x + Attn(x); out = norm1(y + mlp(norm2(y))).
    T is not restricted here (the reference asserts T == 2 at :201 but its body
    is T-generic).
    """
    H, W = input_resolution
    ws, shift = ix.effective_window(input_resolution, window_size, shift_size)
    B, T, L, C = x.shape
    assert L == H * W, "input feature has wrong size"
    # the qkv / proj Linears act per token, so they commute with the window gather / scatter
    a = _sub(p, "attn.")
    qkv = F.linear(x, a["qkv.weight"], a.get("qkv.bias"))
    core = attention_core_unrolled(qkv, a["relative_position_bias_table"], (H, W), ws, shift, num_heads)
    attn_out = F.linear(core, a["proj.weight"], a["proj.bias"])

    y = x + attn_out                                                           # (:234)
    z = y + mlp(layer_norm(y, p["norm2.weight"], p["norm2.bias"]), _sub(p, "mlp."))
    return layer_norm(z, p["norm1.weight"], p["norm1.bias"])                   # (:235)


def patch_merging(x: torch.Tensor, p: Params, input_resolution) -> torch.Tensor:
    """x [B, T, L, C] -> [B, T, L/4, 2C]   (swin_512.py:255-277).

    Channel order of the concatenation: (even h, even w), (odd h, even w),
    (even h, odd w), (odd h, odd w).
    """
    H, W = input_resolution
    B, T, L, C = x.shape
    assert L == H * W and H % 2 == 0 and W % 2 == 0
    g = x.reshape(B, T, H // 2, 2, W // 2, 2, C)              # [.., h2, dh, w2, dw, C]
    g = g.permute(0, 1, 2, 4, 5, 3, 6)                        # [.., h2, w2, dw, dh, C]
    g = g.reshape(B, T, L // 4, 4 * C)                        # chunk index = dw*2 + dh
    g = layer_norm(g, p["norm.weight"], p["norm.bias"])
    return F.linear(g, p["reduction.weight"])


PAIRS = ((slice(0, 2), slice(2, 4)), (slice(1, 3),), (slice(0, 2), slice(2, 4)))   # swin_512.py:287


def _pair_layer(x: torch.Tensor, p: Params, layer_idx: int, pairs, res, heads, ws) -> torch.Tensor:
    """_single_layer_forward (swin_512.py:302-307): every pair reads the *input*,
    frames outside any pair pass through."""
    out = x.clone()
    for sl in pairs:
        y = x[:, sl]
        y = swin_block(y, _sub(p, f"layers.{layer_idx}.0."), res, heads, ws, 0)
        y = swin_block(y, _sub(p, f"layers.{layer_idx}.1."), res, heads, ws, ws // 2)
        out[:, sl] = y
    return out


def swin_layer_v5(x: torch.Tensor, p: Params, dim: int = 512, input_resolution=(64, 80),
                  num_heads: int = 4, pairs=PAIRS):
    """x [B, 4, C, H, W] -> (x3 [B,4,C,H,W], x6 [B,4,2C,H/2,W/2])  (swin_512.py:309-327).

    Stage 1: window 8 / shift 4 at (H, W); stage 2: window 4 / shift 2 at
    (H/2, W/2) with 2*dim channels (:289-298).
    """
    B, T, C, H, W = x.shape
    assert C == dim and (H, W) == tuple(input_resolution)
    t = x.permute(0, 1, 3, 4, 2).reshape(B, T, H * W, C)
    for li in range(3):
        t = _pair_layer(t, p, li, pairs[li], (H, W), num_heads, 8)
    out1 = t.permute(0, 1, 3, 2).reshape(B, T, C, H, W)
    t = patch_merging(t, _sub(p, "downsample."), (H, W))
    for li in range(3):
        t = _pair_layer(t, p, 3 + li, pairs[li], (H // 2, W // 2), num_heads, 4)
    out2 = t.permute(0, 1, 3, 2).reshape(B, T, 2 * C, H // 2, W // 2)
    return out1, out2


# ----------------------------------------------------------------------------
# deterministic parameter / input generators shared by goldens, tests and bench
# ----------------------------------------------------------------------------

def _randn(gen: torch.Generator, *shape, std: float = 1.0) -> torch.Tensor:
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std


def make_attention_params(dim: int, ws: int, num_heads: int, seed: int, bias_std: float = 0.5) -> dict:
    """WindowAttention state (keys as the reference's, swin_512.py:81-104).  The
    bias table is scaled to sigma 0.5 so that index bugs are visible."""
    g = torch.Generator().manual_seed(seed)
    return {
        "relative_position_bias_table": _randn(g, (2 * ws - 1) ** 2, num_heads, std=bias_std),
        "relative_position_index": torch.from_numpy(ix.relative_position_index(ws)),
        "qkv.weight": _randn(g, 3 * dim, dim, std=dim ** -0.5),
        "qkv.bias": _randn(g, 3 * dim, std=0.1),
        "proj.weight": _randn(g, dim, dim, std=dim ** -0.5),
        "proj.bias": _randn(g, dim, std=0.1),
    }


def make_block_params(dim: int, input_resolution, num_heads: int, window_size: int, shift_size: int,
                      seed: int, mlp_ratio: float = 4.0) -> dict:
    ws, shift = ix.effective_window(input_resolution, window_size, shift_size)
    g = torch.Generator().manual_seed(seed + 7919)
    hidden = int(dim * mlp_ratio)
    p = {"attn." + k: v for k, v in make_attention_params(dim, ws, num_heads, seed).items()}
    p.update({
        "norm1.weight": 1.0 + _randn(g, dim, std=0.1), "norm1.bias": _randn(g, dim, std=0.1),
        "norm2.weight": 1.0 + _randn(g, dim, std=0.1), "norm2.bias": _randn(g, dim, std=0.1),
        "mlp.fc1.weight": _randn(g, hidden, dim, std=dim ** -0.5), "mlp.fc1.bias": _randn(g, hidden, std=0.1),
        "mlp.fc2.weight": _randn(g, dim, hidden, std=hidden ** -0.5), "mlp.fc2.bias": _randn(g, dim, std=0.1),
    })
    if shift > 0:
        H, W = input_resolution
        p["attn_mask"] = torch.from_numpy(ix.shift_attn_mask(H, W, ws, shift))
    return p


def make_layer_params(dim: int, input_resolution, num_heads: int, seed: int) -> dict:
    """SwinTransformerLayerv5 state_dict (swin_512.py:281-300)."""
    H, W = input_resolution
    p = {}
    for li in range(6):
        d, res, ws = (dim, (H, W), 8) if li < 3 else (2 * dim, (H // 2, W // 2), 4)
        for bi, shift in enumerate((0, ws // 2)):
            blk = make_block_params(d, res, num_heads, ws, shift, seed + 100 * li + 10 * bi)
            p.update({f"layers.{li}.{bi}.{k}": v for k, v in blk.items()})
    g = torch.Generator().manual_seed(seed + 4242)
    p["downsample.reduction.weight"] = _randn(g, 2 * dim, 4 * dim, std=(4 * dim) ** -0.5)
    p["downsample.norm.weight"] = 1.0 + _randn(g, 4 * dim, std=0.1)
    p["downsample.norm.bias"] = _randn(g, 4 * dim, std=0.1)
    return p


def make_features(seed: int, *shape) -> torch.Tensor:
    """relu(N(0,1)) features, the kernel-level synthetic input of SURVEY 8d config 2."""
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(*shape, generator=g, dtype=torch.float32))
