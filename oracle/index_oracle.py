"""Integer side of the window-attention path, restated as closed forms in numpy.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Everything here must match
the reference bit-for-bit; the fixtures in ``tests/golden/index_*.npz`` were
produced by the reference constructors / ``torch.roll`` + ``window_partition``
+ ``window_reverse`` run on ``arange`` tensors (``oracle/make_goldens.py``).

Reference sites (paths relative to ``/root/reference``):
  * relative_position_index  seg18/net/Ours/swin_512.py:88-99
  * shift attn_mask          seg18/net/Ours/swin_512.py:171-192
  * window_partition         seg18/net/Ours/swin_512.py:26-38
  * window_reverse           seg18/net/Ours/swin_512.py:57-71
  * roll (+/- shift)         seg18/net/Ours/swin_512.py:210-213, 226-229
  * posMask / negMask        pixcontrast_18/contrast/models/PixPro_swin_v5.py:48-69
  * nearest label resize     pixcontrast_18/contrast/models/PixPro_swin_v5.py:585-590
"""
from __future__ import annotations

import numpy as np

MASK_FILL = -100.0  # swin_512.py:190 -- the reference fills with -100, not -inf


def relative_position_index(ws: int) -> np.ndarray:
    """[N, N] int64 with N = ws*ws; entry (i, j) indexes the bias table.

    idx = (h_i - h_j + ws - 1) * (2*ws - 1) + (w_i - w_j + ws - 1)
    with h = i // ws, w = i % ws  (swin_512.py:88-99).
    """
    n = np.arange(ws * ws)
    h, w = n // ws, n % ws
    dh = h[:, None] - h[None, :] + ws - 1
    dw = w[:, None] - w[None, :] + ws - 1
    return (dh * (2 * ws - 1) + dw).astype(np.int64)


def _band(p: np.ndarray, extent: int, ws: int, shift: int) -> np.ndarray:
    # the three slices (0,-ws), (-ws,-shift), (-shift,None) of swin_512.py:174-179
    return (p >= extent - ws).astype(np.int64) + (p >= extent - shift).astype(np.int64)


def shift_region_ids(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """[H, W] int64 region id (0..8) in *shifted* coordinates (swin_512.py:173-184)."""
    rh = _band(np.arange(H), H, ws, shift)
    rw = _band(np.arange(W), W, ws, shift)
    return 3 * rh[:, None] + rw[None, :]


def shift_attn_mask(H: int, W: int, ws: int, shift: int) -> np.ndarray | None:
    """[nW, N, N] float32 in {0, -100}; None when shift == 0 (swin_512.py:171-194)."""
    if shift == 0:
        return None
    ids = shift_region_ids(H, W, ws, shift)
    per_win = (ids.reshape(H // ws, ws, W // ws, ws)
                  .transpose(0, 2, 1, 3).reshape(-1, ws * ws))
    differ = per_win[:, None, :] != per_win[:, :, None]
    return np.where(differ, np.float32(MASK_FILL), np.float32(0.0)).astype(np.float32)


def window_gather_index(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """[nW, N] int64: flat source token (h*W + w) read by window ``win`` slot ``n``.

    Composition of roll(-shift) and window_partition (swin_512.py:210-218):
      src = ((win // nWw)*ws + n // ws + shift) % H , ((win % nWw)*ws + n % ws + shift) % W
    The scatter on the way back (window_reverse + roll(+shift), :224-231) writes
    to the same coordinates.
    """
    nWh, nWw = H // ws, W // ws
    win = np.arange(nWh * nWw)[:, None]
    n = np.arange(ws * ws)[None, :]
    hh = ((win // nWw) * ws + n // ws + shift) % H
    ww = ((win % nWw) * ws + n % ws + shift) % W
    return (hh * W + ww).astype(np.int64)


def effective_window(input_resolution, window_size: int, shift_size: int):
    """Clamp rule of SwinTransformerBlock.__init__ (swin_512.py:154-158)."""
    if min(input_resolution) <= window_size:
        return min(input_resolution), 0
    assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
    return window_size, shift_size


def label_match(l1: np.ndarray, l2: np.ndarray) -> np.ndarray:
    """posMask as a label compare: [B, HW, HW] float32, 1 where labels agree.

    The reference builds one-hot rows and multiplies them (PixPro_swin_v5.py:48-57);
    the product of two one-hot rows is 1 iff the labels are equal.  ``.long()``
    truncates toward zero, hence the astype.
    """
    a = np.trunc(np.asarray(l1, dtype=np.float64)).astype(np.int64).reshape(l1.shape[0], -1)
    b = np.trunc(np.asarray(l2, dtype=np.float64)).astype(np.int64).reshape(l2.shape[0], -1)
    return (a[:, :, None] == b[:, None, :]).astype(np.float32)


def label_differ(l1: np.ndarray, l2: np.ndarray) -> np.ndarray:
    """negMask = 1 - posMask (PixPro_swin_v5.py:59-69)."""
    return np.float32(1.0) - label_match(l1, l2)


def nearest_resize_index(src: int, dst: int) -> np.ndarray:
    """Source index picked by F.interpolate(mode='nearest') for each output index.

    ATen's legacy 'nearest' uses floor(dst_index * src/dst) computed in float32
    (PixPro_swin_v5.py:585-590 calls it with size=[H, W]).  For the reference's
    256x448 -> 32x56 case this is exactly [::8].
    """
    scale = np.float32(src) / np.float32(dst)
    idx = np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, src - 1)


def downsample_labels(mask: np.ndarray, H: int, W: int) -> np.ndarray:
    """[B,1,Hs,Ws] -> [B,1,H,W] nearest (PixPro_swin_v5.py:585-590)."""
    ih = nearest_resize_index(mask.shape[-2], H)
    iw = nearest_resize_index(mask.shape[-1], W)
    return mask[..., ih[:, None], iw[None, :]]
