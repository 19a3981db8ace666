"""GPU: window-attention core kernels against the CPU oracle (oracle.swin_oracle.attention_core_unrolled)
on the same seeded, bf16-rounded inputs.  Tolerance: 2e-2 relative (bf16 path, BASELINE north_star)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_BF16 = 2e-2

# (B, T, H, W, C, nH, ws, shift)
CASES = [
    (2, 2, 16, 24, 256, 2, 8, 0),     # L=128, hd=128, unshifted
    (2, 2, 16, 24, 256, 2, 8, 4),     # shifted: wrap windows + mask
    (1, 2, 8, 12, 512, 2, 4, 2),      # stage-2 like: L=32, hd=256, 4 windows per tile
    (1, 2, 8, 12, 256, 2, 4, 2),      # the layer-golden's layers.4.1 call: 6 windows, partial last tile, hd=128
    (3, 2, 8, 12, 256, 4, 4, 0),      # hd=64
    (2, 1, 16, 24, 256, 2, 8, 4),     # T=1: L=64, 2 windows per tile
    (4, 1, 8, 12, 128, 2, 4, 0),      # T=1, ws=4: L=16, 8 windows per tile
    (4, 1, 8, 12, 128, 2, 4, 2),      # ... shifted: 4-row quadrant boxes (512-byte smem offsets)
    (3, 1, 8, 12, 128, 2, 4, 0),      # ... partial last tile
    (3, 1, 8, 12, 128, 2, 4, 2),
    (1, 2, 64, 80, 512, 4, 8, 4),     # the real stage-1 geometry
    (1, 2, 32, 40, 1024, 4, 4, 2),    # the real stage-2 geometry
    # config 5 (CaDIS-shaped 512x960 crop -> 64x120 tokens): window / head-count sweep
    (1, 2, 64, 120, 512, 4, 8, 4),    # hd 128
    (1, 2, 64, 120, 512, 8, 8, 4),    # hd 64
    (1, 2, 64, 120, 512, 8, 4, 2),    # ws 4: 480 windows, 4 per tile
    (1, 1, 64, 120, 512, 4, 8, 0),    # T = 1
    (1, 2, 32, 60, 1024, 4, 4, 2),    # stage 2 of the same crop
    # head_dim 32 (16 heads at C=512): two heads share a 64-channel chunk
    (2, 2, 16, 24, 128, 4, 8, 4),
    (2, 2, 16, 24, 128, 4, 8, 0),
    (1, 2, 8, 12, 256, 8, 4, 2),
    (3, 1, 8, 12, 64, 2, 4, 2),
    (1, 2, 64, 120, 512, 16, 8, 4),
    # irregular geometry: 7x7 windows (98 or 49 tokens, padding rows in the tile), shifts other than ws/2,
    # small windows with many per tile
    (2, 2, 14, 21, 128, 2, 7, 3),
    (2, 2, 14, 21, 128, 2, 7, 0),
    (3, 1, 14, 21, 128, 2, 7, 3),     # T=1: two 49-token windows per tile, odd window count
    (1, 2, 56, 84, 512, 4, 7, 3),     # config 5: ws 7 on the 56x84 crop
    (1, 2, 56, 84, 512, 16, 7, 3),    # ... with 16 heads (head_dim 32)
    (2, 2, 16, 24, 128, 2, 8, 3),     # unequal rectangles (5 | 3)
    (2, 2, 8, 12, 128, 2, 4, 1),
    (2, 2, 9, 12, 64, 1, 3, 1),       # 18-token windows, 7 per tile
    (1, 2, 10, 15, 128, 2, 5, 2),     # 50-token windows
    # odd windows in padded slots (49 of 64, 25 of 32, 50 of 64 rows): compile-time paths, and the look-up-table
    # path on the same slots for shifts other than ws // 2
    (3, 1, 14, 21, 128, 2, 7, 0),
    (1, 1, 56, 84, 512, 16, 7, 3),    # the shape north_star names: 49 tokens, head_dim 32
    (1, 1, 56, 84, 512, 8, 7, 0),
    (3, 1, 10, 15, 128, 2, 5, 2),     # 25-token windows, 4 per tile, 18 windows (partial last tile)
    (3, 1, 10, 15, 128, 2, 5, 0),
    (3, 2, 10, 15, 128, 4, 5, 0),     # 50-token windows, head_dim 32
    (1, 1, 60, 120, 512, 16, 5, 2),
    (2, 1, 14, 21, 128, 2, 7, 2),     # look-up-table path, padded slots
    (2, 2, 10, 15, 128, 2, 5, 1),
    (2, 1, 9, 12, 64, 1, 3, 1),       # 9-token windows, 14 per tile (packed slots)
    (2, 2, 8, 12, 128, 2, 2, 1),      # 2x2 windows: 8 tokens, 16 per tile
    (2, 1, 8, 12, 128, 4, 2, 0),      # 4 tokens, 32 per tile, head_dim 32
]


def make_case(B, T, H, W, C, nH, ws, shift, seed=0):
    g = torch.Generator().manual_seed(seed)
    qkv = (torch.randn(B, T, H * W, 3 * C, generator=g) * 1.2).to(torch.bfloat16)
    table = torch.randn((2 * ws - 1) ** 2, nH, generator=g) * 0.5
    return qkv, table


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_T%d_%dx%d_C%d_h%d_ws%d_s%d" % c)
def test_winattn_fwd_matches_oracle(case):
    from oracle import swin_oracle as so
    from stswincl_b200 import ops
    B, T, H, W, C, nH, ws, shift = case
    qkv, table = make_case(*case)
    ref = so.attention_core_unrolled(qkv.float(), table, (H, W), ws, shift, nH)
    out, lse2 = ops.winattn_fwd(qkv.cuda(), table.cuda(), H, W, nH, ws, shift)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float().cpu(), ref) < TOL_BF16


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_T%d_%dx%d_C%d_h%d_ws%d_s%d" % c)
def test_winattn_bwd_matches_oracle(case):
    from oracle import swin_oracle as so
    from stswincl_b200 import ops
    B, T, H, W, C, nH, ws, shift = case
    qkv, table = make_case(*case)
    g = torch.Generator().manual_seed(5)
    d_out = torch.randn(B, T, H * W, C, generator=g).to(torch.bfloat16)
    q32 = qkv.float().requires_grad_(True)
    t32 = table.clone().requires_grad_(True)
    ref = so.attention_core_unrolled(q32, t32, (H, W), ws, shift, nH)
    (ref * d_out.float()).sum().backward()
    qkv_d, table_d = qkv.cuda(), table.cuda()
    out, lse2 = ops.winattn_fwd(qkv_d, table_d, H, W, nH, ws, shift)
    d_table = torch.zeros_like(table_d)
    colsum = torch.zeros(3 * C, device="cuda")
    d_qkv = ops.winattn_bwd(qkv_d, table_d, lse2, d_out.cuda(), H, W, nH, ws, shift, d_table, colsum)
    torch.cuda.synchronize()
    assert torch.isfinite(d_qkv.float()).all()
    assert rel_err(d_qkv.float().cpu(), q32.grad) < TOL_BF16
    assert rel_err(d_table.cpu(), t32.grad) < TOL_BF16
    assert rel_err(colsum.cpu(), q32.grad.sum((0, 1, 2))) < TOL_BF16


@pytest.mark.parametrize("kw,msg", [
    (dict(C=256, nH=16, ws=8, shift=4), "head_dim 16"),
    (dict(C=512, nH=4, ws=16, shift=0), "tokens per window"),    # 16x16x2 = 512 tokens
    (dict(C=512, nH=4, ws=8, shift=8), "shift 8"),
    (dict(C=512, nH=4, ws=8, shift=4, H=60), "multiples of the window size"),
])
def test_unsupported_geometry_fails_loudly(kw, msg):
    """No silent fallback: shapes outside the kernels' envelope raise with a reason."""
    from stswincl_b200 import ops
    from stswincl_b200._lib import StswinError
    H, W, C, nH, ws, shift = kw.get("H", 64), kw.get("W", 80), kw["C"], kw["nH"], kw["ws"], kw["shift"]
    qkv = torch.zeros(1, 2, H * W, 3 * C, dtype=torch.bfloat16, device="cuda")
    table = torch.zeros((2 * ws - 1) ** 2, nH, device="cuda")
    with pytest.raises(StswinError, match=msg):
        ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)
