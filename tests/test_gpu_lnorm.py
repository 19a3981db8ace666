"""GPU: LayerNorm fwd/bwd (+ PatchMerging gather) and layout transposes against the CPU oracle."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to(torch.bfloat16)


@pytest.mark.parametrize("M,C", [(1000, 512), (4096, 1024), (300, 128), (777, 256), (512, 2048), (1234, 384), (163840, 512), (9, 64)])
def test_layernorm_fwd_bwd(M, C):
    from oracle import swin_oracle as so
    from stswincl_b200 import ops
    x = _mk((M, C), 1, 2.0, 0.5)
    dy = _mk((M, C), 2)
    dres = _mk((M, C), 3)
    g = torch.Generator().manual_seed(4)
    gamma = 1.0 + 0.2 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    x32 = x.float().requires_grad_(True)
    g32 = gamma.clone().requires_grad_(True)
    b32 = beta.clone().requires_grad_(True)
    ref = so.layer_norm(x32, g32, b32)
    (ref * dy.float()).sum().backward()
    y, mean, rstd = ops.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda())
    assert rel_err(y.float().cpu(), ref) < 1e-2
    dgamma = torch.zeros(C, device="cuda"); dbeta = torch.zeros(C, device="cuda"); cs = torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gamma.cuda(), dgamma, dbeta, dres=dres.cuda(), dx_colsum=cs)
    torch.cuda.synchronize()
    dx_ref = x32.grad + dres.float()
    assert rel_err(dx.float().cpu(), dx_ref) < 1e-2
    assert rel_err(dgamma.cpu(), g32.grad) < 1e-2
    assert rel_err(dbeta.cpu(), b32.grad) < 1e-2
    assert rel_err(cs.cpu(), dx.float().cpu().sum(0)) < 1e-3


@pytest.mark.parametrize("BT,H,W,C", [(4, 16, 24, 128), (8, 64, 80, 512), (3, 8, 12, 64)])
def test_patch_merging_layernorm(BT, H, W, C):
    from oracle import swin_oracle as so
    from stswincl_b200 import ops
    x = _mk((BT, H * W, C), 1, 2.0, 0.5)
    g = torch.Generator().manual_seed(4)
    gamma = 1.0 + 0.2 * torch.randn(4 * C, generator=g)
    beta = 0.2 * torch.randn(4 * C, generator=g)
    dy = _mk((BT * H * W // 4, 4 * C), 2)
    x32 = x.float().requires_grad_(True)
    # oracle: the gather + norm half of patch_merging (identity reduction)
    gat = x32.reshape(1, BT, H // 2, 2, W // 2, 2, C).permute(0, 1, 2, 4, 5, 3, 6).reshape(BT * H * W // 4, 4 * C)
    ref = so.layer_norm(gat, gamma, beta)
    (ref * dy.float()).sum().backward()
    y, mean, rstd = ops.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda(), patch_merge_hw=(H, W))
    assert rel_err(y.float().cpu(), ref) < 1e-2
    dgamma = torch.zeros(4 * C, device="cuda"); dbeta = torch.zeros(4 * C, device="cuda")
    dx = ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gamma.cuda(), dgamma, dbeta, patch_merge_hw=(H, W))
    torch.cuda.synchronize()
    assert rel_err(dx.float().cpu(), x32.grad) < 1e-2
    assert rel_err(dbeta.cpu(), dy.float().sum(0)) < 1e-2


@pytest.mark.parametrize("dt_in,dt_out", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_transpose(dt_in, dt_out):
    from stswincl_b200 import ops
    x = torch.randn(6, 100, 70, generator=torch.Generator().manual_seed(0)).to(dt_in)
    out = ops.transpose(x.cuda(), dt_out)
    torch.cuda.synchronize()
    ref = x.to(torch.bfloat16).float().transpose(1, 2) if torch.bfloat16 in (dt_in, dt_out) else x.transpose(1, 2)
    assert torch.equal(out.float().cpu(), ref.contiguous())


@pytest.mark.parametrize("shape", [(3, 5120, 512), (2, 1280, 1024), (2, 200, 72), (1, 64, 64), (2, 136, 264), (1, 8, 8)])
def test_transpose_bf16_vectorised(shape):
    """bf16 -> bf16 with both extents multiples of 8 takes the register-transpose kernel (8x8 blocks per thread, 128x128 per
    CTA, 16-byte accesses); ragged tiles too."""
    from stswincl_b200 import ops
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    out = ops.transpose(x.cuda(), torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), x.transpose(1, 2).contiguous())


def test_copy_frames_strided_batches():
    """stswin_copy_strided: frame slices of a [B, 4, L, C] tensor in and out (the middle Swin layer's glue)."""
    from stswincl_b200 import ops
    x = torch.randn(3, 4, 37, 64, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).cuda()
    mid = torch.empty(3, 2, 37, 64, dtype=torch.bfloat16, device="cuda")
    ops.copy_frames(mid, x[:, 1:3])
    assert torch.equal(mid, x[:, 1:3])
    out = torch.zeros_like(x)
    ops.copy_frames(out[:, 0], x[:, 0])
    ops.copy_frames(out[:, 3], x[:, 3])
    ops.copy_frames(out[:, 1:3], mid)
    torch.cuda.synchronize()
    assert torch.equal(out, x)
    with pytest.raises(Exception):
        ops.copy_frames(out[:, 0, :, 1:], x[:, 0, :, 1:])        # inner dims not dense

