"""GPU: tcgen05 GEMM (stswin_gemm_bf16) against torch fp32 matmul on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


SHAPES = [(128, 256, 64), (256, 512, 512), (384, 1536, 512), (1000, 520, 200), (128, 256, 2048), (4096, 2048, 512)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("majors", [(False, False), (False, True), (True, True)])
def test_gemm_plain(M, N, K, majors):
    from stswincl_b200 import ops
    a_mn, b_mn = majors
    A = _mk((M, K), 1, K ** -0.5)
    B = _mk((N, K), 2)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    ref = A.float() @ B.float().t() + bias
    a_in = A.t().contiguous() if a_mn else A
    b_in = B.t().contiguous() if b_mn else B
    out = ops.gemm(a_in, b_in, a_mn_major=a_mn, b_mn_major=b_mn, bias=bias)
    torch.cuda.synchronize()
    assert _rel(out.float(), ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (1000, 520, 200)])
def test_gemm_epilogues(M, N, K):
    from stswincl_b200 import ops
    A = _mk((M, K), 1, K ** -0.5)
    B = _mk((N, K), 2)
    R = _mk((M, N), 4)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    acc = A.float() @ B.float().t()
    # residual
    out = ops.gemm(A, B, mode=ops.EPI_BIAS_RES, bias=bias, aux=R)
    assert _rel(out.float(), acc + bias + R.float()) < 1e-2
    # erf-GELU with its derivative as second output + column sums
    cs = torch.zeros(N, device="cuda")
    dg = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    h = ops.gemm(A, B, mode=ops.EPI_BIAS_GELU, bias=bias, out2=dg, colsum=cs)
    uref = (acc + bias).double().requires_grad_(True)
    href = torch.nn.functional.gelu(uref)
    href.sum().backward()
    assert _rel(h.float(), href.detach().float()) < 1e-2
    assert _rel(dg.float(), uref.grad.float()) < 1e-2
    assert _rel(cs, h.float().sum(0)) < 1e-3
    # forward-only GELU (inference path)
    h1 = ops.gemm(A, B, mode=ops.EPI_BIAS_GELU_FWD, bias=bias)
    assert _rel(h1.float(), href.detach().float()) < 1e-2
    # multiply by aux (dgrad through GELU)
    d = ops.gemm(A, B, mode=ops.EPI_MUL_AUX, aux=R)
    assert _rel(d.float(), acc * R.float()) < 1e-2


def test_gemm_gelu_derivative_as_bytes():
    """STSWIN_EPI_BIAS_GELU_Q8 / STSWIN_EPI_MUL_AUX_Q8: GELU' stored as one byte per element (step 1.28 / 255) and consumed by
    the x-aux epilogue of the backward; ragged M / N edges included."""
    from stswincl_b200 import ops
    for (M, N, K) in ((300, 272, 192), (512, 2048, 512)):
        A, B = _mk((M, K), 11), _mk((N, K), 12)
        bias = torch.randn(N, generator=torch.Generator().manual_seed(13)).cuda()
        q = torch.full((M, N), 77, dtype=torch.uint8, device="cuda")
        cs = torch.zeros(N, device="cuda")
        h = ops.gemm(A, B, mode=ops.EPI_BIAS_GELU_Q8, bias=bias, out2=q, colsum=cs)
        u = (A.float() @ B.float().t() + bias).double().requires_grad_(True)
        href = torch.nn.functional.gelu(u)
        href.sum().backward()
        assert _rel(h.float(), href.detach().float()) < 1e-2
        deq = q.float() * (1.28 / 255.0) - 0.14
        assert (deq - u.grad.float()).abs().max() < 0.5 * 1.28 / 255 + 1.5e-3      # half a step + the tanh-form error
        # backward side: D = acc * dequant(aux) (+ column sums), against the same product with the exact derivative
        A2, B2 = _mk((M, 64), 14), _mk((64, N), 15)
        cs2 = torch.zeros(N, device="cuda")
        d = ops.gemm(A2, B2, b_mn_major=True, mode=ops.EPI_MUL_AUX_Q8, aux=q, colsum=cs2)
        acc2 = A2.float() @ B2.float()
        assert _rel(d.float(), acc2 * deq) < 1e-2
        assert _rel(d.float(), acc2 * u.grad.float()) < 1e-2
        assert _rel(cs2, d.float().sum(0)) < 1e-3


def test_gemm_gelu_outliers():
    """Pre-activations far outside the usual range (|u| up to ~60): GELU(u) -> u or 0 and GELU'(u) -> 1 or 0.
    (The polynomial inside the tanh form changes sign beyond |u| ~ 11 unless u^2 is clamped.)"""
    from stswincl_b200 import ops
    M, N, K = 256, 256, 64
    A = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    B = torch.zeros(N, K, dtype=torch.bfloat16, device="cuda")
    bias = torch.linspace(-60.0, 60.0, N, device="cuda")
    dg = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    h = ops.gemm(A, B, mode=ops.EPI_BIAS_GELU, bias=bias, out2=dg)
    u = bias.double().requires_grad_(True)
    href = torch.nn.functional.gelu(u)
    href.sum().backward()
    assert (h.float() - href.detach().float()[None, :]).abs().max() < 0.26      # bf16 spacing at 60 is 0.25
    assert (dg.float() - u.grad.float()[None, :]).abs().max() < 1e-2


@pytest.mark.parametrize("M,N,K,splits", [(512, 512, 8192, 16), (1536, 512, 4096, 6), (200, 300, 1000, 3)])
def test_gemm_wgrad_splitk(M, N, K, splits):
    """dW[M,N] += dy^T x : both operands stored [K, .] (MN-major), fp32 TMA add-reduction."""
    from stswincl_b200 import ops
    if N % 4:
        pytest.skip("fp32 ldd must be a multiple of 4")
    dy = _mk((K, M), 5)
    x = _mk((K, N + (8 - N % 8) % 8), 6)[:, :N]
    base = torch.randn(M, N, generator=torch.Generator().manual_seed(7)).cuda()
    out = base.clone()
    ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, mode=ops.EPI_F32_REDUCE, out=out, k_splits=splits)
    ref = base + dy.float().t() @ x.float()
    assert _rel(out, ref) < 2e-3
