"""Development probe: run the tcgen05 GEMM variants and describe WHERE errors are (rows / cols / k)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops

def mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).cuda()

def describe(out, ref, tag):
    err = (out.double() - ref.double()).abs()
    rel = float(err.max() / ref.double().abs().max())
    bad = err > 1e-2 * ref.abs().max()
    info = {"tag": tag, "rel": rel, "bad_frac": float(bad.float().mean())}
    if bad.any():
        rows = bad.any(1).nonzero().flatten().tolist()
        cols = bad.any(0).nonzero().flatten().tolist()
        info["bad_rows"] = [rows[0], rows[-1], len(rows)]
        info["bad_cols"] = [cols[0], cols[-1], len(cols)]
        info["sample_out"] = out[rows[0], cols[0]:cols[0] + 4].float().tolist()
        info["sample_ref"] = ref[rows[0], cols[0]:cols[0] + 4].float().tolist()
        info["nan"] = bool(torch.isnan(out.float()).any())
    print(json.dumps(info), flush=True)
    return rel

for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 512), (1000, 520, 200)]:
    for (a_mn, b_mn) in [(False, False), (False, True), (True, True)]:
        A = mk((M, K), 1, K ** -0.5); B = mk((N, K), 2)
        ref = A.float() @ B.float().t()
        a_in = A.t().contiguous() if a_mn else A
        b_in = B.t().contiguous() if b_mn else B
        try:
            out = ops.gemm(a_in, b_in, a_mn_major=a_mn, b_mn_major=b_mn)
            torch.cuda.synchronize()
            describe(out.float(), ref, f"M{M}N{N}K{K} a_mn={a_mn} b_mn={b_mn}")
        except Exception as e:
            print("EXC", M, N, K, a_mn, b_mn, repr(e)[:300], flush=True)
            sys.exit(1)

# timing of the forward shapes of one stage-1 / stage-2 block call at B=8 clips (2 pairs)
def bench(M, N, K, **kw):
    A = mk((M, K), 1, K ** -0.5); B = mk((N, K), 2)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for _ in range(3): ops.gemm(A, B, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(A, B, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t0 = time.time()
    for _ in range(3): torch.matmul(A, B.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): torch.matmul(A, B.t())
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / 10
    print(json.dumps({"bench": [M, N, K], "ms": ms, "tflops": 2 * M * N * K / ms / 1e9, "cublas_ms": ms_ref,
                      "cublas_tflops": 2 * M * N * K / ms_ref / 1e9}), flush=True)

for shp in [(163840, 1536, 512), (163840, 512, 512), (163840, 2048, 512), (163840, 512, 2048), (40960, 3072, 1024), (40960, 4096, 1024)]:
    bench(*shp)
