"""ncu / timing target: a few launches of one GEMM shape.  usage: prof_gemm.py M N K [mode] [b_mn] [colsum]"""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
M, N, K = (int(a) for a in sys.argv[1:4])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
b_mn = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
use_cs = bool(int(sys.argv[6])) if len(sys.argv) > 6 else (mode in (3, 7))
g = torch.Generator().manual_seed(0)
A = (torch.randn(M, K, generator=g) * K ** -0.5).to(torch.bfloat16).cuda()
B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
if b_mn:
    B = B.t().contiguous()
bias = torch.randn(N, generator=g).cuda()
aux = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda() if mode in (1, 3) else None
if mode == 7:
    aux = torch.randint(0, 256, (M, N), generator=g, dtype=torch.uint8).cuda()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
out2 = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if mode == 2 else None
if mode == 6:
    out2 = torch.empty(M, N, dtype=torch.uint8, device="cuda")
cs = torch.zeros(N, device="cuda") if use_cs else None
kw = dict(b_mn_major=b_mn, mode=mode, bias=None if mode in (3, 7) else bias, aux=aux, out=out, out2=out2, colsum=cs)
for _ in range(4):
    ops.gemm(A, B, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gemm(A, B, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"gemm M{M} N{N} K{K} mode{mode} b_mn{int(b_mn)} colsum{int(use_cs)}: {ms:.4f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
