"""ncu target: a few launches of one GEMM shape.  usage: prof_gemm.py M N K [mode]"""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
M, N, K = (int(a) for a in sys.argv[1:4])
g = torch.Generator().manual_seed(0)
A = (torch.randn(M, K, generator=g) * K ** -0.5).to(torch.bfloat16).cuda()
B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(4):
    ops.gemm(A, B, out=out)
torch.cuda.synchronize()
