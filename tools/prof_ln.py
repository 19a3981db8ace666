"""LayerNorm forward / backward at the bench row counts: achieved GB/s against the algorithmic bytes."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
for M, C in ((163840, 512), (81920, 512), (40960, 1024), (20480, 1024)):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, C, generator=g, device="cuda").to(torch.bfloat16)
    dy = torch.randn(M, C, generator=g, device="cuda").to(torch.bfloat16)
    dr = torch.randn(M, C, generator=g, device="cuda").to(torch.bfloat16)
    gamma = torch.randn(C, device="cuda"); beta = torch.randn(C, device="cuda")
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda"); cs = torch.zeros(C, device="cuda")
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    fns = (("fwd", 4 * C, lambda: ops.layernorm_fwd(x, gamma, beta)),
           ("bwd+res+colsum", 8 * C, lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db, dres=dr, dx_colsum=cs)),
           ("bwd", 6 * C, lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db)))
    for name, bpr, fn in fns:
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"M={M} C={C} {name:15s} {ms:.4f} ms  {bpr * M / ms / 1e6:.0f} GB/s")
