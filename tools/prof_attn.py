"""ncu target: window-attention fwd+bwd at the stage-1 / stage-2 bench geometry (B2 = 16 pair-batches)."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Bp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
if stage == 1:
    H, W, C, nH, ws, shift = 64, 80, 512, 4, 8, 4
else:
    H, W, C, nH, ws, shift = 32, 40, 1024, 4, 4, 2
T = 2
import os
if os.environ.get("PROF_GEOM"):        # "H,W,C,nH,ws,shift,T,B": any other geometry (e.g. 64,120,512,8,8,0,1,8)
    H, W, C, nH, ws, shift, T, Bp = (int(v) for v in os.environ["PROF_GEOM"].split(","))
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(Bp, T, H * W, 3 * C, generator=g, device="cuda").to(torch.bfloat16)
table = torch.randn((2 * ws - 1) ** 2, nH, generator=g, device="cuda") * 0.5
do = torch.randn(Bp, T, H * W, C, generator=g, device="cuda").to(torch.bfloat16)
for _ in range(3):
    out, lse = ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)
    dt = torch.zeros_like(table)
    dq = ops.winattn_bwd(qkv, table, lse, do, H, W, nH, ws, shift, dt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda: ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)),
                 ("bwd", lambda: ops.winattn_bwd(qkv, table, lse, do, H, W, nH, ws, shift, dt))):
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    toks = Bp * T * H * W
    byt = (8 if name == "fwd" else 14) * C * toks
    print(name, "stage", stage, "ms", round(ms, 4), "GB/s", round(byt / ms / 1e6, 1))
