"""Timeline of CTA 0 of the window-attention forward kernel (debug build, tools/build_trace.sh).
Prints, per work item, the cycle offsets of the named points of each warp role."""
import os, sys, ctypes
import torch
sys.path.insert(0, ".")
from stswincl_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_trace", "libstswin_trace.so")
_lib._stale = lambda: False
from stswincl_b200 import ops
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
Bp = 16
if stage == 1:
    H, W, C, nH, ws, shift = 64, 80, 512, 4, 8, 4
else:
    H, W, C, nH, ws, shift = 32, 40, 1024, 4, 4, 2
T = 2
if os.environ.get("TRACE_GEOM"):       # "H,W,C,nH,ws,shift,T,B": any other geometry (e.g. 64,120,512,16,8,0,1,8)
    H, W, C, nH, ws, shift, T, Bp = (int(v) for v in os.environ["TRACE_GEOM"].split(","))
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(Bp, T, H * W, 3 * C, generator=g, device="cuda").to(torch.bfloat16)
table = torch.randn((2 * ws - 1) ** 2, nH, generator=g, device="cuda") * 0.5
do = torch.randn(Bp, T, H * W, C, generator=g, device="cuda").to(torch.bfloat16)
lib = _lib.load()
NPT = 16
buf = torch.zeros(64 * NPT, dtype=torch.int64, device="cuda")
out, lse = ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)
dt = torch.zeros_like(table)
if which == "bwd":
    ops.winattn_bwd(qkv, table, lse, do, H, W, nH, ws, shift, dt)
torch.cuda.synchronize()
fn = getattr(lib, "stswin_debug_trace_" + which)
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
if which == "fwd":
    ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)
    names = ["ld_first", "ld_last", "S_issued", "PV_begin", "PV_issued", "item_start", "geom_done", "S_ready",
             "pass1_done", "P_free", "pass2_done", "O_ready", "epi_done"]
else:
    ops.winattn_bwd(qkv, table, lse, do, H, W, nH, ws, shift, dt)
    names = ["ldH_first", "ldH_last", "ldL_first", "ldL_last", "SdP_issued", "out_begin", "out_issued", "item_start",
             "SdP_ready", "pass1_done", "pass2_done", "drain_beg", "drain_done", "P_free", "dS_free"]
torch.cuda.synchronize()
assert fn(0) == 0
t = buf.view(64, NPT).cpu()
base = int(t[4][t[4] > 0].min())
print("cycles relative to the earliest event of item 4;", which, "stage", stage)
print("item " + " ".join(f"{n:>10s}" for n in names))
for it in range(4, 16):
    print(f"{it:4d} " + " ".join(f"{int(t[it, k]) - base:10d}" if t[it, k] > 0 else f"{'-':>10s}" for k in range(len(names))))
ends = t[4:30, 12 if which == 'fwd' else len(names) - 1]
print("avg cycles per item (last point, items 4..29):", float((ends[-1] - ends[0]) / (len(ends) - 1)))
