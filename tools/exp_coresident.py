"""Does a small co-running kernel stall the persistent kernels?  20 GEMM launches (163840 x 2048 x 512, bias mode) on the
current stream, with and without a spin kernel (G CTAs, 512 threads, 64 KB of shared memory each, `ms` milliseconds) on a
side stream.  A persistent kernel with one CTA per SM and a static item -> CTA map cannot finish before every one of its
CTAs has been resident."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
spin = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "libspin.so"))
spin.spin_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
M, N, K = 163840, 2048, 512
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
side = torch.cuda.Stream()
def run(ctas, ms):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if ctas:
        spin.spin_launch(ctas, 512, 64 * 1024, int(ms * 1e6), side.cuda_stream)
    e0.record()
    for _ in range(20):
        ops.gemm(A, B, mode=0, bias=bias, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(2): run(0, 0)
print("20 GEMMs alone: %.3f ms" % run(0, 0))
for ctas, ms in ((2, 3.0), (8, 3.0), (2, 1.0)):
    print("20 GEMMs next to %d spinning CTAs for %.1f ms: %.3f ms" % (ctas, ms, run(ctas, ms)))
