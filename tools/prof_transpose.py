"""Timing of the bf16 layout transpose at the bench shapes ([B*T, C, H*W] <-> [B*T, H*W, C])."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
for (b, R, Cc) in ((32, 512, 5120), (32, 5120, 512), (32, 1024, 1280), (32, 1280, 1024)):
    x = torch.randn(b, R, Cc, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        y = ops.transpose(x, torch.bfloat16)
    assert torch.equal(y, x.transpose(1, 2).contiguous())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.transpose(x, torch.bfloat16)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"transpose bf16 [{b},{R},{Cc}]: {ms:.4f} ms  {4 * x.numel() / ms / 1e6:.0f} GB/s")
