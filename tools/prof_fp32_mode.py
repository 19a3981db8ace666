"""ncu target / timing of the fp32-accurate mode: SwinTransformerLayerv5 forward + backward, 2 clips, precision='fp32'."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import swin
dev = "cuda"
m = swin.SwinTransformerLayerv5(dim=512, input_resolution=(64, 80), num_heads=4).to(dev)
m.precision = "fp32"
x = torch.relu(torch.randn(2, 4, 512, 64, 80, device=dev))
wa, wb = torch.randn_like(x) * 0.1, torch.randn(2, 4, 1024, 32, 40, device=dev) * 0.1
def step():
    m.zero_grad(set_to_none=True)
    a, b = m(x)
    ((a * wa).sum() + (b * wb).sum()).backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
print("fp32 mode step: %.2f ms" % (e0.elapsed_time(e1) / 3))
