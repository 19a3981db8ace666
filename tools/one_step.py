"""ncu target: warm-up step + one measured training step of the Swin head (same workload as bench.py)."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import swin
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = swin.SwinTransformerLayerv5().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=3e-5, fused=True)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.relu(torch.randn(B, 4, 512, 64, 80, generator=g, device=dev)).to(torch.bfloat16)
g1 = (torch.randn(B, 4, 512, 64, 80, generator=g, device=dev) * 0.1).to(torch.bfloat16)
g2 = (torch.randn(B, 4, 1024, 32, 40, generator=g, device=dev) * 0.1).to(torch.bfloat16)
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    y1, y2 = model(x)
    loss = (y1.float() * g1).sum() + (y2.float() * g2).sum()
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
