"""Per-kernel-name GPU time of one training step of the Swin head (torch.profiler, CUPTI): shows
what the step spends outside our own kernels (optimizer, casts, adds, copies)."""
import sys, json, collections
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from stswincl_b200 import swin
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = swin.SwinTransformerLayerv5().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=3e-5, fused=True)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.relu(torch.randn(B, 4, 512, 64, 80, generator=g, device=dev)).to(torch.bfloat16)
g1 = (torch.randn(B, 4, 512, 64, 80, generator=g, device=dev) * 0.1).to(torch.bfloat16)
g2 = (torch.randn(B, 4, 1024, 32, 40, generator=g, device=dev) * 0.1).to(torch.bfloat16)
def step():
    opt.zero_grad(set_to_none=True)
    y1, y2 = model(x)
    torch.autograd.backward([y1, y2], [g1, g2])
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name[:90]
        tot[k][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot[k][1] += 1
rows = sorted(tot.items(), key=lambda kv: -kv[1][0])
total = sum(v[0] for v in tot.values())
print("total_us", total)
for k, (t, n) in rows[:40]:
    print(f"{t:10.1f} us {n:5d}  {100*t/total:5.1f}%  {k}")
