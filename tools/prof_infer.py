"""Forward-only (no_grad) throughput of the Swin head: the key-encoder side of the pre-training model."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import swin
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = swin.SwinTransformerLayerv5().to(dev)
x = torch.relu(torch.randn(B, 4, 512, 64, 80, device=dev)).to(torch.bfloat16)
for grad in (True, False):
    ctxm = torch.enable_grad() if grad else torch.no_grad()
    with ctxm:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3): model(x)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):          # launch overhead of ~750 kernels would hide the device time
            out = model(x)
        for _ in range(3): graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): graph.replay()
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"forward only (CUDA graph), grad={grad}: {ms:.2f} ms per {B} clips = {B * 4 / ms * 1e3:.0f} frames/s")
    del graph, out
