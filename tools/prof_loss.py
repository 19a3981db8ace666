"""Pixel contrastive loss (HP-2) at pre-training shapes: C 256, 32x56, 12 classes, 5 key sets; per-family device time."""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import contrast, ops
dev = torch.device("cuda", 0)
K, C, H, W = 12, 256, 32, 56
for N in (4, 16, 32):
    gen = torch.Generator(device=dev).manual_seed(N)
    labels = [torch.randint(0, K, (N, 1, 4, 7), generator=gen, device=dev).float().repeat_interleave(8, 2).repeat_interleave(8, 3) for _ in range(6)]
    emb = [torch.nn.functional.normalize(torch.randn(N, C, H, W, generator=gen, device=dev), dim=1) for _ in range(6)]
    q = emb[0].clone().requires_grad_(True)
    def step():
        q.grad = None
        contrast.pixel_contrast_loss(q, emb[1:], labels[0], labels[1:], K, validate_labels=False).backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = 2 * 5 * 2 * (H * W) ** 2 * C * N
    prof = ops.EventProfiler()
    ops.set_profiler(prof)
    step()
    ops.set_profiler(None)
    fam = prof.summary()
    print(f"N={N}: {ms:.3f} ms fwd+bwd, {flops / ms / 1e9:.1f} TFLOP/s (dense form)", {k: (round(v['ms'], 4), round(v['work'] / v['ms'] / 1e9, 1)) for k, v in fam.items()})
