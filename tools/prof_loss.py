"""Pixel contrastive loss (HP-2) at pre-training shapes: C 256, 32x56, 12 classes; the symmetric two-call step of
ConsistencyLoss.forward (PixPro_swin_v5.py:584-597: 2 queries x 5 key sets, 6 full-resolution label maps), forward +
backward, replayed from a CUDA graph; per-family device time from CUDA events on an eager step.
    python tools/prof_loss.py [N ...]         (default 4 16 32)"""
import sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import contrast, ops
dev = torch.device("cuda", 0)
K, C, H, W = 12, 256, 32, 56
Ns = [int(a) for a in sys.argv[1:]] or [4, 16, 32]
for N in Ns:
    for normalize in (False, True):
        gen = torch.Generator(device=dev).manual_seed(N)
        labels = [torch.randint(0, K, (N, 1, 8, 14), generator=gen, device=dev).float().repeat_interleave(32, 2).repeat_interleave(32, 3) for _ in range(6)]
        emb = [torch.randn(N, C, H, W, generator=gen, device=dev) for _ in range(8)]
        if not normalize:
            emb = [torch.nn.functional.normalize(e, dim=1) for e in emb]
        q1, q2 = emb[6].requires_grad_(True), emb[7].requires_grad_(True)
        def step():
            q1.grad = None; q2.grad = None
            contrast.consistency_loss_tail(q1, q2, *emb[:6], *labels, K, normalize=normalize).backward()
        for _ in range(3): step()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        flops = 2 * 2 * 5 * 2 * (H * W) ** 2 * C * N
        prof = ops.EventProfiler()
        ops.set_profiler(prof)
        step()
        ops.set_profiler(None)
        fam = prof.summary()
        print(f"N={N} normalize={normalize}: {ms:.4f} ms per symmetric step (fwd+bwd, graph replay), {flops / ms / 1e9:.1f} TFLOP/s (dense form)",
              {k: (round(v['ms'], 4), round(v['work'] / v['ms'] / 1e9, 1)) for k, v in fam.items()}, flush=True)
