"""Device timings of the training-step kernels (OHEM cross-entropy, key-encoder EMA, LARS-scaled SGD) at the
reference's sizes, as achieved HBM GB/s against MEASURED_PEAKS.json, next to the reference's eager op sequence
on the same GPU (torch ops: the sort of losses.py:34, the per-tensor loops of lars.py / PixPro_swin_v5.py:266).
    python tools/bench_trainaux.py > gpurun_out/trainaux.json
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stswincl_b200 import losses, optim, swin  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = torch.device("cuda", 0)


def timeit(fn, n=20, warm=3, graph=True):
    """Average device time of fn.  Our paths are replayed from a CUDA graph (the host side of one call -- ctypes
    pointer tables, allocations -- would otherwise bound these sub-millisecond launches); the eager torch
    sequences have host reads and are timed as they run in the reference."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            fn = gr.replay
            fn()
            torch.cuda.synchronize()
        except Exception as e:
            print(f"graph capture failed: {type(e).__name__}: {e}", file=sys.stderr)
            torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def row(name, ms, nbytes, ref_ms=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    r = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 3)}
    if ref_ms is not None:
        r["torch_eager_same_gpu_ms"] = round(ref_ms, 3)
    return name, r


out = {}
# ---- OHEM: train_swin.py:123 shape, batch 8, 12 classes, 512x640, n_min = H*W/16 per the reference (not scaled by batch)
B, K, H, W = 8, 12, 512, 640
g = torch.Generator(device=dev).manual_seed(0)
labels = torch.randint(0, K, (B, H, W), generator=g, device=dev)
for tag, margin in (("threshold_branch", 0.0), ("topk_branch", 12.0)):
    logits = torch.randn(B, K, H, W, generator=g, device=dev) * 1.5
    logits += margin * torch.nn.functional.one_hot(labels, K).permute(0, 3, 1, 2).float()
    logits.requires_grad_(True)
    crit = losses.OhemCELoss2D(H * W // 16)

    def fwd():
        return crit(logits, labels)

    def fwdbwd():
        logits.grad = None
        crit(logits, labels).backward()

    def ref_fwdbwd():
        logits.grad = None
        l = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-1, reduction="none").view(-1)
        l, _ = torch.sort(l, descending=True)
        l = l[l > crit.thresh] if l[crit.n_min] > crit.thresh else l[:crit.n_min]
        l.mean().backward()

    px = B * H * W
    t_f, t_fb, t_ref = timeit(fwd), timeit(fwdbwd), timeit(ref_fwdbwd, n=5, warm=2, graph=False)
    k, v = row(f"ohem_fwd_{tag}", t_f, px * (K * 4 + 8 + 4))
    out[k] = v
    k, v = row(f"ohem_fwd_bwd_{tag}", t_fb, px * (K * 4 + 8 + 4) + px * (K * 4 + 8 + 4) + (px * K * 4 if margin == 0.0 else 0), t_ref)
    out[k] = v
    del logits

# ---- EMA + LARS over the Swin-head parameter set (96.6 M fp32 parameters, 150 tensors)
q = swin.SwinTransformerLayerv5(dim=512, input_resolution=(64, 80), num_heads=4).to(dev)
kk = swin.SwinTransformerLayerv5(dim=512, input_resolution=(64, 80), num_heads=4).to(dev)
qp, kp = list(q.parameters()), list(kk.parameters())
n_bytes = sum(p.numel() for p in qp) * 4


def ref_ema():
    with torch.no_grad():
        for a, b in zip(qp, kp):
            b.data = b.data * 0.99 + a.data * (1. - 0.99)


k, v = row("ema_update", timeit(lambda: optim.momentum_update(qp, kp, 0.99)), 3 * n_bytes, timeit(ref_ema, n=5, warm=2, graph=False))
v["tensors"] = len(qp)
out[k] = v

opt = optim.LARS(torch.optim.SGD(optim.add_weight_decay(q, 1e-5), lr=0.1, momentum=0.9))
grads = [torch.randn_like(p) * 0.01 for p in qp]


def lars_step():
    for p, gr in zip(qp, grads):
        p.grad = gr
    opt.step()


def ref_lars_step():          # lars.py:109-152 with torch ops (one .norm() host read per weight tensor)
    with torch.no_grad():
        for group in opt.param_groups:
            wd, ignore = group['weight_decay'], group.get('ignore')
            for p in group['params']:
                gr = p.grad
                if wd > 0:
                    gr = gr.add(p, alpha=wd)
                if ignore is not None and not ignore:
                    pn, gn = p.norm(), gr.norm()
                    a = 1.0
                    if pn > 0 and gn > 0:
                        a = 0.001 * pn / (gn + 1e-8)
                    gr = gr.mul(a)
                buf = opt.state[p]['momentum_buffer']
                buf.mul_(0.9).add_(gr)
                p.add_(buf, alpha=-0.1)


lars_step()
t = timeit(lars_step)
for p, gr in zip(qp, grads):
    p.grad = gr
t_ref = timeit(ref_lars_step, n=5, warm=2, graph=False)
decay_bytes = sum(p.numel() for p in qp if p.dim() > 1) * 4
k, v = row("lars_sgd_step", t, 6 * n_bytes + 2 * decay_bytes, t_ref)     # update: p, g, buf in and out; norms: p, g of the LARS group
out[k] = v
print(json.dumps({"device": torch.cuda.get_device_name(0), "hbm_peak_GBps": PEAK, "items": out}, indent=1))
