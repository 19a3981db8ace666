"""Component timings of BASELINE.md section 3 (items 2-5): GPU kernels (CUDA events) next to the CPU
oracle port of the reference on this box's host cores.  Prints one JSON document.

  python tools/bench_components.py [--no-cpu]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from stswincl_b200 import contrast, swin      # noqa: E402

PEAKS = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
dev = torch.device("cuda", 0)
do_cpu = "--no-cpu" not in sys.argv


def gpu_ms(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_ms(fn, reps=1):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


out = {"device": torch.cuda.get_device_name(0), "cpu_threads": os.cpu_count(), "torch": torch.__version__, "items": {}}
if do_cpu:
    from oracle import loss_oracle as lo, swin_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)

# ---- item 3: one stage-1 and one stage-2 block, forward + backward, shifted, B = 1 clip pair
for name, dim, res, ws in (("block_stage1", 512, (64, 80), 8), ("block_stage2", 1024, (32, 40), 4)):
    heads, shift = 4, ws // 2
    m = swin.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift).to(dev)
    x = torch.relu(torch.randn(1, 2, res[0] * res[1], dim, device=dev)).to(torch.bfloat16).requires_grad_(True)
    g = torch.randn_like(x)

    def f():
        x.grad = None
        m.zero_grad(set_to_none=True)
        y = m(x)
        y.backward(g)
    ms = gpu_ms(f)
    tok = 2 * res[0] * res[1]
    flops = 3 * (24 * dim * dim + 4 * 2 * ws * ws * dim) * tok          # fwd + 2x bwd
    item = {"gpu_ms": ms, "gpu_tflops": flops / ms / 1e9, "algorithmic_gflop_fwd_bwd": flops / 1e9}
    if do_cpu:
        p = so.make_block_params(dim, res, heads, ws, shift, seed=1)
        leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != "attn_mask" else v) for k, v in p.items()}
        xc = so.make_features(2, 1, 2, res[0] * res[1], dim).requires_grad_(True)

        def fc():
            y = so.swin_block(xc, leaf, res, heads, ws, shift)
            y.sum().backward()
        item["cpu_ms"] = cpu_ms(fc)
    out["items"][name] = item

# ---- item 4: WindowAttention (qkv + core + proj) at [80, 2, 64, 512] with the shift mask, forward only and fwd+bwd
for name, dim, res, ws in (("window_attention_stage1", 512, (64, 80), 8), ("window_attention_stage2", 1024, (32, 40), 4)):
    heads, shift = 4, ws // 2
    a = swin.WindowAttention(dim, (ws, ws), heads).to(dev)
    x = torch.relu(torch.randn(1, 2, res[0] * res[1], dim, device=dev)).to(torch.bfloat16).requires_grad_(True)
    g = torch.randn_like(x)
    geom = (res[0], res[1], heads, ws, shift, 0.0)

    def fwd():
        return swin._AttentionFn.apply(x, a.relative_position_bias_table, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, geom)

    def fb():
        x.grad = None
        a.zero_grad(set_to_none=True)
        fwd().backward(g)
    with torch.no_grad():
        ms_f = gpu_ms(fwd)
    ms_fb = gpu_ms(fb)
    tok = 2 * res[0] * res[1]
    f_fwd = (8 * dim * dim + 4 * 2 * ws * ws * dim) * tok
    out["items"][name] = {"gpu_fwd_ms": ms_f, "gpu_fwd_tflops": f_fwd / ms_f / 1e9, "gpu_fwd_bwd_ms": ms_fb,
                          "gpu_fwd_bwd_tflops": 3 * f_fwd / ms_fb / 1e9, "algorithmic_gflop_fwd": f_fwd / 1e9}

# ---- item 2: the whole layer, forward + backward, B = 1 clip
m = swin.SwinTransformerLayerv5().to(dev)
x = torch.relu(torch.randn(1, 4, 512, 64, 80, device=dev)).to(torch.bfloat16).requires_grad_(True)
g1, g2 = torch.randn(1, 4, 512, 64, 80, device=dev).to(torch.bfloat16), torch.randn(1, 4, 1024, 32, 40, device=dev).to(torch.bfloat16)


def layer():
    x.grad = None
    m.zero_grad(set_to_none=True)
    y1, y2 = m(x)
    torch.autograd.backward([y1, y2], [g1, g2])
ms = gpu_ms(layer)
item = {"gpu_ms": ms, "gpu_tflops": 4.0e12 / ms / 1e9, "algorithmic_tflop_fwd_bwd": 4.0, "frames_per_s": 4 / (ms * 1e-3)}
if do_cpu:
    p = so.make_layer_params(512, (64, 80), 4, seed=0)
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("attn_mask") else v) for k, v in p.items()}
    xc = so.make_features(1, 1, 4, 512, 64, 80).requires_grad_(True)

    def lc():
        y1, y2 = so.swin_layer_v5(xc, leaf, 512, (64, 80), 4)
        (y1.sum() + y2.sum()).backward()
    item["cpu_ms"] = cpu_ms(lc)
out["items"]["layer_1clip"] = item

# ---- item 5: regression_loss forward + backward, C = 256, 32x56, 12 classes
for N in (2, 4):
    K, C, H, W = 12, 256, 32, 56
    gen = torch.Generator(device=dev).manual_seed(N)
    labels = [torch.randint(0, K, (N, 1, 4, 7), generator=gen, device=dev).float().repeat_interleave(8, 2).repeat_interleave(8, 3) for _ in range(6)]
    emb = [torch.nn.functional.normalize(torch.randn(N, C, H, W, generator=gen, device=dev), dim=1) for _ in range(6)]
    q = emb[0].clone().requires_grad_(True)

    def lossfb():
        q.grad = None
        contrast.pixel_contrast_loss(q, emb[1:], labels[0], labels[1:], K, validate_labels=False).backward()
    ms = gpu_ms(lossfb)
    flops = 2 * 5 * 2 * (H * W) ** 2 * C * N                               # dense form, fwd + bwd(dq)
    item = {"gpu_ms": ms, "gpu_tflops": flops / ms / 1e9, "algorithmic_gflop_fwd_bwd": flops / 1e9}
    if do_cpu:
        lab = lo.make_label_maps(1, 6, N, H, W, K)
        em = lo.make_embeddings(2, lab, C, K)
        qc = em[0].clone().requires_grad_(True)

        def lc2():
            qc.grad = None
            lo.regression_loss(qc, em[1:], lab[0], lab[1:], K).backward()
        item["cpu_ms"] = cpu_ms(lc2)
    out["items"][f"regression_loss_N{N}"] = item

print(json.dumps(out, indent=1))
