"""BASELINE config 5: shifted-window attention roofline sweep over window size and head count on a CaDIS-shaped
OS-8 feature map (540x960 -> 64x120 after padding to the window grid).  Reports forward / backward time and the
achieved fraction of the HBM roofline (8*C / 14*C bytes per token) for every (window, heads, frames) point the
kernels accept.  Output: one JSON object."""
import json, sys
import torch
sys.path.insert(0, ".")
from stswincl_b200 import ops
PEAK = 6542.1   # GB/s, MEASURED_PEAKS.json
dev = "cuda"
out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": PEAK, "points": []}
C, B = 512, 8
for (H, W, ws) in ((64, 120, 8), (64, 120, 4), (56, 84, 7), (60, 120, 5), (64, 120, 2)):
    for T in (1, 2):
        if T * ws * ws > 128:
            continue
        for nH in (4, 8, 16):
            hd = C // nH
            if not (hd == 32 or hd % 64 == 0):
                continue
            for shift in (0, ws // 2):
                g = torch.Generator(device=dev).manual_seed(0)
                qkv = torch.randn(B, T, H * W, 3 * C, generator=g, device=dev).to(torch.bfloat16)
                table = torch.randn((2 * ws - 1) ** 2, nH, generator=g, device=dev) * 0.5
                do = torch.randn(B, T, H * W, C, generator=g, device=dev).to(torch.bfloat16)
                try:
                    o, lse = ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)
                except Exception as e:      # outside the supported geometry: reported, not hidden
                    out["points"].append({"H": H, "W": W, "ws": ws, "T": T, "heads": nH, "shift": shift, "error": str(e)[:80]})
                    continue
                dt = torch.zeros_like(table)
                res = {}
                for name, bpt, fn in (("fwd", 8 * C, lambda: ops.winattn_fwd(qkv, table, H, W, nH, ws, shift)),
                                      ("bwd", 14 * C, lambda: ops.winattn_bwd(qkv, table, lse, do, H, W, nH, ws, shift, dt))):
                    for _ in range(2): fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5): fn()
                    e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 5
                    gbs = bpt * B * T * H * W / ms / 1e6
                    res[name + "_ms"] = round(ms, 4); res[name + "_gbs"] = round(gbs, 1); res[name + "_frac"] = round(gbs / PEAK, 3)
                flops = 4 * (T * ws * ws) * C * B * T * H * W
                res["fwd_tflops"] = round(flops / res["fwd_ms"] / 1e9, 1)
                out["points"].append({"H": H, "W": W, "ws": ws, "T": T, "heads": nH, "head_dim": hd, "shift": shift, **res})
print(json.dumps(out, indent=1))
