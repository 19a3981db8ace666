"""Per-quantity parity of SwinTransformerBlock forward + backward against the CPU oracle at the shipped geometries
(the numbers behind tests/test_gpu_swin.py::test_real_geometry_block_forward_backward_vs_oracle)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import rel_err
from oracle import swin_oracle as so
from stswincl_b200 import swin
import os
if os.environ.get("GELU_GRAD_Q8") is not None:      # A/B of the one-byte GELU' storage
    swin.GELU_GRAD_Q8 = os.environ["GELU_GRAD_Q8"] != "0"
print("GELU_GRAD_Q8 =", swin.GELU_GRAD_Q8)
CASES = [("S1_unshifted", 512, (64, 80), 4, 8, 0), ("S1_shifted", 512, (64, 80), 4, 8, 4),
         ("S2_unshifted", 1024, (32, 40), 4, 4, 0), ("S2_shifted", 1024, (32, 40), 4, 4, 2)]
for tag, dim, res, heads, ws, shift in CASES:
    L = res[0] * res[1]
    params = so.make_block_params(dim, res, heads, ws, shift, seed=91)
    m = swin.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift)
    m.load_state_dict(params, strict=True); m = m.cuda()
    x = so.make_features(92, 2, 2, L, dim)
    w = (so.make_features(93, 2, 2, L, dim) - 0.4).to(torch.bfloat16).float()
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != "attn_mask" else v) for k, v in params.items()}
    xr = x.to(torch.bfloat16).float().requires_grad_(True)
    ref = so.swin_block(xr, leaf, res, heads, ws, shift)
    (ref * w).sum().backward()
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    (y * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    out = {"y": rel_err(y.cpu(), ref), "dx": rel_err(xg.grad.cpu(), xr.grad)}
    for n, p in m.named_parameters():
        out[n] = rel_err(p.grad.cpu(), leaf[n].grad)
    print(tag, {k: round(v, 4) for k, v in out.items()}, flush=True)
