"""Small-shape run of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
the fused pixel-loss step (bf16 and fp32 mode, fused normalize, mixed-label groups), one fp32-mode block forward +
backward, FusedAdam, colsum, gather_cast."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import loss_oracle as lo, swin_oracle as so
from stswincl_b200 import contrast, ops, optim, swin

dev = "cuda"
N, C, H, W, K = 2, 64, 8, 14, 12
for coarse in ((4, 7), (8, 14)):
    full = lo.make_label_maps(1, 6, N, 8 * H, 8 * W, K, coarse=coarse)
    ds = [torch.nn.functional.interpolate(m, size=[H, W], mode="nearest") for m in full]
    raw = lo.make_embeddings(2, ds + ds[:2], C, K)
    for prec in ("bf16", "fp32"):
        p1, p2 = raw[6].to(dev).requires_grad_(True), raw[7].to(dev).requires_grad_(True)
        loss = contrast.consistency_loss_tail(p1, p2, *[r.to(dev) for r in raw[:6]], *[m.to(dev) for m in full], K, normalize=True,
                                              precision=prec)
        loss.backward()
        print("tail", coarse, prec, float(loss), float(p1.grad.abs().sum()))
dim, res, heads, ws, shift = 128, (16, 24), 2, 8, 4
params = so.make_block_params(dim, res, heads, ws, shift, seed=3)
blk = swin.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift)
blk.load_state_dict(params, strict=True)
blk = blk.to(dev)
x = so.make_features(4, 1, 2, res[0] * res[1], dim).to(dev).requires_grad_(True)
for prec in ("bf16", "fp32"):
    blk.precision = prec
    blk.zero_grad(); x.grad = None
    y = blk(x)
    y.sum().backward()
    print("block", prec, float(y.abs().sum()), float(x.grad.abs().sum()))
opt = optim.FusedAdam(blk.parameters(), lr=1e-3)
opt.step()
out = torch.zeros(dim, device=dev)
ops.colsum(torch.randn(100, dim, device=dev).to(torch.bfloat16), out)
src = [torch.randn(n, device=dev) for n in (5, 4096, 1003)]
dst = [torch.empty(n, device=dev, dtype=torch.bfloat16) for n in (5, 4096, 1003)]
ops.gather_cast(dst, src)
torch.cuda.synchronize()
print("done")
