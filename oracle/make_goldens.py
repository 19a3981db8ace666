"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE.  Run in the authoring container only (needs
/root/reference):   python -m oracle.make_goldens

Every fixture stores the reference's outputs (and autograd gradients) for inputs
and weights that are regenerated from seeds by ``oracle.swin_oracle.make_*`` /
``oracle.loss_oracle.make_*`` -- plus a float64 checksum of those inputs so that
RNG drift between machines is detected instead of silently compared.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import index_oracle as ix
from . import loss_oracle as lo
from . import ref_shims
from . import swin_oracle as so

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

INDEX_CASES = [  # (H, W, ws, shift)
    (64, 80, 8, 4), (32, 40, 4, 2), (32, 56, 8, 4), (16, 28, 4, 2), (64, 120, 8, 4),
    (16, 24, 8, 4), (8, 12, 4, 2), (56, 84, 7, 3), (16, 24, 8, 0), (8, 8, 8, 4),
]


def checksum(*tensors) -> float:
    return float(sum(t.double().abs().sum().item() for t in tensors))


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def gen_index(sw) -> None:
    out = {}
    for (H, W, ws, shift) in INDEX_CASES:
        tag = f"{H}x{W}_ws{ws}_s{shift}"
        blk = sw.SwinTransformerBlock(8, (H, W), 2, window_size=ws, shift_size=shift)
        ews, eshift = blk.window_size, blk.shift_size
        out[f"relidx_{tag}"] = _np(blk.attn.relative_position_index).astype(np.int16)
        if blk.attn_mask is not None:
            out[f"mask_{tag}"] = _np(blk.attn_mask).astype(np.int8)      # values {0,-100}
        # gather map: run roll + window_partition on a token-id image
        ids = torch.arange(H * W, dtype=torch.float32).view(1, H, W, 1)
        shifted = torch.roll(ids, shifts=(-eshift, -eshift), dims=(1, 2)) if eshift > 0 else ids
        win = sw.window_partition(shifted, ews).view(-1, ews * ews)
        out[f"gather_{tag}"] = _np(win).astype(np.int32)
        # scatter map: window_reverse + roll back must be the inverse permutation (T=2)
        nW = win.shape[0]
        slot = torch.arange(nW * 2 * ews * ews, dtype=torch.float32).view(nW, 2, ews, ews, 1)
        rev = sw.window_reverse(slot, ews, H, W, 2).view(2, H, W, 1)
        back = torch.roll(rev, shifts=(eshift, eshift), dims=(1, 2)) if eshift > 0 else rev
        out[f"scatter_{tag}"] = _np(back.view(2, H * W)).astype(np.int32)
        out[f"eff_{tag}"] = np.array([ews, eshift], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "index_cases.npz"), **out)


def _load(module, params) -> None:
    missing, unexpected = module.load_state_dict(params, strict=True)
    assert not missing and not unexpected


def _grads(module, names):
    sd = dict(module.named_parameters())
    return {n: _np(sd[n].grad) for n in names}


ATTN_CASES = [  # (tag, dim, ws, heads, T, (H, W), shift, B)
    ("a0", 128, 4, 2, 2, (8, 12), 2, 2),
    ("a1", 128, 4, 2, 1, (8, 12), 0, 2),
    ("a2", 256, 8, 2, 2, (16, 24), 4, 1),
]


def gen_attention(sw) -> None:
    out = {}
    for (tag, dim, ws, heads, T, (H, W), shift, B) in ATTN_CASES:
        nW = (H // ws) * (W // ws)
        params = so.make_attention_params(dim, ws, heads, seed=11)
        m = sw.WindowAttention(dim, (ws, ws), heads)
        _load(m, params)
        mask_np = ix.shift_attn_mask(H, W, ws, shift)
        mask = torch.from_numpy(mask_np) if mask_np is not None else None
        x = (so.make_features(21, B * nW, T, ws * ws, dim) - 0.4).requires_grad_(True)
        w = so.make_features(22, B * nW, T, ws * ws, dim) - 0.4          # upstream gradient
        y = m(x, mask=mask)
        (y * w).sum().backward()
        out[f"{tag}_y"] = _np(y)
        out[f"{tag}_dx"] = _np(x.grad)
        for n, g in _grads(m, ["relative_position_bias_table", "qkv.bias", "proj.bias"]).items():
            out[f"{tag}_d_{n}"] = g
        g = _grads(m, ["qkv.weight", "proj.weight"])
        out[f"{tag}_d_qkv.weight_rows"] = g["qkv.weight"][::17]          # row subsample
        out[f"{tag}_d_proj.weight_rows"] = g["proj.weight"][::17]
        out[f"{tag}_insum"] = np.array(checksum(x, w, *[v for v in params.values() if v.is_floating_point()]))
    np.savez_compressed(os.path.join(OUT, "swin_attention.npz"), **out)


BLOCK_CASES = [  # (tag, dim, (H, W), heads, ws, shift, T, B)
    ("b0", 128, (16, 24), 2, 8, 0, 2, 1),
    ("b1", 128, (16, 24), 2, 8, 4, 2, 1),
    ("b2", 256, (8, 12), 4, 4, 2, 2, 2),
    ("b3", 128, (16, 24), 2, 8, 4, 1, 1),     # T=1 through the assert-free subclass
]


def gen_block(sw) -> None:
    class AnyT(sw.SwinTransformerBlock):
        """forward of swin_512.py:196-237 re-dispatched with the T==2 assert defeated:
        the body is T-generic, so fold T into pairs of 'virtual' frames is NOT done --
        we instead call the pieces in the reference's own order."""

        def forward(self, x_v):
            H, W = self.input_resolution
            B, T, L, C = x_v.shape
            ws, s = self.window_size, self.shift_size
            shortcut = x_v.reshape(B * T, L, C)
            x = x_v.reshape(B * T, H, W, C)
            if s > 0:
                x = torch.roll(x, shifts=(-s, -s), dims=(1, 2))
            xw = sw.window_partition(x, ws).view(B, T, -1, ws * ws, C).permute(0, 2, 1, 3, 4).contiguous().view(-1, T, ws * ws, C)
            aw = self.attn(xw, mask=self.attn_mask)
            x = sw.window_reverse(aw, ws, H, W, T).view(B * T, H, W, C)
            if s > 0:
                x = torch.roll(x, shifts=(s, s), dims=(1, 2))
            x = shortcut + x.view(B * T, L, C)
            x = self.norm1(x + self.mlp(self.norm2(x)))
            return x.view(B, T, L, C)

    out = {}
    for (tag, dim, (H, W), heads, ws, shift, T, B) in BLOCK_CASES:
        params = so.make_block_params(dim, (H, W), heads, ws, shift, seed=31)
        cls = sw.SwinTransformerBlock if T == 2 else AnyT
        m = cls(dim, (H, W), heads, window_size=ws, shift_size=shift)
        _load(m, params)
        x = so.make_features(41, B, T, H * W, dim).requires_grad_(True)
        w = so.make_features(42, B, T, H * W, dim) - 0.4
        y = m(x)
        (y * w).sum().backward()
        out[f"{tag}_y"] = _np(y)
        out[f"{tag}_dx"] = _np(x.grad)
        small = ["attn.relative_position_bias_table", "attn.qkv.bias", "attn.proj.bias", "norm1.weight",
                 "norm1.bias", "norm2.weight", "norm2.bias", "mlp.fc1.bias", "mlp.fc2.bias"]
        for n, g in _grads(m, small).items():
            out[f"{tag}_d_{n}"] = g
        for n, g in _grads(m, ["attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight"]).items():
            out[f"{tag}_d_{n}_rows"] = g[::23]
        out[f"{tag}_insum"] = np.array(checksum(x, w, *[v for v in params.values() if v.is_floating_point()]))
    np.savez_compressed(os.path.join(OUT, "swin_block.npz"), **out)


LAYER_CASE = dict(dim=128, res=(16, 24), heads=2, B=1, seed=51)


def gen_layer(sw) -> None:
    c = LAYER_CASE
    H, W = c["res"]
    params = so.make_layer_params(c["dim"], c["res"], c["heads"], c["seed"])
    m = sw.SwinTransformerLayerv5(dim=c["dim"], input_resolution=c["res"], num_heads=c["heads"])
    _load(m, params)
    x = so.make_features(61, c["B"], 4, c["dim"], H, W).requires_grad_(True)
    w1 = so.make_features(62, c["B"], 4, c["dim"], H, W) - 0.4
    w2 = so.make_features(63, c["B"], 4, 2 * c["dim"], H // 2, W // 2) - 0.4
    y1, y2 = m(x)
    ((y1 * w1).sum() + (y2 * w2).sum()).backward()
    out = {"y1": _np(y1), "y2": _np(y2), "dx": _np(x.grad)}
    names = ["layers.0.0.attn.relative_position_bias_table", "layers.1.1.attn.relative_position_bias_table",
             "layers.4.1.attn.relative_position_bias_table", "layers.2.1.norm1.weight", "layers.5.0.mlp.fc2.bias",
             "downsample.norm.weight", "downsample.norm.bias"]
    for n, g in _grads(m, names).items():
        out[f"d_{n}"] = g
    out["d_downsample.reduction.weight_rows"] = _grads(m, ["downsample.reduction.weight"])["downsample.reduction.weight"][::29]
    out["d_layers.1.0.attn.qkv.weight_rows"] = _grads(m, ["layers.1.0.attn.qkv.weight"])["layers.1.0.attn.qkv.weight"][::29]
    out["insum"] = np.array(checksum(x, w1, w2, *[v for v in params.values() if v.is_floating_point()]))
    np.savez_compressed(os.path.join(OUT, "swin_layer.npz"), **out)


LOSS_CASES = [  # (tag, N, C, H, W, class_num, special)
    ("l0", 2, 64, 8, 14, 12, None),
    ("l1", 1, 256, 16, 24, 12, None),
    ("l2", 2, 64, 8, 14, 9, "absent_class"),
    ("l3", 2, 64, 8, 14, 18, "single_class_set"),
    ("l4", 4, 64, 8, 14, 26, None),
]


def loss_case_inputs(tag, N, C, H, W, K, special):
    """Shared by the golden generator and the tests."""
    labels = lo.make_label_maps(71, 6, N, H, W, K)
    if special == "absent_class":          # remove class 3 from key set adj2 (index 3)
        labels[3] = torch.where(labels[3] == 3, torch.full_like(labels[3], 4.0), labels[3])
    if special == "single_class_set":      # neg3 (index 5) is one class everywhere
        labels[5] = torch.full_like(labels[5], 2.0)
    emb = lo.make_embeddings(72, labels, C, K)
    return labels, emb


def gen_loss(px) -> None:
    out = {}
    for (tag, N, C, H, W, K, special) in LOSS_CASES:
        labels, emb = loss_case_inputs(tag, N, C, H, W, K, special)
        q = emb[0].clone().requires_grad_(True)
        loss = px.regression_loss(q, emb[1], emb[2], emb[3], emb[4], emb[5],
                                  labels[0], labels[1], labels[2], labels[3], labels[4], labels[5], K)
        loss.backward()
        out[f"{tag}_loss"] = np.array(loss.item(), dtype=np.float64)
        out[f"{tag}_dq"] = _np(q.grad)
        out[f"{tag}_insum"] = np.array(checksum(*labels, *emb))
    # posMask / negMask on a tiny label pair
    l1, l2 = lo.make_label_maps(73, 2, 2, 4, 6, 5, coarse=(2, 3))
    out["pos_small"] = _np(px.posMask(l1, l2, 5)).astype(np.int8)
    out["neg_small"] = _np(px.negMask(l1, l2, 5)).astype(np.int8)
    # nearest label down-sampling, the reference's 256x448 -> 32x56 case and an uneven one
    big = lo.make_label_maps(74, 1, 1, 256, 448, 12, coarse=(16, 28))[0]
    out["down_32x56"] = _np(torch.nn.functional.interpolate(big, size=[32, 56], mode="nearest")).astype(np.int8)
    odd = lo.make_label_maps(75, 1, 1, 100, 150, 12, coarse=(10, 15))[0]
    out["down_odd_24x40"] = _np(torch.nn.functional.interpolate(odd, size=[24, 40], mode="nearest")).astype(np.int8)
    # the symmetric ConsistencyLoss tail (PixPro_swin_v5.py:584-597) on full-res labels
    N, C, H, W, K = 2, 64, 8, 14, 12
    full = lo.make_label_maps(76, 6, N, 64, 112, K, coarse=(4, 7))
    ds = [torch.nn.functional.interpolate(m, size=[H, W], mode="nearest") for m in full]
    emb = lo.make_embeddings(77, ds, C, K)
    pred_1, pred_2 = emb[0], lo.make_embeddings(78, ds[1:2], C, K)[0]
    proj_1, proj_2 = lo.make_embeddings(79, ds[0:1], C, K)[0], emb[1]
    tail = (px.regression_loss(pred_1, proj_2, emb[2], emb[3], emb[4], emb[5], ds[0], ds[1], ds[2], ds[3], ds[4], ds[5], K)
            + px.regression_loss(pred_2, proj_1, emb[2], emb[3], emb[4], emb[5], ds[1], ds[0], ds[2], ds[3], ds[4], ds[5], K))
    out["tail_loss"] = np.array(tail.item(), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "loss_cases.npz"), **out)


def gen_state_dict_contract(sw) -> None:
    """Key names / shapes of the reference layer's state_dict (tests/test_state_dict_contract.py)."""
    import json
    small = sw.SwinTransformerLayerv5(dim=128, input_resolution=(16, 24), num_heads=2)
    default = sw.SwinTransformerLayerv5()
    out = {"layer_128_16x24_h2": {k: list(v.shape) for k, v in small.state_dict().items()},
           "layer_default": {k: list(v.shape) for k, v in default.state_dict().items()},
           "layer_default_param_count": sum(p.numel() for p in default.parameters())}
    with open(os.path.join(OUT, "state_dict_contract.json"), "w") as f:
        json.dump(out, f)


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    sw = ref_shims.import_swin()
    px = ref_shims.import_pixpro()
    gen_index(sw)
    gen_attention(sw)
    gen_block(sw)
    gen_layer(sw)
    gen_loss(px)
    gen_state_dict_contract(sw)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
