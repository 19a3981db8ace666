"""CPU: the oracle restatement against the fixtures produced by the reference
itself (oracle/make_goldens.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import index_oracle as ix
from oracle import loss_oracle as lo
from oracle import swin_oracle as so
from oracle import make_goldens as mg
from conftest import GOLDEN, rel_err

TOL = 2e-5   # fp32 restatement vs fp32 reference: different summation order only


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("case", mg.INDEX_CASES)
def test_index_closed_forms_exact(case):
    H, W, ws, shift = case
    g = _load("index_cases.npz")
    tag = f"{H}x{W}_ws{ws}_s{shift}"
    ews, eshift = ix.effective_window((H, W), ws, shift)
    assert [ews, eshift] == g[f"eff_{tag}"].tolist()
    assert np.array_equal(ix.relative_position_index(ews), g[f"relidx_{tag}"].astype(np.int64))
    m = ix.shift_attn_mask(H, W, ews, eshift)
    if eshift == 0:
        assert m is None and f"mask_{tag}" not in g.files
    else:
        assert np.array_equal(m, g[f"mask_{tag}"].astype(np.float32))
    gi = ix.window_gather_index(H, W, ews, eshift)
    assert np.array_equal(gi, g[f"gather_{tag}"].astype(np.int64))
    # scatter: slot id stored at each source token must be (win, t, n) of the gather map
    sc = g[f"scatter_{tag}"].astype(np.int64)           # [2, H*W] -> slot = (win*2 + t)*N + n
    N = ews * ews
    for t in range(2):
        win, n = np.divmod(np.arange(gi.size), N)
        assert np.array_equal(sc[t][gi.reshape(-1)], (win * 2 + t) * N + n)


def test_mask_nonzero_only_in_last_row_and_column():
    m = ix.shift_attn_mask(64, 80, 8, 4)
    nz = np.nonzero(np.abs(m).reshape(80, -1).sum(1))[0]
    expect = sorted(set(range(70, 80)) | set(range(9, 80, 10)))
    assert nz.tolist() == expect and len(expect) == 17


@pytest.mark.parametrize("case", mg.ATTN_CASES, ids=lambda c: c[0])
def test_window_attention_matches_reference(case):
    tag, dim, ws, heads, T, (H, W), shift, B = case
    g = _load("swin_attention.npz")
    nW = (H // ws) * (W // ws)
    params = so.make_attention_params(dim, ws, heads, seed=11)
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in params.items()}
    x = (so.make_features(21, B * nW, T, ws * ws, dim) - 0.4).requires_grad_(True)
    w = so.make_features(22, B * nW, T, ws * ws, dim) - 0.4
    assert abs(mg.checksum(x, w, *[v for v in params.values() if v.is_floating_point()]) - float(g[f"{tag}_insum"])) < 1e-6 * float(g[f"{tag}_insum"])
    mask_np = ix.shift_attn_mask(H, W, ws, shift)
    mask = torch.from_numpy(mask_np) if mask_np is not None else None
    y = so.window_attention(x, leaf, ws, heads, mask)
    (y * w).sum().backward()
    assert rel_err(y, g[f"{tag}_y"]) < TOL
    assert rel_err(x.grad, g[f"{tag}_dx"]) < TOL
    assert rel_err(leaf["relative_position_bias_table"].grad, g[f"{tag}_d_relative_position_bias_table"]) < TOL
    assert rel_err(leaf["qkv.bias"].grad, g[f"{tag}_d_qkv.bias"]) < TOL
    assert rel_err(leaf["qkv.weight"].grad[::17], g[f"{tag}_d_qkv.weight_rows"]) < TOL
    assert rel_err(leaf["proj.weight"].grad[::17], g[f"{tag}_d_proj.weight_rows"]) < TOL


@pytest.mark.parametrize("case", mg.BLOCK_CASES, ids=lambda c: c[0])
def test_block_matches_reference(case):
    tag, dim, (H, W), heads, ws, shift, T, B = case
    g = _load("swin_block.npz")
    params = so.make_block_params(dim, (H, W), heads, ws, shift, seed=31)
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != "attn_mask" else v) for k, v in params.items()}
    x = so.make_features(41, B, T, H * W, dim).requires_grad_(True)
    w = so.make_features(42, B, T, H * W, dim) - 0.4
    y = so.swin_block(x, leaf, (H, W), heads, ws, shift)
    (y * w).sum().backward()
    assert rel_err(y, g[f"{tag}_y"]) < TOL
    assert rel_err(x.grad, g[f"{tag}_dx"]) < 5 * TOL
    for n in ["attn.relative_position_bias_table", "attn.qkv.bias", "norm1.weight", "norm2.bias", "mlp.fc1.bias", "mlp.fc2.bias"]:
        assert rel_err(leaf[n].grad, g[f"{tag}_d_{n}"]) < 5 * TOL, n
    for n in ["attn.qkv.weight", "mlp.fc1.weight", "mlp.fc2.weight"]:
        assert rel_err(leaf[n].grad[::23], g[f"{tag}_d_{n}_rows"]) < 5 * TOL, n


def test_layer_matches_reference():
    c = mg.LAYER_CASE
    g = _load("swin_layer.npz")
    H, W = c["res"]
    params = so.make_layer_params(c["dim"], c["res"], c["heads"], c["seed"])
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("attn_mask") else v) for k, v in params.items()}
    x = so.make_features(61, c["B"], 4, c["dim"], H, W).requires_grad_(True)
    w1 = so.make_features(62, c["B"], 4, c["dim"], H, W) - 0.4
    w2 = so.make_features(63, c["B"], 4, 2 * c["dim"], H // 2, W // 2) - 0.4
    y1, y2 = so.swin_layer_v5(x, leaf, c["dim"], c["res"], c["heads"])
    ((y1 * w1).sum() + (y2 * w2).sum()).backward()
    assert rel_err(y1, g["y1"]) < 5 * TOL
    assert rel_err(y2, g["y2"]) < 5 * TOL
    assert rel_err(x.grad, g["dx"]) < 2e-4
    for n in ["layers.0.0.attn.relative_position_bias_table", "layers.4.1.attn.relative_position_bias_table",
              "layers.2.1.norm1.weight", "downsample.norm.weight"]:
        assert rel_err(leaf[n].grad, g[f"d_{n}"]) < 2e-4, n
    assert rel_err(leaf["downsample.reduction.weight"].grad[::29], g["d_downsample.reduction.weight_rows"]) < 2e-4


@pytest.mark.parametrize("case", mg.LOSS_CASES, ids=lambda c: c[0])
def test_regression_loss_matches_reference(case):
    tag, N, C, H, W, K, special = case
    g = _load("loss_cases.npz")
    labels, emb = mg.loss_case_inputs(*case)
    assert abs(mg.checksum(*labels, *emb) - float(g[f"{tag}_insum"])) < 1e-6 * float(g[f"{tag}_insum"])
    q = emb[0].clone().requires_grad_(True)
    loss = lo.regression_loss(q, emb[1:], labels[0], labels[1:], K)
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}_loss"])) < 1e-5 * abs(float(g[f"{tag}_loss"]))
    assert rel_err(q.grad, g[f"{tag}_dq"]) < 1e-4
    proto = lo.regression_loss_prototype(emb[0].double(), [e.double() for e in emb[1:]], labels[0], labels[1:], K)
    assert abs(proto.item() - float(g[f"{tag}_loss"])) < 1e-5 * abs(float(g[f"{tag}_loss"]))


def test_label_masks_and_downsampling_exact():
    g = _load("loss_cases.npz")
    l1, l2 = lo.make_label_maps(73, 2, 2, 4, 6, 5, coarse=(2, 3))
    assert np.array_equal(ix.label_match(l1.numpy(), l2.numpy()), g["pos_small"].astype(np.float32))
    assert np.array_equal(ix.label_differ(l1.numpy(), l2.numpy()), g["neg_small"].astype(np.float32))
    big = lo.make_label_maps(74, 1, 1, 256, 448, 12, coarse=(16, 28))[0].numpy()
    assert np.array_equal(ix.downsample_labels(big, 32, 56), g["down_32x56"].astype(np.float32))
    assert np.array_equal(ix.downsample_labels(big, 32, 56), big[..., ::8, ::8])
    odd = lo.make_label_maps(75, 1, 1, 100, 150, 12, coarse=(10, 15))[0].numpy()
    assert np.array_equal(ix.downsample_labels(odd, 24, 40), g["down_odd_24x40"].astype(np.float32))


def test_consistency_tail_matches_reference():
    g = _load("loss_cases.npz")
    N, C, H, W, K = 2, 64, 8, 14, 12
    full = lo.make_label_maps(76, 6, N, 64, 112, K, coarse=(4, 7))
    ds = [torch.from_numpy(ix.downsample_labels(m.numpy(), H, W)) for m in full]
    emb = lo.make_embeddings(77, ds, C, K)
    pred_1, pred_2 = emb[0], lo.make_embeddings(78, ds[1:2], C, K)[0]
    proj_1, proj_2 = lo.make_embeddings(79, ds[0:1], C, K)[0], emb[1]
    tail = lo.consistency_tail(pred_1, pred_2, proj_1, proj_2, emb[2:], full[0], full[1], full[2:], K)
    assert abs(tail.item() - float(g["tail_loss"])) < 1e-5 * abs(float(g["tail_loss"]))


def test_out_of_range_label_raises():
    labels, emb = mg.loss_case_inputs("l0", 1, 64, 8, 14, 12, None)
    bad = labels[2].clone(); bad[0, 0, 0, 0] = 12.0
    with pytest.raises(RuntimeError):
        lo.regression_loss(emb[0][:1], [e[:1] for e in emb[1:]], labels[0][:1], [labels[1][:1], bad[:1], *[l[:1] for l in labels[3:]]], 12)


def test_restated_caller_model_matches_the_reference_full_model():
    """oracle/tswin_oracle.py (TswinPlus / ResNet18_OS8 / ASPP restated around a pluggable Swin head) against
    tests/golden/tswinplus.npz, written from the UNMODIFIED reference model by oracle/make_goldens_tswin.py: the state_dict
    key list is identical (strict load of the synthetic weights) and the eval-mode logits of the 512x640 clip agree
    to fp32 rounding, with the functional head oracle plugged in for seg18/net/Ours/swin_512.py."""
    import json
    from oracle import tswin_oracle as to
    g = np.load(os.path.join(GOLDEN, "tswinplus.npz"))
    keys = json.loads(str(g["keys"]))
    model = to.TswinPlus(12, to.OracleSwin(to.swin_shapes_from_keys(keys)))
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert ours == keys
    model.load_state_dict(to.synth_state_dict(model.state_dict(), 5), strict=True)
    model.eval()
    with torch.no_grad():
        logits = model(to.make_clip(6))
    assert float(np.abs(logits[0, :, ::8, ::8].numpy() - g["logits_sub"]).max()) < 1e-3 * float(g["logit_absmax"])
    am = logits.argmax(1)[0].numpy()
    decided = g["margin"].astype(np.float32) > 1e-3 * float(g["logit_absmax"])
    assert np.array_equal(am[decided], g["argmax"][decided]) and decided.mean() > 0.99
