"""CPU: the training-step oracle (oracle/trainaux_oracle.py) against the fixtures produced by the reference's own
OhemCELoss2D, LARS and PixPro._momentum_update_key_encoder (oracle/make_goldens_trainaux.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import make_goldens_trainaux as mgt
from oracle import trainaux_oracle as ta
from conftest import GOLDEN, rel_err

TOL = 2e-5   # fp32 restatement vs fp32 reference


def _g():
    return np.load(os.path.join(GOLDEN, "trainaux_cases.npz"))


def test_both_ohem_branches_and_the_boundary_are_covered():
    g = _g()
    branches = {c[0]: int(g[f"ohem_{c[0]}_branch"]) for c in ta.OHEM_CASES}
    assert set(branches.values()) == {0, 1}
    assert branches["boundary_eq"] == 0 and branches["boundary_gt"] == 1
    assert int(g["ohem_boundary_eq_nmin"]) == int(g["ohem_boundary_gt_nmin"]) + 1


@pytest.mark.parametrize("idx", range(len(ta.OHEM_CASES)))
def test_ohem_oracle_vs_reference(idx):
    tag, B, K, H, W, _, margin, ign = ta.OHEM_CASES[idx]
    g = _g()
    logits, labels = ta.make_ohem_case(ta.ohem_seed(idx), B, K, H, W, margin, ign)
    assert abs(float(logits.double().abs().sum() + labels.double().abs().sum()) - float(g[f"ohem_{tag}_insum"])) < 1e-6
    x = logits.clone().requires_grad_(True)
    loss = ta.ohem_ce(x, labels, int(g[f"ohem_{tag}_nmin"]))
    loss.backward()
    assert abs(float(loss) - float(g[f"ohem_{tag}_loss"])) < TOL * abs(float(g[f"ohem_{tag}_loss"]))
    assert rel_err(x.grad, g[f"ohem_{tag}_dlogits"]) < TOL


@pytest.mark.parametrize("run", mgt.LARS_RUNS)
def test_lars_oracle_vs_reference(run):
    tag, wd, mom = run
    g = _g()
    params = ta.make_param_set(7)
    assert abs(float(sum(v.double().abs().sum() for v in params.values())) - float(g["lars_insum"])) < 1e-6
    bufs = {k: None for k in params}
    for step in range(3):
        grads = ta.make_grads(20 + step, params)
        for k in params:
            one_d = params[k].dim() == 1
            p, gr, b = ta.lars_sgd_step(params[k], grads[k], bufs[k], lr=0.5, momentum=mom, weight_decay=0.0 if one_d else wd,
                                        lars=not one_d)
            params[k], bufs[k] = p, b
            assert rel_err(p, g[f"lars_{tag}_s{step}_p_{k}"]) < TOL, (k, step)
            assert float((gr.double() - torch.as_tensor(g[f"lars_{tag}_s{step}_g_{k}"]).double()).abs().max()) <= \
                TOL * max(1e-30, float(np.abs(g[f"lars_{tag}_s{step}_g_{k}"]).max())), (k, step)
            if mom != 0:
                assert rel_err(b, g[f"lars_{tag}_s{step}_b_{k}"]) < TOL or float(b.abs().max()) == 0.0, (k, step)


def test_ema_oracle_vs_reference_bit_exact():
    g = _g()
    q, k = ta.make_param_set(31), ta.make_param_set(32)
    for step in (3, 4):
        m = ta.cosine_momentum(0.99, step, 40)
        k = {n: ta.ema(k[n], q[n], m) for n in k}
    assert np.allclose(g["ema_momenta"], [ta.cosine_momentum(0.99, 3, 40), ta.cosine_momentum(0.99, 4, 40)], rtol=0, atol=0)
    for n in k:
        assert np.array_equal(k[n].numpy(), g[f"ema_{n.replace('.', '_')}"]), n
