"""Generate tests/golden/trainaux_cases.npz by running the UNMODIFIED reference classes
(OhemCELoss2D, LARS + add_weight_decay, PixPro._momentum_update_key_encoder).

TEST INFRASTRUCTURE.  Authoring container only (needs /root/reference):
    python -m oracle.make_goldens_trainaux
"""
from __future__ import annotations

import importlib.util
import math
import os
import types

import numpy as np
import torch

from . import ref_shims
from . import trainaux_oracle as ta

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
LARS_RUNS = [("wd", 1e-5, 0.9), ("wd0", 0.0, 0.9), ("nomom", 1e-4, 0.0)]      # (tag, weight decay, momentum)


def _load(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_shims.REF_ROOT, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _np(t):
    return t.detach().cpu().numpy().copy()          # a copy: parameters are updated in place afterwards


class _Holder(torch.nn.Module):
    """named_parameters() in the insertion order of a dict of tensors (dots are not allowed in names)."""

    def __init__(self, tensors):
        super().__init__()
        for k, v in tensors.items():
            self.register_parameter(k.replace(".", "_"), torch.nn.Parameter(v.clone()))


def gen_ohem(out) -> None:
    losses = _load("seg18/utils/losses.py", "ref_seg18_losses")
    for i, (tag, B, K, H, W, n_min, margin, ign) in enumerate(ta.OHEM_CASES):
        logits, labels = ta.make_ohem_case(ta.ohem_seed(i), B, K, H, W, margin, ign)
        if n_min < 0:                    # exactly n_min, resp. n_min + 1, losses above the threshold
            px = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-1, reduction="none").view(-1)
            n_min = int((px > -math.log(ta.OHEM_THRESH)).sum()) + 1 + n_min
        out[f"ohem_{tag}_nmin"] = np.array(n_min)
        x = logits.clone().requires_grad_(True)
        crit = losses.OhemCELoss2D(n_min)
        loss = crit(x, labels)
        loss.backward()
        out[f"ohem_{tag}_loss"] = np.array(loss.item(), dtype=np.float64)
        out[f"ohem_{tag}_dlogits"] = _np(x.grad)
        out[f"ohem_{tag}_insum"] = np.array(float(logits.double().abs().sum() + labels.double().abs().sum()))
        px = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-1, reduction="none").view(-1)
        out[f"ohem_{tag}_branch"] = np.array(int(torch.sort(px, descending=True)[0][n_min] > crit.thresh))


def gen_lars(out) -> None:
    lars = _load("pixcontrast_18/contrast/lars.py", "ref_contrast_lars")
    for tag, wd, mom in LARS_RUNS:
        params = ta.make_param_set(7)
        model = _Holder(params)
        opt = lars.LARS(torch.optim.SGD(lars.add_weight_decay(model, wd), lr=0.5, momentum=mom))
        named = dict(model.named_parameters())
        for step in range(3):
            grads = ta.make_grads(20 + step, params)
            for k in params:
                named[k.replace(".", "_")].grad = grads[k].clone()
            opt.step()
            for k in params:
                p = named[k.replace(".", "_")]
                out[f"lars_{tag}_s{step}_p_{k}"] = _np(p)
                out[f"lars_{tag}_s{step}_g_{k}"] = _np(p.grad)
                if mom != 0:
                    out[f"lars_{tag}_s{step}_b_{k}"] = _np(opt.state[p]["momentum_buffer"])
    out["lars_insum"] = np.array(float(sum(v.double().abs().sum() for v in ta.make_param_set(7).values())))


def gen_ema(out) -> None:
    px = ref_shims.import_pixpro()
    q, k = ta.make_param_set(31), ta.make_param_set(32)
    names = ["encoder_1", "encoder_2", "encoder_3", "proj1", "proj2", "proj3", "projector"]
    keys = ["encoder_k_1", "encoder_k_2", "encoder_k_3", "proj_k_1", "proj_k_2", "proj_k_3", "projector_k"]
    stub = types.SimpleNamespace(pixpro_momentum=0.99, k=3, K=40, pixpro_ins_loss_weight=0.)
    items = list(q.keys())
    for i, (nq, nk) in enumerate(zip(names, keys)):       # spread the tensors over the seven encoder / head pairs
        sub = items[i::len(names)]
        setattr(stub, nq, _Holder({n: q[n] for n in sub}))
        setattr(stub, nk, _Holder({n: k[n] for n in sub}))
    for rep in range(2):
        px.PixPro._momentum_update_key_encoder(stub)
    assert stub.k == 5
    for i, (nq, nk) in enumerate(zip(names, keys)):
        for n, p in getattr(stub, nk).named_parameters():
            out[f"ema_{n}"] = _np(p)
    out["ema_momenta"] = np.array([ta.cosine_momentum(0.99, 3, 40), ta.cosine_momentum(0.99, 4, 40)], dtype=np.float64)


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    out = {}
    gen_ohem(out)
    gen_lars(out)
    gen_ema(out)
    path = os.path.join(OUT, "trainaux_cases.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), len(out), "arrays;",
          {k: int(v) for k, v in out.items() if k.endswith("_branch")})


if __name__ == "__main__":
    main()
