"""CPU oracle for the training-step kernels around the hot paths.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Functional torch fp32 restatements of
  * OhemCELoss2D.forward            seg18/utils/losses.py:29-40
  * LARS.step over torch.optim.SGD  pixcontrast_18/contrast/lars.py:109-152
  * the key-encoder momentum update pixcontrast_18/contrast/models/PixPro_swin_v5.py:258-289
plus the seeded input builders shared by the golden generator (oracle/make_goldens_trainaux.py)
and the tests.  Parity status: PINNED against the reference classes themselves
(tests/golden/trainaux_cases.npz, checked by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

OHEM_THRESH = 0.7

# (tag, B, K, H, W, n_min, margin, ignore_fraction): margin is added to the logit of the true class, so a large
# margin makes most pixels easy (top-n_min branch, losses.py:39) and margin 0 keeps them hard (threshold branch, :37)
OHEM_CASES = [
    ("hard", 2, 12, 16, 24, 48, 0.0, 0.0),
    ("easy", 2, 12, 16, 24, 48, 12.0, 0.0),
    ("easy_ignore", 2, 11, 16, 24, 48, 12.0, 0.3),
    ("easy_odd", 1, 12, 7, 9, 5, 12.0, 0.2),
    ("hard_ignore_odd", 1, 12, 7, 9, 5, 0.5, 0.2),
    ("mixed_36", 1, 36, 12, 20, 40, 5.0, 0.1),
    # n_min -1 / -2: set by the golden generator to (number of losses above the threshold) / (that number - 1):
    # the two sides of the `loss[n_min] > thresh` test (losses.py:36); stored in the fixture as ohem_<tag>_nmin
    ("boundary_eq", 2, 5, 8, 8, -1, 4.0, 0.0),
    ("boundary_gt", 2, 5, 8, 8, -2, 4.0, 0.0),
    ("mostly_ignored", 1, 6, 8, 12, 60, 0.0, 0.7),      # fewer labelled pixels than n_min: zero losses enter the top n_min
]


def ohem_seed(idx: int) -> int:
    """Seed of OHEM_CASES[idx]; the two boundary cases share their inputs."""
    return 100 + (idx - 1 if OHEM_CASES[idx][0] == "boundary_gt" else idx)


def make_ohem_case(seed: int, B: int, K: int, H: int, W: int, margin: float, ignore_fraction: float, ignore_index: int = -1):
    """-> logits [B,K,H,W] fp32, labels [B,H,W] int64 (some set to ignore_index)."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, K, (B, H, W), generator=g)
    logits = torch.randn(B, K, H, W, generator=g) * 1.5
    logits = logits + margin * F.one_hot(labels, K).permute(0, 3, 1, 2).float() * (0.75 + 0.5 * torch.rand(B, 1, H, W, generator=g))
    if ignore_fraction > 0:
        drop = torch.rand(B, H, W, generator=g) < ignore_fraction
        labels = torch.where(drop, torch.full_like(labels, ignore_index), labels)
    return logits, labels


def ohem_ce(logits: torch.Tensor, labels: torch.Tensor, n_min: int, thresh: float = OHEM_THRESH, ignore_index: int = -1):
    """losses.py:32-40 without the sort: the k-th largest loss decides the branch, top-k replaces sorted[:n_min]."""
    t = -math.log(thresh)                                               # :26
    loss = F.cross_entropy(logits, labels, ignore_index=ignore_index, reduction="none").view(-1)   # :33
    kth = torch.topk(loss.detach(), n_min + 1).values[n_min]           # == sorted_desc[n_min], :36
    if kth > t:
        return loss[loss > t].mean()                                    # :37
    return torch.topk(loss, n_min).values.mean()                        # :39


def lars_sgd_step(p, g, buf, *, lr, momentum, weight_decay, lars, trust_coef=0.001, eps=1e-8, dampening=0.0, nesterov=False):
    """One tensor of LARS.step: -> (new p, the gradient left in p.grad, new momentum buffer).
    lars.py:121-135 (decay, norms, adaptive lr) then SGD with weight decay zeroed (:140-146)."""
    if weight_decay > 0:
        g = g + weight_decay * p
    if lars:
        pn, gn = p.norm(), g.norm()
        if pn > 0 and gn > 0:
            g = g * (trust_coef * pn / (gn + eps))
    d = g
    if momentum != 0:
        buf = g.clone() if buf is None else momentum * buf + (1 - dampening) * g
        d = g + momentum * buf if nesterov else buf
    return p - lr * d, g, buf


def ema(k: torch.Tensor, q: torch.Tensor, m: float) -> torch.Tensor:
    """PixPro_swin_v5.py:266-267."""
    return k * m + q * (1. - m)


def cosine_momentum(base: float, k: int, K: int) -> float:
    """PixPro_swin_v5.py:263."""
    return 1. - (1. - base) * (math.cos(math.pi * k / K) + 1) / 2.


def make_param_set(seed: int):
    """A small parameter set shaped like a slice of the pre-training model: 2-D weights (LARS group), 1-D biases /
    norm weights (ignore group), one all-zero weight and sizes that are not multiples of 4."""
    g = torch.Generator().manual_seed(seed)
    shapes = {"fc1.weight": (48, 32), "fc1.bias": (48,), "norm.weight": (48,), "norm.bias": (48,), "fc2.weight": (10, 48),
              "fc2.bias": (10,), "conv.weight": (6, 3, 3, 3), "odd.weight": (7, 9), "zero.weight": (5, 8), "big.weight": (130, 67)}
    params = {k: torch.randn(*s, generator=g) * 0.3 for k, s in shapes.items()}
    params["zero.weight"].zero_()
    return params


def make_grads(seed: int, params):
    g = torch.Generator().manual_seed(seed)
    grads = {k: torch.randn(*v.shape, generator=g) * 0.05 for k, v in params.items()}
    grads["odd.weight"].zero_()                       # grad_norm == 0 together with weight_decay 0 in one of the runs
    return grads
