"""Floating-point oracle for HP-2 (label-guided pixel contrastive loss).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates, for CPU torch
(fp32/fp64, differentiable):

  * regression_loss         pixcontrast_18/contrast/models/PixPro_swin_v5.py:71-129
  * F.normalize(dim=1)      pixcontrast_18/contrast/models/PixPro_swin_v5.py:330,362,400,...
  * ConsistencyLoss tail    pixcontrast_18/contrast/models/PixPro_swin_v5.py:584-597

in the per-row form of SURVEY.md Appx B.1: for query pixel i with label l_i and
key sets s in (k, adj1, adj2, adj3, neg3),

  P_i = sum_s sum_j [l_i == l_sj] z_sij / (sum_s sum_j [l_i == l_sj] + 1e-6)
  N_i = sum_s ( sum_j [l_i != l_sj] z_sij / (sum_j [l_i != l_sj] + 1e-6) )
  loss = -mean_{b,i} log( e^P / (e^P + e^N) + 1e-6 )

A second, independent form (per-class key sums, Appx B.2) is provided as a
cross-check.  Pinned against the reference by ``tests/golden/loss_*.npz``.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import index_oracle as ix

EPS_CNT = 1e-6     # PixPro_swin_v5.py:119-123
EPS_LOG = 1e-6     # PixPro_swin_v5.py:127-128
EPS_NORM = 1e-12   # F.normalize default


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, dim=1): x / max(||x||_2, 1e-12) over channels."""
    n = x.pow(2).sum(dim=1, keepdim=True).sqrt().clamp_min(EPS_NORM)
    return x / n


def _labels(lbl: torch.Tensor) -> torch.Tensor:
    # .long() truncation of a float / u8 [N,1,H,W] label map (PixPro_swin_v5.py:54)
    return lbl.reshape(lbl.shape[0], -1).to(torch.float64).trunc().to(torch.int64)


def regression_loss(q: torch.Tensor, keys: Sequence[torch.Tensor], label_q: torch.Tensor,
                    labels_k: Sequence[torch.Tensor], class_num: int) -> torch.Tensor:
    """Dense form.  q, keys[s]: [N, C, H, W] (already L2-normalised by the caller,
    as in the reference); label_q, labels_k[s]: [N, 1, H, W].  Returns a 0-dim tensor.

    ``keys`` is ordered (k, adj1, adj2, adj3, neg3) like the reference's positional
    arguments, but any number of sets >= 1 is accepted (extra sets extend the
    P-pool and the N-sum)."""
    Nb, C = q.shape[:2]
    qf = q.reshape(Nb, C, -1)
    lq = _labels(label_q)
    if int(lq.max()) >= class_num or int(lq.min()) < 0:
        raise RuntimeError("Class values must be smaller than num_classes.")   # F.one_hot behaviour
    pos_sum = torch.zeros(qf.shape[0], qf.shape[2], dtype=q.dtype)
    pos_cnt = torch.zeros_like(pos_sum)
    neg_term = torch.zeros_like(pos_sum)
    for kk, lk in zip(keys, labels_k):
        kf = kk.reshape(Nb, C, -1)
        lkk = _labels(lk)
        if int(lkk.max()) >= class_num or int(lkk.min()) < 0:
            raise RuntimeError("Class values must be smaller than num_classes.")
        z = torch.einsum("nci,ncj->nij", qf, kf)
        same = (lq[:, :, None] == lkk[:, None, :]).to(q.dtype)
        diff = 1.0 - same
        pos_sum = pos_sum + (same * z).sum(-1)
        pos_cnt = pos_cnt + same.sum(-1)
        neg_term = neg_term + (diff * z).sum(-1) / (diff.sum(-1) + EPS_CNT)
    P = pos_sum / (pos_cnt + EPS_CNT)
    eP, eN = torch.exp(P), torch.exp(neg_term)
    return -torch.mean(torch.log(eP / (eP + eN) + EPS_LOG))


def regression_loss_prototype(q, keys, label_q, labels_k, class_num: int) -> torch.Tensor:
    """Same value through per-class key sums (SURVEY Appx B.2) -- cross-check only."""
    Nb, C = q.shape[:2]
    qf = q.reshape(Nb, C, -1)
    lq = _labels(label_q)
    pos_sum = torch.zeros(Nb, qf.shape[2], dtype=q.dtype)
    pos_cnt = torch.zeros_like(pos_sum)
    neg_term = torch.zeros_like(pos_sum)
    for kk, lk in zip(keys, labels_k):
        kf = kk.reshape(Nb, C, -1)
        onehot = torch.nn.functional.one_hot(_labels(lk), class_num).to(q.dtype)    # [N, HW, K]
        proto = torch.einsum("ncj,njk->nck", kf, onehot)                            # per-class key sums
        hist = onehot.sum(1)                                                        # [N, K]
        dots = torch.einsum("nci,nck->nik", qf, proto)                              # [N, HW, K]
        same_sum = dots.gather(2, lq[:, :, None]).squeeze(2)
        same_cnt = hist.gather(1, lq)
        all_sum = dots.sum(2)
        pos_sum = pos_sum + same_sum
        pos_cnt = pos_cnt + same_cnt
        neg_term = neg_term + (all_sum - same_sum) / (kf.shape[2] - same_cnt + EPS_CNT)
    P = pos_sum / (pos_cnt + EPS_CNT)
    eP, eN = torch.exp(P), torch.exp(neg_term)
    return -torch.mean(torch.log(eP / (eP + eN) + EPS_LOG))


def consistency_tail(pred_1, pred_2, proj_1, proj_2, shared_keys, mask_1, mask_2, shared_masks,
                     class_num: int) -> torch.Tensor:
    """ConsistencyLoss.forward after the encoders (PixPro_swin_v5.py:584-597):
    nearest label down-sampling to the embedding resolution, then the symmetric sum
    of two regression_loss calls that share the (adj1, adj2, adj3, neg3) key sets."""
    H, W = pred_1.shape[-2:]
    ds = lambda m: torch.from_numpy(ix.downsample_labels(m.numpy(), H, W))
    m1, m2 = ds(mask_1), ds(mask_2)
    ms = [ds(m) for m in shared_masks]
    return (regression_loss(pred_1, [proj_2, *shared_keys], m1, [m2, *ms], class_num)
            + regression_loss(pred_2, [proj_1, *shared_keys], m2, [m1, *ms], class_num))


# ----------------------------------------------------------------------------
# deterministic synthetic inputs
# ----------------------------------------------------------------------------

def make_label_maps(seed: int, n_maps: int, N: int, H: int, W: int, class_num: int,
                    coarse=(4, 7)) -> list:
    """Spatially coherent integer label maps [N,1,H,W] (float32 like the
    reference's loader output after .float()), built by nearest-upsampling a
    coarse random map."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_maps):
        c = torch.randint(0, class_num, (N, 1, coarse[0], coarse[1]), generator=g)
        ih = torch.from_numpy(ix.nearest_resize_index(coarse[0], H))
        iw = torch.from_numpy(ix.nearest_resize_index(coarse[1], W))
        out.append(c[:, :, ih[:, None], iw[None, :]].to(torch.float32).contiguous())
    return out


def make_embeddings(seed: int, labels: Sequence[torch.Tensor], C: int, class_num: int,
                    noise: float = 0.5) -> list:
    """normalize(proto[label] + noise * N(0,1)) -- purely random embeddings give the
    degenerate loss ln 2 and hide bugs."""
    g = torch.Generator().manual_seed(seed)
    proto = torch.randn(class_num, C, generator=g)
    out = []
    for lbl in labels:
        N, _, H, W = lbl.shape
        e = proto[lbl.long().reshape(N, H, W)].permute(0, 3, 1, 2)
        e = e + noise * torch.randn(N, C, H, W, generator=g)
        out.append(l2_normalize(e).contiguous())
    return out
