"""GPU: fused pixel contrastive loss against the reference's goldens and the CPU oracle.
Tolerance 2e-2 relative (bf16 embeddings on the tensor cores); masks / label resizing exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _g():
    return np.load(os.path.join(GOLDEN, "loss_cases.npz"))


@pytest.mark.parametrize("idx", range(5))
def test_regression_loss_vs_reference_golden(idx):
    from oracle import make_goldens as mg
    from stswincl_b200 import contrast
    case = mg.LOSS_CASES[idx]
    tag, N, C, H, W, K, special = case
    g = _g()
    labels, emb = mg.loss_case_inputs(*case)
    q = emb[0].cuda().requires_grad_(True)
    loss = contrast.regression_loss(q, *[e.cuda() for e in emb[1:]], *[l.cuda() for l in labels], K)
    loss.backward()
    torch.cuda.synchronize()
    ref = float(g[f"{tag}_loss"])
    assert abs(float(loss) - ref) < TOL * abs(ref)
    assert rel_err(q.grad.cpu(), g[f"{tag}_dq"]) < TOL


def test_masks_and_label_resize_exact():
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    g = _g()
    l1, l2 = lo.make_label_maps(73, 2, 2, 4, 6, 5, coarse=(2, 3))
    assert np.array_equal(contrast.posMask(l1.cuda(), l2.cuda(), 5).cpu().numpy(), g["pos_small"].astype(np.float32))
    assert np.array_equal(contrast.negMask(l1.cuda(), l2.cuda(), 5).cpu().numpy(), g["neg_small"].astype(np.float32))
    big = lo.make_label_maps(74, 1, 1, 256, 448, 12, coarse=(16, 28))[0]
    assert np.array_equal(contrast.downsample_labels(big.cuda(), 32, 56).cpu().numpy(), g["down_32x56"].astype(np.float32))
    odd = lo.make_label_maps(75, 1, 1, 100, 150, 12, coarse=(10, 15))[0]
    assert np.array_equal(contrast.downsample_labels(odd.cuda(), 24, 40).cpu().numpy(), g["down_odd_24x40"].astype(np.float32))


def test_consistency_tail_vs_reference_golden():
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    g = _g()
    N, C, H, W, K = 2, 64, 8, 14, 12
    full = lo.make_label_maps(76, 6, N, 64, 112, K, coarse=(4, 7))
    ds = [torch.nn.functional.interpolate(m, size=[H, W], mode="nearest") for m in full]
    emb = lo.make_embeddings(77, ds, C, K)
    pred_1, pred_2 = emb[0], lo.make_embeddings(78, ds[1:2], C, K)[0]
    proj_1, proj_2 = lo.make_embeddings(79, ds[0:1], C, K)[0], emb[1]
    c = lambda t: t.cuda()
    tail = contrast.consistency_loss_tail(c(pred_1), c(pred_2), c(proj_1), c(proj_2), c(emb[2]), c(emb[3]), c(emb[4]), c(emb[5]),
                                          *[c(m) for m in full], K)
    assert abs(float(tail) - float(g["tail_loss"])) < TOL * abs(float(g["tail_loss"]))


@pytest.mark.parametrize("N,C,H,W,K", [(2, 256, 32, 56, 12), (1, 128, 16, 28, 9)])
def test_full_size_loss_vs_oracle(N, C, H, W, K):
    """The reference's pre-training geometry: 256 channels, 32x56 = 1792 pixels."""
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    labels = lo.make_label_maps(81, 6, N, H, W, K, coarse=(4, 7))
    emb = lo.make_embeddings(82, labels, C, K)
    q_ref = emb[0].to(torch.bfloat16).float().requires_grad_(True)
    ref = lo.regression_loss(q_ref, [e.to(torch.bfloat16).float() for e in emb[1:]], labels[0], labels[1:], K)
    ref.backward()
    q = emb[0].cuda().requires_grad_(True)
    loss = contrast.regression_loss(q, *[e.cuda() for e in emb[1:]], *[l.cuda() for l in labels], K)
    (3.0 * loss).backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 2e-3 * abs(float(ref))
    assert rel_err(q.grad.cpu(), 3.0 * q_ref.grad) < TOL


def test_fused_normalize_and_extra_key_sets():
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    N, C, H, W, K = 2, 64, 8, 14, 12
    labels = lo.make_label_maps(91, 8, N, H, W, K)
    emb = lo.make_embeddings(92, labels, C, K)
    gen = torch.Generator().manual_seed(93)
    raw = [e * (0.5 + 2.0 * torch.rand(N, 1, H, W, generator=gen)) for e in emb]      # un-normalised inputs
    q_ref = raw[0].clone().requires_grad_(True)
    ref = lo.regression_loss(lo.l2_normalize(q_ref), [lo.l2_normalize(r) for r in raw[1:]], labels[0], labels[1:], K)
    ref.backward()
    q = raw[0].cuda().requires_grad_(True)
    loss = contrast.pixel_contrast_loss(q, [r.cuda() for r in raw[1:]], labels[0].cuda(), [l.cuda() for l in labels[1:]], K,
                                        normalize=True)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < TOL * abs(float(ref))
    assert rel_err(q.grad.cpu(), q_ref.grad) < TOL


def test_thirteen_key_sets_chunked_launches():
    """More than 8 key sets (e.g. key sets gathered from other ranks, SURVEY C3 -- an extension of the
    reference, checked against the oracle's own generalisation) run as several 8-set launches."""
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    N, C, H, W, K = 1, 64, 8, 16, 12
    labels = lo.make_label_maps(101, 14, N, H, W, K)
    emb = lo.make_embeddings(102, labels, C, K)
    q_ref = emb[0].to(torch.bfloat16).float().requires_grad_(True)
    ref = lo.regression_loss(q_ref, [e.to(torch.bfloat16).float() for e in emb[1:]], labels[0], labels[1:], K)
    ref.backward()
    q = emb[0].cuda().requires_grad_(True)
    loss = contrast.pixel_contrast_loss(q, [e.cuda() for e in emb[1:]], labels[0].cuda(), [l.cuda() for l in labels[1:]], K)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 2e-3 * abs(float(ref))
    assert rel_err(q.grad.cpu(), q_ref.grad) < TOL


def test_cross_rank_negatives_single_rank_is_reference_loss():
    """C3 extension: with world_size 1 the gathered form must equal the reference tail."""
    import torch.distributed as dist
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    g = _g()
    N, C, H, W, K = 2, 64, 8, 14, 12
    full = lo.make_label_maps(76, 6, N, 64, 112, K, coarse=(4, 7))
    ds = [torch.nn.functional.interpolate(m, size=[H, W], mode="nearest") for m in full]
    emb = lo.make_embeddings(77, ds, C, K)
    pred_1, pred_2 = emb[0], lo.make_embeddings(78, ds[1:2], C, K)[0]
    proj_1, proj_2 = lo.make_embeddings(79, ds[0:1], C, K)[0], emb[1]
    c = lambda t: t.cuda()
    created = False
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1)
        created = True
    try:
        tail = contrast.consistency_loss_tail(c(pred_1), c(pred_2), c(proj_1), c(proj_2), c(emb[2]), c(emb[3]), c(emb[4]),
                                              c(emb[5]), *[c(m) for m in full], K, cross_rank_negatives=True)
    finally:
        if created:
            dist.destroy_process_group()
    assert abs(float(tail) - float(g["tail_loss"])) < TOL * abs(float(g["tail_loss"]))


def test_out_of_range_label_raises_and_cpu_raises():
    from oracle import make_goldens as mg
    from stswincl_b200 import contrast
    from stswincl_b200._lib import StswinError
    labels, emb = mg.loss_case_inputs("l0", 1, 64, 8, 14, 12, None)
    bad = labels[2].clone(); bad[0, 0, 0, 0] = 12.0
    args = [e[:1].cuda() for e in emb] + [labels[0][:1].cuda(), labels[1][:1].cuda(), bad[:1].cuda()] + [l[:1].cuda() for l in labels[3:]]
    with pytest.raises(RuntimeError):                 # the reference's F.one_hot behaviour, on request (one host read)
        contrast.regression_loss(*args, 12, validate_labels=True)
    # default: checked on the device, no host synchronisation -- the loss is NaN
    assert torch.isnan(contrast.regression_loss(*args, 12)).item()
    with pytest.raises(StswinError):
        contrast.regression_loss(*[e[:1] for e in emb], *[l[:1] for l in labels], 12)
    with pytest.raises(StswinError):                  # labels live in one byte
        contrast.regression_loss(*args, 255)


@pytest.mark.parametrize("N,H,W,Hs,Ws,K,coarse,dtype", [
    (2, 32, 56, 256, 448, 12, (4, 7), torch.float32), (3, 8, 14, 8, 14, 5, (8, 14), torch.uint8),
    (1, 28, 28, 100, 150, 26, (5, 6), torch.int64), (2, 24, 40, 24, 40, 254, (24, 40), torch.float32)])
def test_label_tables_exact(N, H, W, Hs, Ws, K, coarse, dtype):
    """Device label pass (nearest down-sampling, .long(), stable counting sort, group labels, histograms) against
    numpy -- integer work, exact."""
    from oracle import index_oracle as ix, loss_oracle as lo
    from stswincl_b200 import contrast
    maps = lo.make_label_maps(111, 3, N, Hs, Ws, K, coarse=coarse)
    nat, srt, glab, perm, hist, ctl = [t.cpu().numpy() for t in contrast.label_tables([m.to(dtype).cuda() for m in maps], H, W, K)]
    HW, HWp = H * W, (H * W + 255) // 256 * 256
    assert ctl[0] == 0
    for li, m in enumerate(maps):
        ref = ix.downsample_labels(m.numpy(), H, W).reshape(N, HW).astype(np.int64)
        for n in range(N):
            assert np.array_equal(nat[li, n, :HW], ref[n]) and (nat[li, n, HW:] == 255).all()
            order = np.argsort(ref[n], kind="stable")
            assert np.array_equal(srt[li, n, :HW], ref[n][order]) and (srt[li, n, HW:] == 255).all()
            inv = np.empty(HW, dtype=np.int64); inv[order] = np.arange(HW)
            assert np.array_equal(perm[li, n].astype(np.int64) & 0xffff, inv)
            assert np.array_equal(hist[li, n], np.bincount(ref[n], minlength=256))
            groups = srt[li, n].reshape(HWp // 32, 32)
            want = np.where((groups == groups[:, :1]).all(1), groups[:, 0], 254)
            assert np.array_equal(glab[li, n, :HWp // 32], want)


@pytest.mark.parametrize("N,C,H,W,K,coarse", [(2, 256, 32, 56, 12, (4, 7)), (1, 64, 28, 28, 26, (28, 28)), (3, 128, 8, 16, 9, (2, 2))])
def test_symmetric_tail_fused_normalize_vs_oracle(N, C, H, W, K, coarse):
    """ConsistencyLoss tail (:584-597) as ONE fused step with F.normalize inside, against the oracle: full-resolution
    label maps, raw (un-normalised) embeddings, gradients to both query maps.  coarse = (H, W) gives per-pixel random
    labels: nearly every 32-key group is mixed (the per-element path of both kernels)."""
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    full = lo.make_label_maps(121, 6, N, 8 * H, 8 * W, K, coarse=coarse)
    ds = [torch.nn.functional.interpolate(m, size=[H, W], mode="nearest") for m in full]
    gen = torch.Generator().manual_seed(122)
    raw = [e * (0.5 + 2.0 * torch.rand(N, 1, H, W, generator=gen)) for e in lo.make_embeddings(123, ds + ds[:2], C, K)]
    pred = [raw[6].clone().requires_grad_(True), raw[7].clone().requires_grad_(True)]
    nrm = lo.l2_normalize
    bf = lambda t: nrm(t).to(torch.bfloat16).float()
    ref = lo.consistency_tail(nrm(pred[0]), nrm(pred[1]), bf(raw[0]), bf(raw[1]), [bf(r) for r in raw[2:6]], full[0], full[1], full[2:], K)
    (2.0 * ref).backward()
    c = lambda t: t.cuda()
    p1, p2 = c(raw[6]).requires_grad_(True), c(raw[7]).requires_grad_(True)
    loss = contrast.consistency_loss_tail(p1, p2, *[c(r) for r in raw[:6]], *[c(m) for m in full], K, normalize=True)
    (2.0 * loss).backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 3e-3 * abs(float(ref))
    assert rel_err(p1.grad.cpu(), pred[0].grad) < TOL
    assert rel_err(p2.grad.cpu(), pred[1].grad) < TOL


def test_tail_is_cuda_graph_capturable_and_deterministic():
    """No host synchronisation anywhere in the step: capture forward + backward once, replay on new data."""
    from oracle import loss_oracle as lo
    from stswincl_b200 import contrast
    N, C, H, W, K = 2, 128, 16, 28, 12
    def data(seed):
        labels = lo.make_label_maps(seed, 6, N, H, W, K)
        emb = lo.make_embeddings(seed + 1, labels + labels[:2], C, K)
        return [e.cuda() for e in emb], [l.cuda() for l in labels]
    emb, labels = data(131)
    static_e = [e.clone() for e in emb]
    static_l = [l.clone() for l in labels]
    q1, q2 = static_e[6].requires_grad_(True), static_e[7].requires_grad_(True)
    def step():
        q1.grad = None; q2.grad = None
        loss = contrast.consistency_loss_tail(q1, q2, *static_e[:6], *static_l, K)
        loss.backward()
        return loss
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    for seed in (131, 141):
        emb, labels = data(seed)
        with torch.no_grad():
            for d, src in zip(static_e, emb): d.copy_(src)
            for d, src in zip(static_l, labels): d.copy_(src)
        g.replay()
        got, g1 = float(out), q1.grad.clone()
        g.replay()
        assert float(out) == got                                   # bit-reproducible loss (the gradient sums use fp32 atomics)
        assert rel_err(q1.grad, g1) < 1e-5
        e1, e2 = emb[6].clone().requires_grad_(True), emb[7].clone().requires_grad_(True)
        ref = contrast.consistency_loss_tail(e1, e2, *emb[:6], *labels, K)
        ref.backward()
        assert float(ref) == got
        assert rel_err(g1, e1.grad) < 1e-5


class _StubSegmentor(torch.nn.Module):
    """Tiny stand-in for the caller's segmentor (TswinPlusv5: resnet -> swin -> aspp + three projections)."""

    def __init__(self, head):
        super().__init__()
        self.resnet = torch.nn.Conv2d(3, 128, 8, stride=8)
        self.swin = head
        self.aspp = torch.nn.Conv2d(256, 24, 1)
        self.project1, self.project2, self.project3 = torch.nn.Conv2d(128, 8, 1), torch.nn.Conv2d(128, 8, 1), torch.nn.Conv2d(256, 8, 1)


def test_pixpro_wrapper_and_consistency_loss_module():
    """stswincl_b200.pixpro.PixPro around caller-built encoders + contrast.ConsistencyLoss(args, pixpro)
    (PixPro_swin_v5.py:138-597): head swap with strict weight load, reference state_dict key names, no-grad key passes,
    cosine-scheduled multi-tensor EMA, and the fused loss equal to the oracle on the returned embeddings."""
    import math
    import types
    from oracle import loss_oracle as lo, tswin_oracle as to
    from stswincl_b200 import contrast, pixpro, swin
    torch.manual_seed(0)
    shapes = {k: (tuple(v.shape), isinstance(v, torch.nn.Parameter), v.dtype) for k, v in
              list(swin.SwinTransformerLayerv5(128, (16, 24), 2).named_parameters()) + list(swin.SwinTransformerLayerv5(128, (16, 24), 2).named_buffers())}
    def make_segmentor():      # the caller hands over a segmentor whose head is NOT ours (here: the CPU oracle module)
        head = to.OracleSwin(shapes, dim=128, input_resolution=(16, 24), num_heads=2)
        ref = swin.SwinTransformerLayerv5(128, (16, 24), 2)
        head.load_state_dict(ref.state_dict(), strict=True)
        return _StubSegmentor(head).cuda()
    args = types.SimpleNamespace(pixpro_momentum=0.9, pixpro_transform_layer=1, num_instances=64, batch_size=2, epochs=4,
                                 start_epoch=1, data="endo18", pixpro_pos_ratio=0.7)
    pp = pixpro.PixPro(args, make_segmentor, lambda: torch.nn.Conv2d(48, 64, 1).cuda(), fuse_normalize=True).cuda()
    assert isinstance(pp.encoder_2, swin.SwinTransformerLayerv5) and isinstance(pp.encoder_k_2, swin.SwinTransformerLayerv5)
    keys = set(pp.state_dict().keys())
    assert {"encoder_2.layers.0.0.attn.qkv.weight", "encoder_k_2.downsample.reduction.weight", "value_transform.weight",
            "encoder_1.weight", "proj_k_3.bias", "projector_k.weight"} <= keys
    assert all(not p.requires_grad for p in pp.encoder_k_2.parameters()) and all(p.requires_grad for p in pp.encoder_2.parameters())
    with torch.no_grad():                                   # make the query encoder differ from the key encoder
        for p in pp.encoder_2.parameters():
            p.add_(0.01 * torch.randn_like(p))
    before = [p.clone() for p in pp.encoder_k_2.parameters()]
    model = contrast.ConsistencyLoss(args, pp)
    assert model.class_num == 12
    gen = torch.Generator().manual_seed(1)
    ims = [torch.rand(2, 4, 3, 128, 192, generator=gen).cuda() for _ in range(6)]
    masks = [m.cuda() for m in lo.make_label_maps(2, 6, 2, 128, 192, 12, coarse=(2, 3))]
    loss = model(*ims, *masks)
    loss.backward()
    torch.cuda.synchronize()
    m = 1. - (1. - 0.9) * (math.cos(math.pi * 0 / pp.K) + 1) / 2.
    assert pp.k == 1
    for b, kp, qp in zip(before, pp.encoder_k_2.parameters(), pp.encoder_2.parameters()):
        assert torch.equal(kp, b * m + qp.detach() * (1. - m))                      # bit-exact EMA
    assert pp.encoder_2.layers[0][0].attn.qkv.weight.grad is not None and pp.encoder_k_2.layers[0][0].attn.qkv.weight.grad is None
    with torch.no_grad():
        outs = pp(*ims)
    assert len(outs) == 8 and outs[0].shape == (2, 64, 16, 24)
    pp.k -= 1                                               # the second call advanced the schedule; the loss check below only needs embeddings
    nrm = lo.l2_normalize
    emb = [nrm(o.float().cpu()) for o in outs]
    ref = lo.consistency_tail(emb[0], emb[1], emb[2], emb[3], emb[4:8], masks[0].cpu(), masks[1].cpu(), [m_.cpu() for m_ in masks[2:]], 12)
    with torch.no_grad():
        again = contrast.consistency_loss_tail(*outs, *masks, 12, normalize=True)
    assert abs(float(again) - float(ref)) < 5e-3 * abs(float(ref))
