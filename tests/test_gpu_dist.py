"""GPU, 2 ranks, NCCL: a data-parallel step of the Swin head (one clip per rank, dist.SegmentedStep: backward in
segments, one gradient bucket all-reduce per segment on a side stream) equals the 1-rank step on the concatenated
batch -- SURVEY.md section 7 "distributed" row.  Also the cross-rank key all-gather of the contrastive loss (C3) against
the oracle's concatenated generalisation.  Skipped with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import swin_oracle as so
        from stswincl_b200 import dist as sdist, swin
        dim, res, heads = 128, (16, 24), 2
        params = so.make_layer_params(dim, res, heads, seed=7)
        x = so.make_features(8, world, 4, dim, res[0], res[1]).to(dev)
        w1 = (so.make_features(9, world, 4, dim, res[0], res[1]) - 0.4).to(dev)
        w2 = (so.make_features(10, world, 4, 2 * dim, res[0] // 2, res[1] // 2) - 0.4).to(dev)
        ok, worst = True, 0.0
        for use_graph, wire in ((False, torch.float32), (True, torch.float32), (True, torch.bfloat16)):
            full = swin.SwinTransformerLayerv5(dim=dim, input_resolution=res, num_heads=heads)
            full.load_state_dict(params, strict=True)
            full = full.to(dev)
            y1, y2 = full(x)
            (((y1 * w1).sum() + (y2 * w2).sum()) / world).backward()            # the whole batch on one rank
            want = {n: p.grad.clone() for n, p in full.named_parameters()}
            mine = swin.SwinTransformerLayerv5(dim=dim, input_resolution=res, num_heads=heads)
            mine.load_state_dict(params, strict=True)
            mine = mine.to(dev)
            opt = torch.optim.SGD(mine.parameters(), lr=0.0)
            r = slice(rank, rank + 1)
            stepper = sdist.SegmentedStep(mine, opt, lambda y: (y[0] * w1[r]).sum() + (y[1] * w2[r]).sum(), segments=4,
                                          wire_dtype=wire, use_graph=use_graph)
            for _ in range(2):
                stepper.step(x[r].contiguous())
            torch.cuda.synchronize()
            tol = 2e-2 if wire == torch.bfloat16 else 5e-3
            for n, p in mine.named_parameters():
                e = rel_err(stepper._view_of[id(p)].float(), want[n])
                worst = max(worst, e)
                ok = ok and e < tol
            ok = ok and (stepper.captured == use_graph) and len(stepper._buckets) == 4
        ret[rank] = (bool(ok), worst)
    finally:
        dist.destroy_process_group()


def test_two_rank_swin_step_equals_one_rank_step_on_the_concatenated_batch():
    _need_two()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, 29700 + os.getpid() % 200, ret), nprocs=2, join=True)
    assert all(v[0] for v in dict(ret).values()) and len(ret) == 2, dict(ret)


def _c3_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import loss_oracle as lo
        from stswincl_b200 import contrast
        N, C, H, W, K = 2, 64, 8, 16, 12
        per_rank = []
        for r in range(world):                     # every rank builds every rank's data (the oracle needs all of it)
            labels = lo.make_label_maps(200 + r, 6, N, H, W, K)
            emb = lo.make_embeddings(300 + r, labels + labels[:2], C, K)
            per_rank.append((labels, emb))
        labels, emb = per_rank[rank]
        # oracle generalisation: the other ranks' four shared key sets are extra key sets of both queries
        extra_k = [e.to(torch.bfloat16).float() for r in range(world) if r != rank for e in per_rank[r][1][2:6]]
        extra_l = [l for r in range(world) if r != rank for l in per_rank[r][0][2:6]]
        bf = lambda t: t.to(torch.bfloat16).float()
        q1 = emb[6].clone().requires_grad_(True)
        q2 = emb[7].clone().requires_grad_(True)
        ref = (lo.regression_loss(q1, [bf(emb[1]), *[bf(e) for e in emb[2:6]], *extra_k], labels[0], [labels[1], *labels[2:6], *extra_l], K)
               + lo.regression_loss(q2, [bf(emb[0]), *[bf(e) for e in emb[2:6]], *extra_k], labels[1], [labels[0], *labels[2:6], *extra_l], K))
        ref.backward()
        c = lambda t: t.to(dev)
        p1, p2 = c(emb[6]).requires_grad_(True), c(emb[7]).requires_grad_(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loss = contrast.consistency_loss_tail(p1, p2, *[c(e) for e in emb[:6]], *[c(l) for l in labels], K, cross_rank_negatives=True)
        loss.backward()
        torch.cuda.synchronize()
        ok = abs(float(loss) - float(ref)) < 3e-3 * abs(float(ref))
        ok = ok and rel_err(p1.grad.cpu(), q1.grad) < 2e-2 and rel_err(p2.grad.cpu(), q2.grad) < 2e-2
        ret[rank] = (bool(ok), float(loss), float(ref))
    finally:
        dist.destroy_process_group()


def test_cross_rank_key_all_gather_matches_the_concatenated_oracle():
    """C3 on NCCL: every rank's loss sees its own five key sets plus the other ranks' four shared sets."""
    _need_two()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_c3_worker, args=(2, 29900 + os.getpid() % 200, ret), nprocs=2, join=True)
    assert all(v[0] for v in dict(ret).values()) and len(ret) == 2, dict(ret)


def test_data_parallel_two_replica_threads():
    """nn.DataParallel (seg18/train_swin.py:131-135): two replica threads call the kernels concurrently on two devices;
    output and parameter gradients equal the single-device run; the caller's current device is left alone."""
    _need_two()
    from oracle import swin_oracle as so
    from stswincl_b200 import swin
    dim, res, heads, ws, shift = 128, (16, 24), 2, 8, 4
    params = so.make_block_params(dim, res, heads, ws, shift, seed=21)
    x = so.make_features(22, 4, 2, res[0] * res[1], dim)
    w = (so.make_features(23, 4, 2, res[0] * res[1], dim) - 0.4).to(torch.bfloat16).float()
    single = swin.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift)
    single.load_state_dict(params, strict=True)
    single = single.cuda(0)
    y0 = single(x.cuda(0))
    (y0 * w.cuda(0)).sum().backward()
    want = {n: p.grad.clone() for n, p in single.named_parameters()}
    multi = swin.SwinTransformerBlock(dim, res, heads, window_size=ws, shift_size=shift)
    multi.load_state_dict(params, strict=True)
    dp = torch.nn.DataParallel(multi.cuda(0), device_ids=[0, 1])
    for _ in range(3):                                   # repeated: a race would not show on every run
        multi.zero_grad()
        y = dp(x.cuda(0))
        (y * w.cuda(0)).sum().backward()
        torch.cuda.synchronize()
        assert torch.cuda.current_device() == 0
        assert rel_err(y, y0) < 1e-3
        for n, p in multi.named_parameters():
            assert rel_err(p.grad, want[n]) < 5e-3, n


def test_segmented_step_single_gpu_equals_plain_step():
    """dist.SegmentedStep on one GPU (no collective): forward + four backward segments replayed from CUDA graphs, gradients
    handed to FusedAdam through the flat buckets -- equals the plain eager forward / backward / FusedAdam step."""
    from oracle import swin_oracle as so
    from stswincl_b200 import dist as sdist, optim, swin
    dim, res, heads = 128, (16, 24), 2
    params = so.make_layer_params(dim, res, heads, seed=17)
    x = so.make_features(18, 2, 4, dim, res[0], res[1]).cuda().to(torch.bfloat16)
    w1 = (so.make_features(19, 2, 4, dim, res[0], res[1]) - 0.4).cuda().to(torch.bfloat16)
    w2 = (so.make_features(20, 2, 4, 2 * dim, res[0] // 2, res[1] // 2) - 0.4).cuda().to(torch.bfloat16)
    loss_fn = lambda y: torch.sum(y[0] * w1, dtype=torch.float32) + torch.sum(y[1] * w2, dtype=torch.float32)

    def fresh():
        m = swin.SwinTransformerLayerv5(dim=dim, input_resolution=res, num_heads=heads)
        m.load_state_dict(params, strict=True)
        m = m.cuda()
        return m, optim.FusedAdam(m.parameters(), lr=1e-3)

    plain, o_plain = fresh()
    losses_plain = []
    for _ in range(3):
        o_plain.zero_grad(set_to_none=True)
        loss = loss_fn(plain(x))
        loss.backward()
        o_plain.step()
        losses_plain.append(float(loss))
    seg, o_seg = fresh()
    stepper = sdist.SegmentedStep(seg, o_seg, loss_fn, segments=4, wire_dtype=torch.float32, use_graph=True)
    losses_seg = [float(stepper.step(x)) for _ in range(3)]
    torch.cuda.synchronize()
    assert stepper.captured and len(stepper._buckets) == 4
    # first loss identical (same kernels, same weights); later ones after identical-math Adam steps (split-K atomics reorder sums)
    assert abs(losses_seg[0] - losses_plain[0]) <= 1e-5 * abs(losses_plain[0])
    assert all(abs(a - b) < 5e-3 * abs(b) for a, b in zip(losses_seg, losses_plain)), (losses_seg, losses_plain)
    for (n, p), q in zip(plain.named_parameters(), seg.parameters()):
        assert rel_err(q, p) < 2e-2, n
        if p.dim() >= 2:                         # the bf16 shadows follow the fp32 weights in both runs
            assert torch.equal(optim.bf16_weight(q), q.detach().to(torch.bfloat16))
