"""CPU, world_size 2, gloo: the host-side data-parallel logic (clip sharding, gradient averaging,
key-set gathering) -- the N>1 path of bench.py minus the kernels."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stswincl_b200 import dist as sdist


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
        data = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
        target = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
        # single-process reference on the full batch
        ref = [p.clone() for p in model.parameters()]
        loss = ((model(data) - target) ** 2).mean()
        full = torch.autograd.grad(loss, list(model.parameters()))
        # this rank's shard
        idx = sdist.shard_indices(8, rank, world)
        loss = ((model(data[idx]) - target[idx]) ** 2).mean()
        loss.backward()
        n = sdist.average_gradients(model.parameters(), bucket_bytes=64)      # tiny buckets: several collectives
        ok = all(torch.allclose(p.grad, g, atol=1e-6) for p, g in zip(model.parameters(), full)) and n >= 2
        # bf16 on the wire: same up to rounding
        model.zero_grad()
        ((model(data[idx]) - target[idx]) ** 2).mean().backward()
        sdist.average_gradients(model.parameters(), comm_dtype=torch.bfloat16)
        ok = ok and all(torch.allclose(p.grad, g, atol=2e-2, rtol=2e-2) for p, g in zip(model.parameters(), full))
        # one flat buffer, one collective (bench.py --dp flat)
        model.zero_grad()
        ((model(data[idx]) - target[idx]) ** 2).mean().backward()
        flat = sdist.FlatGradients(model.parameters())
        flat.gather(); flat.all_reduce(); flat.bind()
        ok = ok and all(torch.allclose(p.grad, g, atol=1e-6) for p, g in zip(model.parameters(), full))
        ok = ok and all(p.grad.data_ptr() == v.data_ptr() and v.data_ptr() % 16 == 0 for p, v in zip(flat.params, flat.views))
        with torch.no_grad():
            list(model.parameters())[0].add_(float(rank))          # make the replicas differ, then re-synchronise
        sdist.broadcast_parameters(model.parameters())
        ok = ok and torch.equal(list(model.parameters())[0], ref[0])
        # gather: own copy first, then the others
        t = torch.full((2, 3), float(rank))
        got = sdist.gather_key_sets([t, t + 10])
        ok = ok and len(got) == 2 * world and float(got[0][0, 0]) == rank and float(got[1][0, 0]) == 1 - rank \
            and float(got[2][0, 0]) == rank + 10
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_matches_full_batch():
    world = 2
    port = 29600 + os.getpid() % 200
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


class _Staged(torch.nn.Module):
    """Toy model with the ``stages(x)`` contract of SwinTransformerLayerv5 (a chain of (fn, parameters))."""

    def __init__(self):
        super().__init__()
        self.a, self.b, self.c = torch.nn.Linear(6, 8), torch.nn.Linear(8, 8), torch.nn.Linear(8, 3)

    def stages(self, like):
        return [(lambda x: (torch.tanh(self.a(x)),), list(self.a.parameters())),
                (lambda h: (h, torch.tanh(self.b(h))), list(self.b.parameters())),
                (lambda h, g: (self.c(h + g),), list(self.c.parameters()))]

    def forward(self, x):
        state = (x,)
        for fn, _ in self.stages(x):
            state = fn(*state)
        return state


def _seg_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
        target = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
        ok = True
        for segments in (1, 2, 3):
            torch.manual_seed(0)
            full, mine = _Staged(), _Staged()
            mine.load_state_dict(full.state_dict())
            o_full = torch.optim.SGD(full.parameters(), lr=0.1, momentum=0.9)
            o_mine = torch.optim.SGD(mine.parameters(), lr=0.1, momentum=0.9)
            idx = sdist.shard_indices(8, rank, world)
            stepper = sdist.SegmentedStep(mine, o_mine, lambda y: ((y[0] - target[idx]) ** 2).mean(), segments=segments,
                                          wire_dtype=torch.float32, use_graph=False)
            for _ in range(3):
                o_full.zero_grad()
                ((full(data)[0] - target) ** 2).mean().backward()
                o_full.step()
                stepper.step(data[idx])
            ok = ok and all(torch.allclose(p, q, atol=1e-6) for p, q in zip(full.parameters(), mine.parameters()))
            ok = ok and len(stepper._buckets) == segments
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_segmented_step_two_ranks_equal_the_full_batch_step():
    """dist.SegmentedStep (backward cut at stage boundaries, one bucket all-reduce per segment): two ranks on half the
    batch each reproduce the single-process step on the whole batch, for 1 / 2 / 3 segments."""
    world = 2
    port = 29800 + os.getpid() % 200
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_seg_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_indices_cover_and_pad():
    for n, world in [(8, 2), (7, 2), (5, 4), (3, 8)]:
        shards = [sdist.shard_indices(n, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1                      # equal length (wrap-padded)
        assert set(i for s in shards for i in s) == set(range(n))      # every sample owned
    assert sdist.shard_indices(7, 1, 2) == [1, 3, 5, 0]
    assert sdist.scaled_lr(1.0, 4, 8) == 4 * 8 / 256
