"""World-size-2 gloo tests (CPU) of the data-parallel plumbing: sharding, the single flat all-reduce with the
loss scalars riding in its tail, and the equivalence  mean-gradient of the global batch == all-reduced sum of
per-rank unnormalised gradients * si_unscale."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from avsi_b200 import parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        B, T, F, P = 6, 5, 7, 40                     # global batch, frames, bins, parameters
        rng = np.random.default_rng(0)              # identical on both ranks
        W = torch.tensor(rng.standard_normal((F, P // F + 1))[:, :1].repeat(P, 1).T[:P].copy())   # dummy, unused
        x = torch.tensor(rng.standard_normal((B, T, F)))
        target = torch.tensor(rng.standard_normal((B, T, F)))
        theta = torch.tensor(rng.standard_normal(F), requires_grad=True)
        mask = torch.tensor((rng.uniform(size=(B, T, 1)) > 0.3).astype(np.float64)).expand(B, T, F)
        # reference: one device, whole batch, loss = mean |target - x*theta|
        ref_loss = (target - x * theta).abs().mean()
        ref_grad, = torch.autograd.grad(ref_loss, theta)
        # data parallel: each rank computes the UNNORMALISED sum over its shard
        lo, hi = parallel.shard_batch(B, rank, world)
        th = theta.detach().clone().requires_grad_(True)
        d = (target[lo:hi] - x[lo:hi] * th).abs()
        g, = torch.autograd.grad(d.sum(), th)
        flat = torch.zeros(F + parallel.LOSS_TAIL, dtype=torch.float64)
        flat[:F] = g
        m = mask[lo:hi]
        sums = torch.stack([(d * (1 - m)).sum(), (1 - m).sum(), (d * m).sum(), m.sum(), d.sum(),
                            torch.tensor(float(d.numel()), dtype=torch.float64)]).detach()
        parallel.pack_loss_tail(flat, F, sums)
        parallel.all_reduce_flat(flat)
        grad = flat[:F] * parallel.si_unscale(B // world, world, T, F)
        hole, valid, func = parallel.global_losses(flat, F)
        dfull = (target - x * theta.detach()).abs()
        ok = (torch.allclose(grad, ref_grad, rtol=1e-12, atol=1e-14)
              and abs(float(func) - float(ref_loss)) < 1e-12
              and abs(float(hole) - float((dfull * (1 - mask)).sum() / (1 - mask).sum())) < 1e-12
              and abs(float(valid) - float((dfull * mask).sum() / mask.sum())) < 1e-12)
        files = ['s%03d' % i for i in range(11)]
        mine = parallel.shard_list(files, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        flat_all = sorted(sum(gathered, []))
        ok = ok and flat_all == sorted(files) and len(set(map(tuple, gathered))) == world
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_flat_allreduce_matches_single_device_gloo_world2():
    world = 2
    ctx = mp.get_context('spawn')
    mgr = ctx.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}


def test_shard_batch_rejects_ragged():
    from avsi_b200 import parallel
    assert parallel.shard_batch(8, 1, 4) == (2, 4)
    with pytest.raises(ValueError):
        parallel.shard_batch(10, 0, 4)


def _lockstep_worker(rank, world, port, out):
    """Ragged shards: rank 0 holds 5 full batches + a partial one, rank 1 holds 4 full batches.  Both must leave the
    loop after 4 steps, having issued the same number of collectives (the all-reduce inside the loop must not hang)."""
    import torch.distributed as dist
    from avsi_b200 import parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        bs = 3
        n_full = 5 if rank == 0 else 4

        def gen():
            for i in range(n_full):
                yield (np.arange(bs), 'full%d' % i)
            if rank == 0:
                yield (np.arange(bs - 1), 'partial')
        steps, total = 0, torch.zeros(1, dtype=torch.float64)
        for b in parallel.lockstep(gen(), bs, None, 'cpu'):
            assert len(b[0]) == bs
            g = torch.full((1,), float(rank + 1), dtype=torch.float64)
            dist.all_reduce(g)                        # stands for the gradient all-reduce of the step
            total += g
            steps += 1
        # a second epoch right after (validation pass): still in step
        steps2 = sum(1 for _ in parallel.lockstep(iter([(np.arange(bs), 'v')] * (2 + rank)), bs, None, 'cpu'))
        out[rank] = (steps, float(total), steps2)
    finally:
        dist.destroy_process_group()


def test_lockstep_drops_the_ragged_tail_on_every_rank_gloo_world2():
    world = 2
    ctx = mp.get_context('spawn')
    mgr = ctx.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_lockstep_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: (4, 12.0, 2), 1: (4, 12.0, 2)}


def test_lockstep_is_plain_iteration_without_a_process_group():
    from avsi_b200 import parallel
    got = list(parallel.lockstep(iter([(np.arange(3),), (np.arange(2),)]), 3))
    assert [len(b[0]) for b in got] == [3, 2]                 # single process: the partial batch is kept


def test_shard_list_keeps_a_rank_consistent_shuffle():
    import random
    from avsi_b200 import parallel
    files = ['f%02d' % i for i in range(10)]
    shuffled = list(files)
    random.Random(5).shuffle(shuffled)
    a, b = parallel.shard_list(shuffled, 0, 2, keep_order=True), parallel.shard_list(shuffled, 1, 2, keep_order=True)
    assert a == shuffled[0::2] and b == shuffled[1::2] and sorted(a + b) == files
    assert parallel.shard_list(shuffled, 0, 2) == files[0::2]
