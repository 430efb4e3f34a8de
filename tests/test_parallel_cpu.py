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
