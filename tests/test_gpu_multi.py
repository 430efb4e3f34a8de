"""Two-GPU NCCL parity with the REAL model: a data-parallel step on 2 ranks equals the
1-rank step on the concatenated batch -- for av-blstm (SI: global mean over B*T*F) and av-blstm-ssnn-ctc (MTL: the hole
normaliser sum(1-m) is a GLOBAL ratio, models.py:1947, all-reduced before the loss kernel) -- and a 2-rank train()
job on an odd number of TFRecord files ends cleanly on both ranks.  Skipped on boxes with fewer than 2 GPUs
(run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

pytestmark = pytest.mark.gpu


def _free_port():
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(name, batch, idx, T, pg, device, cfg_over):
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    cfg = synth.default_config(name, batch_size=len(idx), audio_len=batch['wav'].shape[1], ctc_loss=0.05, **cfg_over)
    cls, inp = models.MODEL_REGISTRY[name]
    video = av_sync.video_pipeline(batch['landmarks'][idx], T, batch['vmean'][idx], batch['vstd'][idx], device=device)
    if cls.MTL:
        m = cls(batch['seq_len'][idx], batch['lab_len'][idx], batch['wav'][idx], batch['mask'][idx], batch['labels'][idx],
                batch['mean'], batch['std'], 0.0, cfg, video_features=video, input=inp, device=device, process_group=pg)
    else:
        m = cls(batch['seq_len'][idx], batch['wav'][idx], batch['mask'][idx], batch['mean'], batch['std'], 0.0, cfg,
                video_features=video, input=inp, device=device, process_group=pg)
    m.assign_vars(init_canonical(m.engine.layout, seed=5, bias_scale=0.05))
    return m


def _dp_worker(rank, world, port, name, out):
    import torch.distributed as dist
    from avsi_b200 import synth
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        B, audio_len = 6, 11520                                  # T = 60 >= 2 * 24 + 1 CTC states
        batch = synth.make_batch(B, audio_len=audio_len, seed=31)
        # unequal hole sizes per rank so that a per-rank hole normaliser would differ from the global one
        batch['mask'][:] = 1.0
        for b in range(B):
            batch['mask'][b, 5:5 + 3 * (b + 1)] = 0.0
        T = batch['T']
        dev = 'cuda:%d' % rank
        lo, hi = rank * B // world, (rank + 1) * B // world
        res = {}
        for opt in ('sgd',):
            dp = _model(name, batch, np.arange(lo, hi), T, dist.group.WORLD, dev, dict(optimizer_type=opt, starter_learning_rate=0.05))
            g_dp = dp.canonical_gradients(reduce=True)
            loss_hole_dp = float(dp.loss_hole) if dp.MTL else None
            dp.feed(dropout_rate=0.0)
            dp.train_op()
            dp.feed(dropout_rate=0.0)
            dp.train_op()
            th_dp = dp.engine.theta.clone()
            if rank == 0:
                one = _model(name, batch, np.arange(B), T, None, dev, dict(optimizer_type=opt, starter_learning_rate=0.05))
                th0 = one.engine.theta.clone()
                g_one = one.canonical_gradients()
                one.feed(dropout_rate=0.0)
                one.train_op()
                one.feed(dropout_rate=0.0)
                one.train_op()
                ga = np.concatenate([g_dp[k].ravel() for k in sorted(g_one)])
                gb = np.concatenate([g_one[k].ravel() for k in sorted(g_one)])
                res['grad_rel'] = float(np.linalg.norm(ga - gb) / np.linalg.norm(gb))
                upd = (one.engine.theta - th0).double()
                res['update_rel'] = float(((th_dp - one.engine.theta).double().norm() / upd.norm()).item())
                res['update_norm'] = float(upd.norm().item())
            # every rank must hold the same weights after the step
            gathered = [torch.empty_like(th_dp) for _ in range(world)]
            dist.all_gather(gathered, th_dp)
            res['replicas_equal'] = bool(all(torch.equal(gathered[0], g) for g in gathered))
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', ['av-blstm', 'av-blstm-ssnn-ctc'])
def test_two_rank_step_equals_one_rank_step_on_the_concatenated_batch(name):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    res = dict(out)
    assert res[0]['replicas_equal'] and res[1]['replicas_equal']
    # same arithmetic, different summation order (per-rank partial sums, split-K atomics): far inside the fp16 budget
    assert res[0]['grad_rel'] < 5e-4, res[0]
    assert res[0]['update_norm'] > 0 and res[0]['update_rel'] < 2e-3, res[0]


def _train_worker(rank, world, port, cfg, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from avsi_b200 import training
        model = training.train(cfg)
        th = model.engine.theta
        gathered = [torch.empty_like(th) for _ in range(world)]
        dist.all_gather(gathered, th)
        out[rank] = (int(model.global_step), bool(all(torch.equal(gathered[0], g) for g in gathered)),
                     bool(torch.isfinite(th).all()))
    finally:
        dist.destroy_process_group()


def test_two_rank_train_job_with_ragged_shards(tmp_path):
    """11 training files over 2 ranks at 2 utterances per rank and step: rank 0 holds 6 files, rank 1 holds 5 -> 2 full
    lock-step steps per epoch, the tails (2 and 1 utterances) are dropped, nobody hangs, replicas stay identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from test_gpu_model import _write_dataset
    root = str(tmp_path / 'data')
    os.makedirs(root)
    audio_len = 11520
    _write_dataset(root, 11, 4, audio_len, seed=80)
    np.save(os.path.join(root, 'mean.npy'), np.full(257, 6.0))
    np.save(os.path.join(root, 'std.npy'), np.full(257, 2.0))
    exp = str(tmp_path / 'exp' / 'dp')
    cfg = str(tmp_path / 'blstm_ctc.config')
    with open(cfg, 'w') as f:
        f.write('root_folder = %s\nexp_folder = %s\nmodel = av-blstm-ssnn-ctc\naudio_feat_dim = 257\nvideo_feat_dim = 136\n'
                'audio_len = %d\nbatch_size = 4\nnet_dim = [250,250,250]\nstarter_learning_rate = 0.001\nmax_n_epochs = 2\n'
                'n_earlystop_epochs = 5\nlr_decay = 1.0\noptimizer_type = adam\nl2 = 0.0\ndropout_rate = 0.0\nctc_loss = 0.001\n'
                'num_asr_labels = 33\naudio_feat_mean = %s\naudio_feat_std = %s\n'
                % (root, exp, audio_len, os.path.join(root, 'mean.npy'), os.path.join(root, 'std.npy')))
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, cfg, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0, 'rank hung or failed'
    res = dict(out)
    assert res[0] == res[1] == (4, True, True), res
