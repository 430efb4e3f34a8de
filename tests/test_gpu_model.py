"""Model-level parity (B200): front end -> stacked BLSTM -> loss -> BPTT -> Adam against the float64
oracle on identical synthetic inputs and weights.  Tolerances are BASELINE.json's: spectra 1e-5
relative, BLSTM outputs and gradients 2e-3 relative L2."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 2e-3


def _oracle_inputs(batch, input_type):
    from oracle import stft as ostft
    from oracle import video as ovideo
    B, T = batch['wav'].shape[0], batch['T']
    vid = np.stack([ovideo.video_features(batch['landmarks'][b].astype(np.float64), T, batch['vmean'][b].astype(np.float64),
                                          batch['vstd'][b].astype(np.float64)) for b in range(B)])
    _, tsn, audio = ostft.frontend(batch['wav'].astype(np.float64), batch['mask'].astype(np.float64),
                                   batch['mean'].astype(np.float64), batch['std'].astype(np.float64), None)
    net_in = {'a': audio, 'v': vid, 'av': np.concatenate([audio, vid], 2)}[input_type]
    return tsn, net_in


def _build(model_name, B, audio_len, seed, seq_len=None, bias_scale=0.05, **cfg_kw):
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    batch = synth.make_batch(B, audio_len=audio_len, seed=seed)
    if seq_len is not None:
        batch['seq_len'] = np.asarray(seq_len, np.int32)
    cfg = synth.default_config(model_name, batch_size=B, audio_len=audio_len, **cfg_kw)
    cls, inp = models.MODEL_REGISTRY[model_name]
    video = av_sync.video_pipeline(batch['landmarks'], batch['T'], batch['vmean'], batch['vstd'])
    if cls.MTL:
        model = cls(batch['seq_len'], batch['lab_len'], batch['wav'], batch['mask'], batch['labels'], batch['mean'],
                    batch['std'], 0.0, cfg, video_features=video, input=inp)
    else:
        model = cls(batch['seq_len'], batch['wav'], batch['mask'], batch['mean'], batch['std'], 0.0, cfg,
                    video_features=video, input=inp)
    canon = init_canonical(model.engine.layout, seed=seed + 1, bias_scale=bias_scale)
    model.assign_vars(canon)
    return model, batch, canon, inp


def _check_grads(grads, ograds, what, tol=TOL):
    ga = np.concatenate([grads[k].ravel() for k in sorted(ograds)])
    gb = np.concatenate([ograds[k].ravel() for k in sorted(ograds)])
    worst = max((rel_l2(grads[k], ograds[k]), k) for k in ograds if np.linalg.norm(ograds[k]) > 0)
    assert rel_l2(ga, gb) < tol, '%s: grad rel-L2 %.3e, worst %s' % (what, rel_l2(ga, gb), worst)
    assert worst[0] < 2 * tol, '%s: worst per-variable grad rel-L2 %.3e at %s' % (what, worst[0], worst[1])


def _through_f16(a):
    return np.asarray(a, np.float64).astype(np.float16).astype(np.float64)


@pytest.mark.parametrize('model_name,B,audio_len', [('av-blstm', 3, 9600), ('a-blstm', 5, 4800), ('v-blstm', 2, 4800),
                                                    ('av-blstm', 20, 2400), ('av-blstm', 1, 100), ('a-blstm', 2, 193),
                                                    ('av-blstm', 230, 150)])
def test_si_forward_loss_gradients(model_name, B, audio_len):
    """Includes the degenerate shapes: one frame (T = 1: no recurrent step at all, shorter than the analysis window),
    T = 2, a single utterance, and T = 1 on the large-batch kernels."""
    from oracle import blstm as oblstm
    T = -(-audio_len // 192)
    seq = np.full(B, T)
    seq[-1] = max(1, T - 3)
    model, batch, canon, inp = _build(model_name, B, audio_len, seed=B, seq_len=seq)
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len']),
                                         canon, 3)
    assert rel_l2(model.target_spec_norm.cpu().numpy(), tsn) < 1e-5
    assert rel_l2(model.net_inputs.cpu().numpy(), net_in) < 1e-3            # fp16 copy of the network input
    assert rel_l2(model.inference.cpu().numpy(), outs['inference']) < TOL
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    for name in ('loss', 'loss_func', 'loss_hole', 'loss_valid'):
        got, want = float(getattr(model, name)), float(outs[name])
        if np.isnan(want):                           # 0 / 0: no reliable (or no masked) bin in the batch, as in the reference
            assert np.isnan(got), name
        else:
            assert abs(got - want) < TOL * abs(want), name
    _check_grads(model.canonical_gradients(), ograds, model_name)


@pytest.mark.parametrize('model_name,B,audio_len', [('av-blstm', 278, 2880), ('a-blstm', 257, 1920)])
def test_si_large_batch_tcgen05_path(model_name, B, audio_len):
    """B >= 225 selects the tcgen05 4-CTA-cluster recurrence kernels (forward + BPTT) and, with M*N*K large enough,
    the CTA-pair GEMM: same oracle, same tolerances; the last batch tile is ragged (278 = 2 * 128 + 22)."""
    from oracle import blstm as oblstm
    T = -(-audio_len // 192)
    seq = np.full(B, T)
    seq[3] = T - 2
    model, batch, canon, inp = _build(model_name, B, audio_len, seed=B, seq_len=seq)
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len']),
                                         canon, 3)
    assert rel_l2(model.target_spec_norm.cpu().numpy(), tsn) < 1e-5
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * abs(float(outs['loss']))
    _check_grads(model.canonical_gradients(), ograds, model_name + ' B=%d' % B)


def test_mtl_large_batch_tcgen05_path():
    from oracle import blstm as oblstm
    B, audio_len = 230, 11520        # T = 60 >= 2 * 24 + 1 label states; B > 224: tcgen05 recurrence
    model, batch, canon, inp = _build('av-blstm-ssnn-ctc', B, audio_len, seed=33, ctc_loss=0.05)
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads(
        'mtl', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len'], labels=batch['labels'],
                    lab_len=batch['lab_len']), canon, 3, ctc_weight=0.05)
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * float(outs['loss'])
    _check_grads(model.canonical_gradients(), ograds, 'mtl B=230')


def test_full_size_batch_consistent_with_small_batch_kernels():
    """BASELINE shape (3 s utterances, T = 250) at B = 256: utterances computed inside the big batch (tcgen05
    recurrence, CTA-pair GEMMs) agree with the same utterances run alone through the small-batch kernels
    (mma.sync recurrence, one-tile GEMM) -- a size-independent property, no oracle needed; and the summed loss
    sums are additive over a split of the batch."""
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    B = 256
    batch = synth.make_batch(B, audio_len=48000, seed=77)
    cfg = synth.default_config('av-blstm', batch_size=B, audio_len=48000)

    def run(idx):
        sub = {k: batch[k][idx] for k in ('wav', 'mask', 'landmarks', 'vmean', 'vstd', 'seq_len')}
        video = av_sync.video_pipeline(sub['landmarks'], batch['T'], sub['vmean'], sub['vstd'])
        m = models.StackedBLSTMModel(sub['seq_len'], sub['wav'], sub['mask'], batch['mean'], batch['std'], 0.0, cfg,
                                     video_features=video, input='av')
        m.assign_vars(init_canonical(m.engine.layout, seed=5, bias_scale=0.05))
        pred = m.prediction.cpu().numpy()
        sums = m._loss_pass(False)['sums'].cpu().numpy().copy()
        return pred, sums
    big, sums_big = run(np.arange(B))
    pick = np.array([0, 100, 127, 128, 255])
    small, _ = run(pick)
    assert rel_l2(big[pick], small) < TOL
    lo, s_lo = run(np.arange(0, 130))
    hi, s_hi = run(np.arange(130, B))
    assert rel_l2(np.concatenate([lo, hi]), big) < TOL
    assert np.allclose(s_lo[:6] + s_hi[:6], sums_big[:6], rtol=2e-3)


def test_si_full_length_utterance():
    """GRID shape: 3 s, T = 250 -- error growth over the 250-step chain stays inside the budget."""
    from oracle import blstm as oblstm
    model, batch, canon, inp = _build('av-blstm', 2, 48000, seed=9)
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len']),
                                         canon, 3)
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    _check_grads(model.canonical_gradients(), ograds, 'av-blstm T=250')


@pytest.mark.parametrize('model_name', ['av-blstm-ssnn-ctc', 'a-blstm-ctc'])
def test_mtl_forward_loss_gradients(model_name):
    from oracle import blstm as oblstm
    B, audio_len = 4, 19200          # T = 100 >= 2 * 24 + 1 label states
    model, batch, canon, inp = _build(model_name, B, audio_len, seed=21, ctc_loss=0.05)
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads(
        'mtl', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len'], labels=batch['labels'],
                    lab_len=batch['lab_len']), canon, 3, ctc_weight=0.05)
    ipt, asr = model.inference
    assert rel_l2(ipt.cpu().numpy(), outs['inference']) < TOL
    assert rel_l2(asr.cpu().numpy(), outs['logits_asr']) < TOL
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss_hole) - float(outs['loss_hole'])) < TOL * float(outs['loss_hole'])
    assert abs(float(model.ctc_loss) - float(outs['ctc_loss'])) < TOL * float(outs['ctc_loss'])
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * float(outs['loss'])
    _check_grads(model.canonical_gradients(), ograds, model_name)
    # `decoding` = beam search (width 20, merge_repeated) on the asr logits, `per` = edit distance / label length:
    # the host routine on the CUDA logits against the restatement on the SAME logits (near-ties between beams could
    # otherwise flip with the 2e-3 logit tolerance), plus greedy decoding against the oracle's best path
    from oracle import ctc as octc
    dec = model.decoding
    got_asr = asr.cpu().numpy()
    assert dec.shape[0] == B and dec.max() < 33
    for b in range(B):
        ref, _ = octc.ctc_beam_search(got_asr[b, :int(batch['seq_len'][b])], beam_width=20, merge_repeated=True)
        assert [int(x) for x in dec[b] if x >= 0] == ref, b
    per = model.per
    for b in range(B):
        hyp = [int(x) for x in dec[b] if x >= 0]
        lab = [int(x) for x in batch['labels'][b, :int(batch['lab_len'][b])]]
        d = np.zeros((len(hyp) + 1, len(lab) + 1), np.int64)
        d[:, 0], d[0, :] = np.arange(len(hyp) + 1), np.arange(len(lab) + 1)
        for i in range(1, len(hyp) + 1):
            for j in range(1, len(lab) + 1):
                d[i, j] = min(d[i - 1, j] + 1, d[i, j - 1] + 1, d[i - 1, j - 1] + (hyp[i - 1] != lab[j - 1]))
        assert abs(per[b] - d[-1, -1] / max(1, len(lab))) < 1e-6
    model.decoder = 'greedy'
    greedy = octc.greedy_decode(np.transpose(got_asr, (1, 0, 2)), batch['seq_len'])
    gd = model.decoding
    for b in range(B):
        assert [int(x) for x in gd[b] if x >= 0] == [int(x) for x in greedy[b]]


def test_train_step_matches_oracle_adam_and_learns():
    from oracle import adam as oadam
    from oracle import blstm as oblstm
    model, batch, canon, inp = _build('av-blstm', 4, 4800, seed=5)
    tsn, net_in = _oracle_inputs(batch, inp)
    inputs = dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len'])
    outs, ograds = oblstm.loss_and_grads('si', inputs, canon, 3)
    l0 = float(model.loss)
    model.train_op()
    new = model.engine.export_canonical()
    # TF Adam, first step: every touched weight moves by ~lr * sign(g)
    upd_ref, upd_got = [], []
    for k in sorted(canon):
        th, _, _ = oadam.adam_tf_step(canon[k], ograds[k], np.zeros_like(canon[k]), np.zeros_like(canon[k]), 1)
        big = np.abs(ograds[k]) > 1e-6 * np.abs(ograds[k]).max()         # sign of tiny gradients is noise
        upd_ref.append((th - canon[k])[big])
        upd_got.append((new[k] - canon[k])[big])
    ur, ug = np.concatenate(upd_ref), np.concatenate(upd_got)
    assert rel_l2(ug, ur) < 0.02
    assert model.global_step == 1
    losses = [l0]
    for _ in range(15):
        model.feed()                 # same batch: invalidate caches
        model.train_op()
    model.feed()
    losses.append(float(model.loss))
    assert np.isfinite(losses[-1]) and losses[-1] < 0.9 * losses[0], losses
    # padded parameters never move
    pad = torch.from_numpy(1.0 - model.engine.layout.pad_mask()).to(model.device)
    assert float((model.engine.theta.abs() * pad).max()) == 0.0


@pytest.mark.parametrize('model_name,B', [('av-blstm', 8), ('av-blstm-ssnn-ctc', 4), ('av-blstm', 256)])
def test_captured_train_step_equals_eager_steps(model_name, B):
    """capture_train_step(): the CUDA-graph replay of feed -> train_op follows the eager steps -- same losses, same
    weights after four updates on four different batches (to a tolerance: the replayed optimiser takes its bias correction
    from the device-resident count), Adam's
    bias correction advancing from the device-resident count, shape changes and unsupported settings refused."""
    from avsi_b200 import _lib, av_sync, synth
    eager, batch, canon, inp = _build(model_name, B, 4800, seed=3)
    graphed, _, _, _ = _build(model_name, B, 4800, seed=3)
    batches = [batch] + [synth.make_batch(B, audio_len=4800, seed=30 + i) for i in range(3)]

    def feeds(b):
        video = av_sync.video_pipeline(b['landmarks'], b['T'], b['vmean'], b['vstd'])
        f = dict(target_sources=b['wav'], masks=b['mask'], sequence_lengths=b['seq_len'], video_features=video)
        if eager.MTL:
            f.update(labels=b['labels'], labels_lengths=b['lab_len'])
        return f
    with pytest.raises(_lib.AvsiError):
        graphed.capture_train_step()                     # no eager step yet
    losses_e, losses_g = [], []
    for m in (eager, graphed):
        m.feed(**feeds(batches[0]))
        m.train_op()
    step = graphed.capture_train_step()
    assert graphed.global_step == 1 and graphed.engine.step_count == 1
    for b in batches[1:]:
        eager.feed(**feeds(b))
        eager.train_op()
        losses_e.append(float(eager.loss))
        step(**feeds(b))
        losses_g.append(float(graphed.loss))
    assert graphed.global_step == eager.global_step == 4 and graphed.engine.step_count == 4
    assert np.allclose(losses_g, losses_e, rtol=2e-4), (losses_g, losses_e)
    we, wg = eager.engine.export_canonical(), graphed.engine.export_canonical()
    moved = np.concatenate([(we[k] - canon[k]).ravel() for k in sorted(canon)])
    diff = np.concatenate([(we[k] - wg[k]).ravel() for k in sorted(canon)])
    assert np.linalg.norm(moved) > 0 and np.linalg.norm(diff) < 2e-2 * np.linalg.norm(moved)
    # eager steps keep working on the same model afterwards, from the same count
    graphed.feed(**feeds(batches[0]))
    graphed.train_op()
    assert graphed.global_step == 5 and np.isfinite(float(graphed.loss))
    with pytest.raises(ValueError):
        step(target_sources=batches[0]['wav'][:, :100])
    graphed.optimizer_choice = 'sgd'
    with pytest.raises(_lib.AvsiError):
        graphed.capture_train_step()


@pytest.mark.parametrize('model_name,B', [('av-blstm', 5), ('av-blstm-ssnn-ctc', 3), ('a-blstm', 130)])
def test_fused_head_l1_equals_two_launches(model_name, B, monkeypatch):
    """avsi_head_l1 (AVSI_FUSE_HEAD_L1=1: head GEMM with the masked-L1 loss in its epilogue) against the default head GEMM +
    avsi_masked_l1: the six loss sums, the CTC loss of the MTL model (it reads the phone columns the fused kernel still
    writes) and the complete parameter gradient."""
    res = {}
    for fuse in ('0', '1'):
        monkeypatch.setenv('AVSI_FUSE_HEAD_L1', fuse)
        model, batch, canon, inp = _build(model_name, B, 4800, seed=11)
        model.compute_gradients()
        sums = model._loss_pass(False, want_pred=False)['sums'].cpu().numpy().copy()
        grads = model.canonical_gradients()
        res[fuse] = (sums, grads, float(model.loss))
    assert np.allclose(res['0'][0][:6], res['1'][0][:6], rtol=1e-6)
    assert abs(res['0'][2] - res['1'][2]) <= 1e-6 * abs(res['0'][2])
    ga = np.concatenate([res['0'][1][k].ravel() for k in sorted(res['0'][1])])
    gb = np.concatenate([res['1'][1][k].ravel() for k in sorted(res['0'][1])])
    assert rel_l2(gb, ga) < 1e-5


def test_variables_roundtrip_and_inference_mode():
    from avsi_b200 import av_sync, models, synth
    model, batch, canon, inp = _build('av-blstm', 2, 4800, seed=7)
    model.build_graph('av-blstm')
    tv = model.train_vars
    assert set(tv) == set('av-blstm/' + k for k in canon)
    assert tv['av-blstm/logits/weights'].shape == (500, 257)
    for k in canon:
        assert np.allclose(tv['av-blstm/' + k], canon[k].astype(np.float32), atol=0)
    cfg = synth.default_config('av-blstm', batch_size=2, audio_len=4800)
    video = av_sync.video_pipeline(batch['landmarks'], batch['T'], batch['vmean'], batch['vstd'])
    m2 = models.StackedBLSTMModel(batch['seq_len'], batch['wav'], batch['mask'], batch['mean'], batch['std'], 0.0, cfg,
                                  video_features=video, input='av', is_training=False)
    m2.build_graph('av-blstm')
    m2.assign_vars(model.all_vars)
    assert torch.equal(m2.prediction, model.prediction)
    enh = m2.enhanced_sources
    assert enh.shape == (2, 4800) and torch.isfinite(enh).all()
    with pytest.raises(Exception):
        m2.train_op()


@pytest.mark.parametrize('opt', ['sgd', 'momentum'])
def test_sgd_and_momentum_updates(opt):
    """models.py:165-173: staircase exponential decay of the rate, GradientDescent / Momentum(0.9) updates."""
    model, batch, canon, inp = _build('a-blstm', 3, 2400, seed=2)
    model.optimizer_choice = opt
    model.starter_learning_rate, model.learning_decay, model.updating_step = 0.5, 0.5, 2
    theta0 = model.engine.theta.clone()
    acc = torch.zeros_like(theta0)
    expect = theta0.clone()
    for step in range(3):
        lr = 0.5 * 0.5 ** (step // 2)
        assert abs(model.learning_rate - lr) < 1e-12
        model.compute_gradients()
        n = model.engine.layout.n_params_padded
        g = model.engine.grad[:n].clone() / (3 * batch['T'] * 257)
        model._cache = {}
        model.train_op()
        if opt == 'momentum':
            acc = 0.9 * acc + g
            expect = expect - lr * acc
        else:
            expect = expect - lr * g
        assert rel_l2(model.engine.theta.cpu().numpy(), expect.cpu().numpy()) < 1e-5
        model.feed(target_sources=batch['wav'])


def _write_dataset(root, n_train, n_val, audio_len, seed):
    """Synthetic GRID-like TFRecord folders + normalisation files + config file for train()."""
    import os
    from avsi_b200 import synth, tfrecord_io as tio
    from oracle import video as ovideo
    T = -(-audio_len // 192)
    for split, n, sd in (('training-set', n_train, seed), ('validation-set', n_val, seed + 1), ('test-set', n_val, seed + 2)):
        d = os.path.join(root, split)
        os.makedirs(d)
        b = synth.make_batch(n, audio_len=audio_len, seed=sd)
        for i in range(n):
            vid = ovideo.video_features(b['landmarks'][i].astype(np.float64), T, b['vmean'][i], b['vstd'][i])
            rec = tio.serialize_sample_fixed(T, int(b['lab_len'][i]), b['wav'][i], vid, b['mask'][i], b['labels'][i],
                                             '%s_%03d' % (split[:2], i))
            tio.write_records(os.path.join(d, 'data_%05d.tfrecord' % (i + 1)), [rec])
        np.save(os.path.join(d, 'seq_lengths.npy'), np.full(n, T))
    np.save(os.path.join(root, 'mean.npy'), np.full(257, 6.0))
    np.save(os.path.join(root, 'std.npy'), np.full(257, 2.0))
    return T


@pytest.mark.parametrize('model_name', ['av-blstm', 'av-blstm-ssnn-ctc'])
def test_train_infer_mask_app_jobs(tmp_path, model_name, capsys):
    """The drop-in job entry points: train(config_file) on TFRecords, checkpoint under TF names,
    infer(...) writing enhanced wavs, mask_app(...) writing masked wavs."""
    import os
    from scipy.io import wavfile
    from avsi_b200 import checkpoint, inference, masking, training
    root = str(tmp_path / 'data')
    os.makedirs(root)
    audio_len = 9600
    T = _write_dataset(root, 12, 4, audio_len, seed=40)
    exp = str(tmp_path / 'exp' / 'run1')
    cfg = str(tmp_path / 'blstm.config')
    with open(cfg, 'w') as f:
        f.write('# test config\nroot_folder = %s\nexp_folder = %s\nmodel = %s\naudio_feat_dim = 257\nvideo_feat_dim = 136\n'
                'audio_len = %d\nbatch_size = 4\nnet_dim = [250,250,250]\nstarter_learning_rate = 0.001\nmax_n_epochs = 3\n'
                'n_earlystop_epochs = 5\nlr_decay = 1.0\noptimizer_type = adam\nl2 = 0.0\ndropout_rate = 0.0\nctc_loss = 0.001\n'
                'audio_feat_mean = %s\naudio_feat_std = %s\n' % (root, exp, model_name, audio_len,
                                                                os.path.join(root, 'mean.npy'), os.path.join(root, 'std.npy')))
    model = training.train(cfg)
    out = capsys.readouterr().out
    assert '+---- Done training: epoch limit reached ----+' in out and 'Model saved in file' in out
    log = open(os.path.join(exp, 'training_log.txt')).read().splitlines()
    rows = [l for l in log if l[:1].isdigit()]
    assert len(rows) == 3 and rows[0].split('\t')[0] == '1'
    first, last = float(rows[0].split('\t')[2].split('|')[0]), float(rows[-1].split('\t')[2].split('|')[0])
    assert np.isfinite(last) and last < first                                   # the training loss went down
    for fn in ('config.txt', 'audio_features_mean.npy', 'audio_features_std.npy', 'sinet.npz'):
        assert os.path.exists(os.path.join(exp, 'netmodel', fn)), fn
    ck = checkpoint.load(os.path.join(exp, 'netmodel', 'sinet'))
    pre = model_name + '/cudnn_lstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/'
    assert ck[pre + 'kernel'].shape == (393 + 250, 1000)
    assert ck[model_name + '/' + pre + 'kernel/Adam_1'].shape == (643, 1000)       # slot scope nests in the model scope
    assert int(ck[model_name + '/Variable']) == 9 and (model_name + '/optimizer/beta1_power') in ck
    # inference job: restores sinet, writes one wav per test utterance
    audio_out = str(tmp_path / 'audio')
    hole = inference.infer(os.path.join(exp, 'netmodel'), os.path.join(root, 'test-set'), audio_out, 'enh', batch_size=2)
    assert np.isfinite(hole)
    rate, w = wavfile.read(os.path.join(audio_out, 'te_000', 'enhanced', 'enh.wav'))
    assert rate == 16000 and w.dtype == np.int16 and len(w) == T * 192
    assert masking.mask_app(os.path.join(root, 'test-set'), audio_out, num_audio_samples=audio_len, batch_size=4) == 4
    rate, mw = wavfile.read(os.path.join(audio_out, 'te_003', 'masked.wav'))
    assert len(mw) == T * 192 and np.abs(mw.astype(np.int32)).max() > 0


def test_feed_accepts_storage_dtypes():
    """int16 / int32 samples and uint8 / bool masks are widened on the device: identical results to the fp32 feed."""
    model, batch, canon, inp = _build('av-blstm', 3, 4800, seed=4)
    ref = model.prediction.cpu().numpy().copy()
    for wav_dt, mask_dt in ((np.int16, np.uint8), (np.int32, np.bool_)):
        model.feed(target_sources=batch['wav'].astype(wav_dt), masks=batch['mask'].astype(mask_dt))
        assert model._fed['target_sources'].dtype == torch.float32 and model._fed['masks'].dtype == torch.float32
        assert np.array_equal(model.prediction.cpu().numpy(), ref)


@pytest.mark.parametrize('inp,apply_mask,B', [('a', False, 4), ('av', True, 3), ('a', True, 230)])
def test_asr_model_forward_loss_gradients(inp, apply_mask, B):
    """models_asr.StackedBLSTMModel: power-2 spectrogram (x mask) -> log-mel-80 -> norm -> 2-layer BLSTM
    -> CTC, against the float64 oracle (same tolerances as the inpainting models)."""
    from avsi_b200 import av_sync, models_asr, synth
    from avsi_b200.layout import init_canonical
    from oracle import blstm as oblstm
    from oracle import stft as ostft
    from oracle import video as ovideo
    audio_len = 11520                                # T = 60 >= 2 * 24 + 1 label states
    batch = synth.make_batch(B, audio_len=audio_len, seed=50 + B)
    T = batch['T']
    cfg = synth.default_config('asr-blstm', batch_size=B, audio_len=audio_len, net_dim=(250, 250))
    rng = np.random.default_rng(1)
    fmean = rng.normal(8.0, 0.5, 80).astype(np.float32)
    fstd = rng.uniform(2.0, 3.0, 80).astype(np.float32)
    video = av_sync.video_pipeline(batch['landmarks'], T, batch['vmean'], batch['vstd']) if inp == 'av' else None
    model = models_asr.StackedBLSTMModel(batch['seq_len'], batch['lab_len'], batch['wav'], batch['mask'], batch['labels'],
                                         fmean, fstd, 0.0, cfg, video_features=video, input=inp, apply_mask=apply_mask)
    canon = init_canonical(model.engine.layout, seed=7, bias_scale=0.05)
    model.assign_vars(canon)
    # oracle features (models_asr.py:30-37)
    st = ostft.get_stft(batch['wav'].astype(np.float64), window_size=24, step_size=12, out_shape=(B, T, 257))
    spec = np.abs(st) ** 2
    if apply_mask:
        spec = spec * batch['mask'].astype(np.float64)
    fb = ostft.get_log_mel_spectrogram(spec)
    net_in = (fb - fmean.astype(np.float64)) / fstd.astype(np.float64)
    if inp == 'av':
        vid = np.stack([ovideo.video_features(batch['landmarks'][b].astype(np.float64), T, batch['vmean'][b].astype(np.float64),
                                              batch['vstd'][b].astype(np.float64)) for b in range(B)])
        net_in = np.concatenate([net_in, vid], 2)
    outs, ograds = oblstm.loss_and_grads('asr', dict(net_in=net_in, seq_len=batch['seq_len'], labels=batch['labels'],
                                                     lab_len=batch['lab_len']), canon, 2)
    assert rel_l2(model.target_fbanks_norm.cpu().numpy(), net_in[:, :, :80]) < 1e-4
    assert rel_l2(model.inference.cpu().numpy(), outs['inference']) < TOL
    assert abs(float(model.ctc_loss) - float(outs['ctc_loss'])) < TOL * float(outs['ctc_loss'])
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * float(outs['loss'])
    grads = model.canonical_gradients()
    if not apply_mask:
        _check_grads(grads, ograds, 'asr ' + inp)
    else:
        # Masked frames enter the network at (log 1e-6 - mean) / std, about -9 here: the cells saturate and the layer-0
        # gradient becomes ill-conditioned in its INPUTS (in the float64 oracle, rounding net_in through fp16 alone moves
        # the cell_0 bw kernel gradient by 7e-3, rounding the weights by 4e-3).  The kernels' own arithmetic is held to
        # TOL against the oracle evaluated at the operands the tensor cores see (inputs and matrices rounded through
        # fp16); against the full-precision oracle the bound is 10 x TOL.
        _check_grads(grads, ograds, 'asr ' + inp + ' (full-precision oracle)', tol=10 * TOL)
        canon16 = {k: (_through_f16(v) if k.endswith(('kernel', 'weights')) else v) for k, v in canon.items()}
        _, ograds16 = oblstm.loss_and_grads('asr', dict(net_in=_through_f16(net_in), seq_len=batch['seq_len'],
                                                        labels=batch['labels'], lab_len=batch['lab_len']), canon16, 2)
        _check_grads(grads, ograds16, 'asr ' + inp + ' (oracle at fp16 operands)')
    th0 = model.engine.theta.clone()
    model.train_op()
    assert not torch.equal(th0, model.engine.theta) and model.per.shape == (B,)


@pytest.mark.parametrize('model_name,B', [('av-blstm', 4), ('av-blstm-ssnn-ctc', 230)])
def test_dropout_forward_and_gradients(model_name, B):
    """dropout_rate > 0 (tf.nn.dropout on the BLSTM outputs ahead of the heads, models.py:117 / :1901): the oracle is
    given the keep mask the kernel drew; outputs, loss and gradients then agree to the usual tolerances.  Every feed
    (= sess.run) draws a new mask."""
    from oracle import blstm as oblstm
    audio_len = 11520 if 'ctc' in model_name else 4800
    kw = dict(ctc_loss=0.05) if 'ctc' in model_name else {}
    model, batch, canon, inp = _build(model_name, B, audio_len, seed=60 + B, **kw)
    rate = 0.3
    model.feed(dropout_rate=rate)
    keep = model.dropout_keep_mask().cpu().numpy()
    assert keep.shape == (B, batch['T'], 500) and abs(keep.mean() - (1 - rate)) < 0.01
    tsn, net_in = _oracle_inputs(batch, inp)
    if 'ctc' in model_name:
        outs, ograds = oblstm.loss_and_grads(
            'mtl', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len'], labels=batch['labels'],
                        lab_len=batch['lab_len']), canon, 3, ctc_weight=0.05, drop=(keep, rate))
    else:
        outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'],
                                                        seq_len=batch['seq_len']), canon, 3, drop=(keep, rate))
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * abs(float(outs['loss']))
    _check_grads(model.canonical_gradients(), ograds, model_name + ' dropout')
    p1 = model.prediction.cpu().numpy()
    model.feed(dropout_rate=rate)                      # next run: another mask
    assert not np.array_equal(model.dropout_keep_mask().cpu().numpy(), keep)
    assert not np.array_equal(model.prediction.cpu().numpy(), p1)
    model.feed(dropout_rate=0.0)
    assert model.dropout_keep_mask() is None


def test_checkpoint_round_trip_through_tf_bundle(tmp_path):
    """saver.save / saver.restore through TensorFlow's own container (tf_bundle.py): variables under the TF names,
    the global step, the Adam slots and the step count recovered from beta1_power."""
    from avsi_b200 import checkpoint, tf_bundle
    model, batch, canon, inp = _build('av-blstm-ssnn-ctc', 3, 11520, seed=9, ctc_loss=0.05)
    for _ in range(3):
        model.train_op()
    prefix = str(tmp_path / 'netmodel' / 'sinet')
    checkpoint.save(model, prefix, fmt='tf')
    raw = tf_bundle.read_bundle(prefix)
    pre = 'cudnn_lstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/kernel'
    assert any(k.endswith(pre) for k in raw) and any(k.endswith(pre + '/Adam_1') for k in raw)
    b1p = [k for k in raw if k.endswith('beta1_power')]
    assert len(b1p) == 1 and abs(float(raw[b1p[0]]) - 0.9 ** 4) < 1e-6      # TF: beta1^(t+1) after t steps
    other, _, _, _ = _build('av-blstm-ssnn-ctc', 3, 11520, seed=10, ctc_loss=0.05)
    checkpoint.restore(other, prefix)
    assert torch.equal(other.engine.theta, model.engine.theta)
    assert torch.equal(other.engine.adam_m, model.engine.adam_m) and torch.equal(other.engine.adam_v, model.engine.adam_v)
    assert other.engine.step_count == 3 and other.global_step == model.global_step == 3
    # a further step on both: `model` steps again on the same feed (activations recomputed with the moved weights),
    # `other` from the restored state: the same bits (every reduction of the step adds in a fixed order)
    model.train_op()
    other.feed(**{k: v for k, v in model._fed.items()})
    other.train_op()
    assert torch.equal(other.engine.theta, model.engine.theta)


def test_restore_reference_style_checkpoints(tmp_path):
    """What a checkpoint written by the reference's TRAINING graph looks like to us: canonical weights (CudnnLSTMSaveable),
    the global step, Adam slots of the opaque cuDNN blob only, float32 beta powers.  Weights and step restore, the
    moments restart (warning), nothing raises KeyError; a checkpoint lacking a weight raises ValueError (what
    training.py:154-166 catches); a momentum run saves and restores its accumulator."""
    from avsi_b200 import checkpoint
    model, batch, canon, inp = _build('av-blstm', 3, 2400, seed=12)
    model.build_graph('av-blstm')
    ref_like = {('av-blstm/' + k): np.asarray(v, np.float32) for k, v in canon.items()}
    ref_like['av-blstm/Variable'] = np.asarray(1500, np.int32)
    ref_like['av-blstm/av-blstm/cudnn_lstm/opaque_kernel/Adam'] = np.zeros(10, np.float32)
    ref_like['av-blstm/av-blstm/cudnn_lstm/opaque_kernel/Adam_1'] = np.zeros(10, np.float32)
    ref_like['av-blstm/optimizer/beta1_power'] = np.asarray(0.9 ** 1501, np.float32)          # underflows to 0
    ref_like['av-blstm/optimizer/beta2_power'] = np.asarray(0.999 ** 1501, np.float32)
    names = sorted(ref_like)
    import json
    arrays = {'v%05d' % i: ref_like[n] for i, n in enumerate(names)}
    arrays['__names__'] = np.frombuffer(json.dumps(names).encode(), np.uint8)
    np.savez(str(tmp_path / 'ref.npz'), **arrays)
    checkpoint.restore(model, str(tmp_path / 'ref'))
    assert model.global_step == 1500 and model.engine.step_count == 1500              # from `Variable`: beta1_power is 0
    assert float(model.engine.adam_m.abs().max()) == 0.0
    got = model.engine.export_canonical()
    assert all(np.array_equal(got[k], np.asarray(canon[k], np.float32)) for k in canon)
    ref_like['av-blstm/optimizer/beta1_power'] = np.asarray(0.9 ** 8, np.float32)             # a young checkpoint: t = 7
    assert checkpoint._adam_step_from(ref_like, model) == 7
    broken = dict(ref_like)
    del broken['av-blstm/logits/weights']
    names = sorted(broken)
    arrays = {'v%05d' % i: broken[n] for i, n in enumerate(names)}
    arrays['__names__'] = np.frombuffer(json.dumps(names).encode(), np.uint8)
    np.savez(str(tmp_path / 'broken.npz'), **arrays)
    with pytest.raises(ValueError):
        checkpoint.restore(model, str(tmp_path / 'broken'))
    # momentum accumulator round trip
    mom, _, _, _ = _build('av-blstm', 3, 2400, seed=12, optimizer_type='momentum')
    mom.build_graph('av-blstm')
    mom.train_op()
    mom.train_op()
    checkpoint.save(mom, str(tmp_path / 'mom'))
    other, _, _, _ = _build('av-blstm', 3, 2400, seed=13, optimizer_type='momentum')
    other.build_graph('av-blstm')
    checkpoint.restore(other, str(tmp_path / 'mom'))
    assert torch.equal(other.engine.momentum_acc, mom.engine.momentum_acc) and float(mom.engine.momentum_acc.abs().max()) > 0
    assert torch.equal(other.engine.theta, mom.engine.theta)


def test_embedding_model_forward_loss_gradients():
    """StackedBLSTMEmbeddingModel (models.py:1120-1472, integration_layer 0): the utterance's 512-d embedding replicated
    over the frames and concatenated to the network input; layer-0 kernel [(393 + 512) + 250, 1000]."""
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    from oracle import blstm as oblstm
    B, audio_len = 3, 4800
    batch = synth.make_batch(B, audio_len=audio_len, seed=14)
    T = batch['T']
    emb = (np.random.default_rng(3).standard_normal((B, 512)) * 0.5).astype(np.float32)
    cfg = synth.default_config('av-blstm-emb', batch_size=B, audio_len=audio_len)
    video = av_sync.video_pipeline(batch['landmarks'], T, batch['vmean'], batch['vstd'])
    cls, inp = models.MODEL_REGISTRY['av-blstm-emb']
    model = cls(batch['seq_len'], batch['wav'], batch['mask'], batch['mean'], batch['std'], 0.0, cfg, video_features=video,
                embeddings=emb, input=inp)
    assert model.engine.layout.in_dim == 393 + 512
    canon = init_canonical(model.engine.layout, seed=15, bias_scale=0.05)
    model.assign_vars(canon)
    assert canon['cudnn_lstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/kernel'].shape == (905 + 250, 1000)
    tsn, net_in = _oracle_inputs(batch, 'av')
    net_in = np.concatenate([net_in, np.repeat(emb[:, None, :].astype(np.float64), T, axis=1)], 2)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len']),
                                         canon, 3)
    assert rel_l2(model.net_inputs.cpu().numpy(), net_in) < 1e-3
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * abs(float(outs['loss']))
    _check_grads(model.canonical_gradients(), ograds, 'av-blstm-emb')
    with pytest.raises(ValueError):
        model.feed(embeddings=emb[:, :100])
        model.prediction


@pytest.mark.parametrize('model_name,B,audio_len', [('av-blstm-ssnn', 3, 4800), ('a-blstm-ssnn', 230, 1920)])
def test_ssnn_model_forward_loss_gradients(model_name, B, audio_len):
    """StackedBLSTMSSNNModel (models.py:718-1117, integration_layer 0): the speaker embedding computed from the corrupted
    spectrogram by the trainable 3-layer network and appended to every input frame -- forward values, the embedding
    itself and the gradients of ALL variables (BLSTM stack, head, speaker_embedding/*) against the float64 oracle."""
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    from oracle import blstm as oblstm
    from oracle import stft as ostft
    batch = synth.make_batch(B, audio_len=audio_len, seed=24)
    T = batch['T']
    cfg = synth.default_config(model_name, batch_size=B, audio_len=audio_len)
    cls, inp = models.MODEL_REGISTRY[model_name]
    video = av_sync.video_pipeline(batch['landmarks'], T, batch['vmean'], batch['vstd'])
    model = cls(batch['seq_len'], batch['wav'], batch['mask'], batch['mean'], batch['std'], 0.0, cfg,
                video_features=video if inp != 'a' else None, input=inp)
    base = {'a': 257, 'av': 393}[inp]
    assert model.engine.layout.in_dim == base + 200
    canon = init_canonical(model.engine.layout, seed=25, bias_scale=0.05)
    assert canon['speaker_embedding/weights_1'].shape == (514, 200) and canon['speaker_embedding/biases_3'].shape == (200,)
    model.assign_vars(canon)
    tsn, net_in = _oracle_inputs(batch, inp)
    audio_feat = tsn * batch['mask']
    delta_inp = ostft.add_delta_features(audio_feat, n_delta=1, N=2)
    outs, ograds = oblstm.loss_and_grads('ssnn', dict(net_in=net_in, delta_inp=delta_inp, target=tsn, mask=batch['mask'],
                                                      seq_len=batch['seq_len']), canon, 3)
    assert rel_l2(model.speaker_embedding.cpu().numpy(), outs['speaker_embedding']) < TOL
    assert rel_l2(model.speaker_embedding_ext.cpu().numpy(), outs['speaker_embedding_ext']) < TOL
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * abs(float(outs['loss']))
    grads = model.canonical_gradients()
    # The whole gradient against the oracle: TOL.  Per variable the dense layers below a leaky ReLU are delicate: the
    # activation has a kink, so a pre-activation within rounding error of 0 whose sign differs between the two
    # evaluations switches its derivative between 1 and 0.3 -- in ANY arithmetic narrower than the oracle's.  The kernels'
    # own arithmetic is therefore held per variable to 2 x TOL against the oracle evaluated WITH THE ACTIVATION PATTERN OF
    # THE CUDA RUN (identical wherever the signs agree), and to 5 x TOL against the plain oracle.
    ga = np.concatenate([grads[k].ravel() for k in sorted(ograds)])
    gb = np.concatenate([ograds[k].ravel() for k in sorted(ograds)])
    assert rel_l2(ga, gb) < TOL, rel_l2(ga, gb)
    slopes = [np.where(model._ssnn['z%d' % k].cpu().numpy() > 0, 1.0, 0.3) for k in (0, 1)]
    outs_p, ograds_p = oblstm.loss_and_grads('ssnn', dict(net_in=net_in, delta_inp=delta_inp, target=tsn, mask=batch['mask'],
                                                          seq_len=batch['seq_len']), canon, 3, slopes=slopes)
    assert rel_l2(outs_p['speaker_embedding'], outs['speaker_embedding']) < 1e-4         # same function up to the near-zero cases
    for k in ograds:
        if np.linalg.norm(ograds[k]) > 0:
            assert rel_l2(grads[k], ograds_p[k]) < 2 * TOL, (k, rel_l2(grads[k], ograds_p[k]), rel_l2(grads[k], ograds[k]))
            assert rel_l2(grads[k], ograds[k]) < 5 * TOL, (k, rel_l2(grads[k], ograds[k]))
    assert all(np.linalg.norm(ograds[k]) > 0 for k in ('speaker_embedding/weights_1', 'speaker_embedding/weights_3',
                                                       'speaker_embedding/biases_2'))
    th0 = model.engine.theta.clone()
    model.train_op()
    moved = model.engine.export_canonical()
    assert not np.array_equal(moved['speaker_embedding/weights_2'], canon['speaker_embedding/weights_2'].astype(np.float32))
    assert not torch.equal(th0, model.engine.theta)


def test_two_step_model_forward_loss_gradients_and_summaries():
    """StackedBLSTM2StepsModel (models.py:240-317): v-blstm predicts the spectrogram from the motion vectors, its
    prediction is the audio input of av-blstm-twosteps; only the av model's variables are trained."""
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    from oracle import blstm as oblstm
    B, audio_len = 3, 4800
    batch = synth.make_batch(B, audio_len=audio_len, seed=16)
    T = batch['T']
    cfg = synth.default_config('av-blstm-twosteps', batch_size=B, audio_len=audio_len)
    video = av_sync.video_pipeline(batch['landmarks'], T, batch['vmean'], batch['vstd'])
    model = models.StackedBLSTM2StepsModel(batch['seq_len'], batch['wav'], batch['mask'], batch['mean'], batch['std'], 0.0,
                                           cfg, video)
    model.build_graph('av-blstm-twosteps')
    canon_v = init_canonical(model.video_model.engine.layout, seed=17, bias_scale=0.05)
    canon_av = init_canonical(model.av_model.engine.layout, seed=18, bias_scale=0.05)
    variables = {'v-blstm/' + k: v for k, v in canon_v.items()}
    variables.update({'av-blstm-twosteps/' + k: v for k, v in canon_av.items()})
    model.assign_vars(variables)
    tsn, net_in = _oracle_inputs(batch, 'av')
    vid = net_in[:, :, 257:]
    common = dict(target=tsn, mask=batch['mask'], seq_len=batch['seq_len'])
    outs_v, _ = oblstm.loss_and_grads('si', dict(net_in=vid, **common), canon_v, 3)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=np.concatenate([outs_v['prediction'], vid], 2), **common), canon_av, 3)
    assert rel_l2(model.video_prediction.cpu().numpy(), outs_v['prediction']) < TOL
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * abs(float(outs['loss']))
    _check_grads(model.canonical_gradients(), ograds, 'av-blstm-twosteps')
    assert set(model.train_vars) == {'av-blstm-twosteps/' + k for k in canon_av}            # models.py:295
    assert any(k.startswith('v-blstm/') for k in model.all_vars)
    v0 = model.video_model.engine.theta.clone()
    a0 = model.av_model.engine.theta.clone()
    model.train_op()
    assert torch.equal(model.video_model.engine.theta, v0) and not torch.equal(model.av_model.engine.theta, a0)
    # summaries (models.py:199-219): flipped [n,F,T,1] images and peak-normalised audio
    model.feed(dropout_rate=0.0)
    sm = model.summaries
    assert set(sm) == {'summary/Target_spectrogram', 'summary/Enhanced_spectrogram', 'summary/Mask',
                       'summary/Target_audio', 'summary/Enhanced_audio'}
    kind, img = sm['summary/Target_spectrogram']
    tgt = model.target_spec_norm
    assert kind == 'image' and tuple(img.shape) == (B, 257, T, 1)
    assert torch.equal(img[1, :, :, 0], tgt[1].t().flip(0))
    kind, aud = sm['summary/Enhanced_audio']
    assert kind == 'audio' and tuple(aud.shape) == (B, audio_len) and abs(float(aud.abs().max()) - 1.0) < 1e-6


def test_train_asr_job(tmp_path, capsys):
    """training_asr.train(config_file) (training_asr.py:23): the shared job loop with the phone-recognition model --
    log-mel statistics as normalisation files, CTC loss / PER as monitored figures, checkpoint under the `asr/<model>` scope."""
    import os
    from avsi_b200 import checkpoint, training_asr
    root = str(tmp_path / 'data')
    os.makedirs(root)
    audio_len = 11520                                  # T = 60 >= 2 * 24 + 1 label states
    _write_dataset(root, 6, 3, audio_len, seed=70)
    np.save(os.path.join(root, 'fb_mean.npy'), np.full(80, 8.0))
    np.save(os.path.join(root, 'fb_std.npy'), np.full(80, 2.5))
    exp = str(tmp_path / 'exp' / 'asr1')
    cfg = str(tmp_path / 'blstm_asr.config')
    with open(cfg, 'w') as f:
        f.write('root_folder = %s\nexp_folder = %s\nmodel = a-blstm\naudio_feat_dim = 257\nvideo_feat_dim = 136\n'
                'audio_len = %d\nbatch_size = 3\nnet_dim = [250,250]\nstarter_learning_rate = 0.001\nmax_n_epochs = 2\n'
                'n_earlystop_epochs = 5\nlr_decay = 1.0\noptimizer_type = adam\nl2 = 0.0\ndropout_rate = 0.0\n'
                'audio_feat_mean = %s\naudio_feat_std = %s\n' % (root, exp, audio_len, os.path.join(root, 'fb_mean.npy'),
                                                                os.path.join(root, 'fb_std.npy')))
    model = training_asr.train(cfg)
    out = capsys.readouterr().out
    assert '+---- Done training: epoch limit reached ----+' in out and 'Model saved in file' in out
    rows = [l for l in open(os.path.join(exp, 'training_log.txt')).read().splitlines() if l[:1].isdigit()]
    assert len(rows) == 2
    ctc = [float(r.split('\t')[2].split('|')[1]) for r in rows]
    assert np.isfinite(ctc).all() and ctc[1] < ctc[0]                        # the CTC loss went down
    ck = checkpoint.load(os.path.join(exp, 'netmodel', 'asrnet'))                    # training_asr.py:308
    pre = 'asr/a-blstm/cudnn_lstm/stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/'
    assert ck[pre + 'kernel'].shape == (80 + 250, 1000)
    # `sinet` is the best-validation checkpoint: written after epoch 1 (step 2) and again after epoch 2 only if it improved
    assert ck['asr/a-blstm/logits/weights'].shape == (500, 34) and int(ck['asr/a-blstm/Variable']) in (2, 4)
    assert model.per.shape == (3,)


def test_inference_siasr_job(tmp_path, capsys):
    """inference_siasr_ctc.infer(...) (inference_siasr_ctc.py:22, the CLI's `inference_siasr`): inpainting model -> enhanced
    waveform -> phone recogniser ON the enhanced waveform -> `enhanced/<prefix>.wav` + `transcriptions/<prefix>.lbl`."""
    import os
    from scipy.io import wavfile
    from avsi_b200 import inference_siasr_ctc, training, training_asr
    root = str(tmp_path / 'data')
    os.makedirs(root)
    audio_len = 11520
    T = _write_dataset(root, 4, 2, audio_len, seed=90)
    np.save(os.path.join(root, 'fb_mean.npy'), np.full(80, 8.0))
    np.save(os.path.join(root, 'fb_std.npy'), np.full(80, 2.5))
    common = ('audio_feat_dim = 257\nvideo_feat_dim = 136\naudio_len = %d\nbatch_size = 2\nstarter_learning_rate = 0.001\n'
              'max_n_epochs = 1\nn_earlystop_epochs = 5\nlr_decay = 1.0\noptimizer_type = adam\nl2 = 0.0\ndropout_rate = 0.0\n' % audio_len)
    exp_si, exp_asr = str(tmp_path / 'exp' / 'si'), str(tmp_path / 'exp' / 'asr')
    cfg_si, cfg_asr = str(tmp_path / 'si.config'), str(tmp_path / 'asr.config')
    with open(cfg_si, 'w') as f:
        f.write('root_folder = %s\nexp_folder = %s\nmodel = av-blstm\nnet_dim = [250,250,250]\n%saudio_feat_mean = %s\naudio_feat_std = %s\n'
                % (root, exp_si, common, os.path.join(root, 'mean.npy'), os.path.join(root, 'std.npy')))
    with open(cfg_asr, 'w') as f:
        f.write('root_folder = %s\nexp_folder = %s\nmodel = a-blstm\nnet_dim = [250,250]\n%saudio_feat_mean = %s\naudio_feat_std = %s\n'
                % (root, exp_asr, common, os.path.join(root, 'fb_mean.npy'), os.path.join(root, 'fb_std.npy')))
    training.train(cfg_si)
    training_asr.train(cfg_asr)
    assert os.path.exists(os.path.join(exp_asr, 'netmodel', 'asrnet.npz'))
    dict_file = str(tmp_path / 'dictionary.txt')
    with open(dict_file, 'w') as f:
        f.write('\n'.join('w%d P%02d P%02d' % (i, i, (i * 7) % 33) for i in range(33)))   # 33 phoneme symbols + 33 words
    # the reference's load_dictionary sorts the distinct symbols: ids index that list
    from avsi_b200.transcription2phonemes import get_labels, get_phonemes_from_labels, load_dictionary
    d = load_dictionary(dict_file)
    assert d[:3] == ['P00', 'P01', 'P02'] and len(d) == 66
    assert get_phonemes_from_labels(get_labels('P03,SP,P01', d), d) == ['P03', 'P01']
    capsys.readouterr()
    audio_out = str(tmp_path / 'audio')
    hole, asr, per = inference_siasr_ctc.infer(os.path.join(exp_si, 'netmodel'), os.path.join(exp_asr, 'netmodel'),
                                               os.path.join(root, 'test-set'), audio_out, 'exp0', dict_file, batch_size=2)
    out = capsys.readouterr().out
    assert 'Loss hole:' in out and 'PER:' in out and np.isfinite([hole, asr, per]).all() and per >= 0.0
    for name in ('te_000', 'te_001'):
        rate, w = wavfile.read(os.path.join(audio_out, name, 'enhanced', 'exp0.wav'))
        assert rate == 16000 and w.dtype == np.int16 and len(w) == T * 192
        lbl = open(os.path.join(audio_out, name, 'transcriptions', 'exp0.lbl')).read()
        assert lbl == '' or all(tok in d for tok in lbl.split(','))
