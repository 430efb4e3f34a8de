"""Parity WHERE THE BENCHMARK RUNS.

bench.py times B = 2048 utterances of T = 250 frames (BASELINE configs[1]) and, for configs[4], T = 1667: at those
sizes the step runs the tcgen05 4-CTA-cluster recurrence / BPTT kernels (lstm4.cu, lstm4_bwd.cu) and the CTA-pair
GEMMs.  The float64 oracle cannot run 2048 x 250 frames in seconds, so:

  * the tcgen05 kernels are forced at a small batch (AVSI_LSTM_FWD / AVSI_LSTM_BWD = l4, re-read through
    avsi_reload_env) and compared with the oracle -- outputs AND gradients, <= 2e-3 relative L2 -- over the full
    250- and 1667-step chains, SI and MTL (the BPTT kernel ships the peers' partial dh as fp16 every step: this is
    where its error growth over the chain is measured);
  * the exact bench launch (B = 2048, T = 250: 16 row tiles, 32 clusters) runs on a batch that TILES 8 distinct
    utterances 256 times: every one of the 2048 predictions must match the oracle's for its utterance, and the
    gradient of the mean loss equals the oracle's gradient on the 8 distinct utterances (a mean over identical
    copies), so the whole parameter gradient of the bench-shaped step is checked against float64;
  * large-magnitude weights: the static loss scale does not overflow fp16 for recurrent weights 3x Glorot, and when
    the activations' gradients DO overflow (head weights x 4e4) the overflow guard skips the step, halves the scale
    and training continues with gradients that match the oracle again.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from test_gpu_model import TOL, _build, _check_grads, _oracle_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture
def l4_forced():
    from avsi_b200 import _lib
    _lib.set_env(AVSI_LSTM_FWD='l4', AVSI_LSTM_BWD='l4')
    yield
    _lib.set_env(AVSI_LSTM_FWD=None, AVSI_LSTM_BWD=None)


def chain_parity(model_name, B, audio_len, seed, **build_kw):
    """(rel-L2 of prediction, of the full gradient, worst per-variable gradient) of the CUDA model against the oracle."""
    from oracle import blstm as oblstm
    mtl = model_name.endswith('ctc')
    model, batch, canon, inp = _build(model_name, B, audio_len, seed=seed, **(dict(ctc_loss=0.05) if mtl else {}),
                                      **build_kw)
    tsn, net_in = _oracle_inputs(batch, inp)
    oin = dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len'])
    if mtl:
        oin.update(labels=batch['labels'], lab_len=batch['lab_len'])
        outs, ograds = oblstm.loss_and_grads('mtl', oin, canon, 3, ctc_weight=0.05)
    else:
        outs, ograds = oblstm.loss_and_grads('si', oin, canon, 3)
    e_pred = rel_l2(model.prediction.cpu().numpy(), outs['prediction'])
    e_loss = abs(float(model.loss) - float(outs['loss'])) / abs(float(outs['loss']))
    grads = model.canonical_gradients()
    ga = np.concatenate([grads[k].ravel() for k in sorted(ograds)])
    gb = np.concatenate([ograds[k].ravel() for k in sorted(ograds)])
    worst = max((rel_l2(grads[k], ograds[k]), k) for k in ograds if np.linalg.norm(ograds[k]) > 0)
    return dict(pred=e_pred, loss=e_loss, grad=rel_l2(ga, gb), worst=worst[0], worst_name=worst[1],
                model=model, grads=grads, ograds=ograds, batch=batch, canon=canon)


@pytest.mark.parametrize('model_name,audio_len', [('av-blstm', 48000), ('av-blstm-ssnn-ctc', 48000), ('av-blstm', 320000)])
def test_tcgen05_recurrence_full_chains_vs_oracle(l4_forced, model_name, audio_len):
    """T = 250 (BASELINE configs[1]) and T = 1667 (configs[4], 20 s) through lstm4_fwd / lstm4_bwd at B = 3."""
    r = chain_parity(model_name, 3, audio_len, seed=41)
    assert r['pred'] < TOL, r['pred']
    assert r['loss'] < TOL, r['loss']
    _check_grads(r['grads'], r['ograds'], '%s T=%d (tcgen05 path)' % (model_name, -(-audio_len // 192)))


def test_tcgen05_recurrence_mtl_20s_chain_vs_oracle(l4_forced):
    """AV-MTL-SI at T = 1667 (320064 samples = 1667 whole hops).  With ctc_loss = 0.05 the gradient is the CTC term's (the
    hole-L1 term is normalised by a count that grows with T), so this is where the CTC kernel's precision over a 1667-frame
    chain shows: 5.9e-3 with un-shifted fp32 log-space alpha / beta, 2.1e-3 with the re-centring but the fast exponential,
    4.2e-4 now (profiles/r02_parity_vs_T.json) -- the same as AV-SI."""
    r = chain_parity('av-blstm-ssnn-ctc', 3, 1667 * 192, seed=41)
    assert r['pred'] < TOL and r['loss'] < TOL
    _check_grads(r['grads'], r['ograds'], 'av-blstm-ssnn-ctc T=1667 (tcgen05 path) grad %.2e' % r['grad'])


def test_mtl_20s_ill_conditioned_instance_is_the_problem_not_the_kernels(l4_forced):
    """The same model on another draw (320000 samples, other labels) misses TOL: 2.1e-3.  Decomposed here:
      * the asr logits are as accurate as everywhere else (3.9e-4 relative, 1e-4 absolute);
      * the CTC kernel, against the float64 oracle evaluated on THE SAME logits, is within 5e-4 (measured 1.8e-4);
      * but the float64 oracle's own CTC gradient moves by 2.5e-3 (4.4e-3 for one utterance) when it is evaluated on those
        logits instead of its own: a CTC posterior weighs alignments by the product of 1667 per-frame probabilities, and
        for this draw a 1e-4 perturbation of the logits shifts log p(l|x) by 0.05.
    Any arithmetic with fp16-level (or bf16 / TF32-level) error in the forward pass is subject to that factor; the
    end-to-end bound for the instance is therefore TOL + the measured conditioning."""
    from avsi_b200 import _lib
    from oracle import ctc as octc
    r = chain_parity('av-blstm-ssnn-ctc', 3, 320000, seed=41)
    model, batch = r['model'], r['batch']
    assert r['pred'] < TOL and r['loss'] < TOL
    _, asr = model.inference
    lg_cuda = asr.cpu().numpy().astype(np.float64)                                    # [B,T,C]
    tsn, net_in = _oracle_inputs(batch, 'av')
    from oracle import blstm as oblstm
    outs, _ = oblstm.loss_and_grads('mtl', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len'],
                                                labels=batch['labels'], lab_len=batch['lab_len']), r['canon'], 3, ctc_weight=0.05)
    lg_or = outs['logits_asr']
    assert rel_l2(lg_cuda, lg_or) < 1e-3

    def oracle_ctc_grad(lg):
        x = torch.tensor(np.transpose(lg, (1, 0, 2)), dtype=torch.float64, requires_grad=True)
        octc.ctc_nll_torch(x, torch.from_numpy(batch['labels']).long(), torch.from_numpy(batch['lab_len']),
                           torch.from_numpy(batch['seq_len'])).sum().backward()
        return x.grad.numpy()
    g_or_or, g_or_cu = oracle_ctc_grad(lg_or), oracle_ctc_grad(lg_cuda)
    conditioning = rel_l2(g_or_cu, g_or_or)
    # the kernel on the model's own logits
    lib, d = _lib.load(), model.device
    B, T, C = lg_cuda.shape
    ldl = 64
    lg = torch.zeros(T * B, ldl, device=d)
    lg[:, :C] = asr.permute(1, 0, 2).reshape(T * B, C)
    tl, tn, ts = (torch.from_numpy(np.ascontiguousarray(batch[k]).astype(np.int32)).to(d) for k in ('labels', 'lab_len', 'seq_len'))
    nll = torch.empty(B, device=d)
    dl = torch.zeros(T * B, ldl, dtype=torch.float16, device=d)
    ws = torch.empty(int(lib.avsi_ctc_workspace_bytes(B, T, tl.shape[1])) // 4 + 4, device=d)
    _lib.check(lib.avsi_ctc_loss(_lib.ptr(lg), ldl, 0, C, _lib.ptr(tl), tl.shape[1], _lib.ptr(tn), _lib.ptr(ts), B, T, 8.0, None,
                                 _lib.ptr(nll), _lib.ptr(dl), ldl, 0, _lib.ptr(ws), _lib.stream_ptr()), 'ctc')
    torch.cuda.synchronize()
    g = dl.cpu().numpy().astype(np.float64).reshape(T, B, ldl)[:, :, :C] / 8.0
    assert rel_l2(g, g_or_cu) < 5e-4, rel_l2(g, g_or_cu)
    assert conditioning > 0.5 * TOL, conditioning          # the instance IS ill-conditioned (else this test is moot)
    assert r['grad'] < TOL + conditioning, (r['grad'], conditioning)


def test_tcgen05_and_mma_recurrences_agree_on_long_chain(l4_forced):
    """The same 1667-step problem through both kernel families: the small-batch mma.sync kernels (fp32 partial dh
    exchange) and the tcgen05 ones (fp16 partials) stay within the budget of EACH OTHER as well."""
    from avsi_b200 import _lib
    r4 = chain_parity('av-blstm', 2, 320000, seed=43)
    _lib.set_env(AVSI_LSTM_FWD='mma', AVSI_LSTM_BWD='mma')
    rm = chain_parity('av-blstm', 2, 320000, seed=43)
    assert rm['pred'] < TOL and rm['grad'] < TOL
    ga = np.concatenate([r4['grads'][k].ravel() for k in sorted(r4['grads'])])
    gb = np.concatenate([rm['grads'][k].ravel() for k in sorted(rm['grads'])])
    assert rel_l2(ga, gb) < TOL


def test_bench_launch_b2048_t250_every_row_tile_vs_oracle():
    """The exact launch bench.py times (B = 2048, T = 250, AV-SI): 8 distinct utterances tiled over the 2048 batch
    rows (utterance b = distinct[b % 8], so every 128-row tile of every cluster holds all 8)."""
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    from oracle import blstm as oblstm
    U, B, audio_len = 8, 2048, 48000
    small = synth.make_batch(U, audio_len=audio_len, seed=77)
    T = small['T']
    idx = np.arange(B) % U
    cfg = synth.default_config('av-blstm', batch_size=B, audio_len=audio_len)
    video = av_sync.video_pipeline(small['landmarks'], T, small['vmean'], small['vstd'])
    vid_big = video[torch.as_tensor(idx, device=video.device)] if torch.is_tensor(video) else video[idx]
    model = models.StackedBLSTMModel(small['seq_len'][idx], small['wav'][idx], small['mask'][idx], small['mean'],
                                     small['std'], 0.0, cfg, video_features=vid_big, input='av')
    canon = init_canonical(model.engine.layout, seed=78, bias_scale=0.05)
    model.assign_vars(canon)
    tsn, net_in = _oracle_inputs(small, 'av')
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=small['mask'], seq_len=small['seq_len']),
                                         canon, 3)
    ref = torch.from_numpy(outs['prediction']).to(model.device)                      # [8, T, F] float64
    pred = model.prediction.double().view(B // U, U, T, -1)                          # row b = tile * 8 + utterance
    err = (pred - ref[None]).flatten(2).norm(dim=2) / ref.flatten(1).norm(dim=1)[None]  # per (tile, utterance)
    assert float(err.max()) < TOL, float(err.max())
    assert abs(float(model.loss) - float(outs['loss'])) < TOL * abs(float(outs['loss']))
    _check_grads(model.canonical_gradients(), ograds, 'av-blstm B=2048 T=250')
    # and the optimiser step from that gradient: finite, guard untouched
    model.train_op()
    skipped, scale = model.engine.guard_state()
    assert skipped == 0 and scale == 1.0
    assert bool(torch.isfinite(model.engine.theta).all())


def _scaled(canon, pattern, factor, rows=None):
    out = {k: np.array(v, copy=True) for k, v in canon.items()}
    for k in out:
        if pattern in k:
            if rows is None:
                out[k] *= factor
            else:
                out[k][rows(out[k]):] *= factor
    return out


def test_large_norm_recurrent_weights_do_not_overflow_the_static_scale(l4_forced):
    """Recurrent kernels 3x their Glorot range (saturating gates, |dh| growing along the chain): T = 250, the loss-scaled
    fp16 gradients stay finite, the guard never fires, parity holds."""
    from oracle import blstm as oblstm
    model, batch, canon, inp = _build('av-blstm', 3, 48000, seed=51)
    H = 250
    big = _scaled(canon, '/kernel', 3.0, rows=lambda k: k.shape[0] - H)              # the h rows of [x ; h]
    model.assign_vars(big)
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len']),
                                         big, 3)
    assert rel_l2(model.prediction.cpu().numpy(), outs['prediction']) < TOL
    _check_grads(model.canonical_gradients(), ograds, "av-blstm, 3x recurrent weights")
    model.train_op()
    assert model.engine.guard_state() == (0, 1.0)
    assert bool(torch.isfinite(model.engine.theta).all())


def test_overflow_guard_skips_and_recovers():
    """Head weights x 4e4: dY = dlogits . W_head^T exceeds 65504 in fp16 -> NaN weight gradients.  The guard must skip
    those steps (weights and Adam state untouched), halve the scale until the backward pass is finite, and the
    gradients at the reduced scale must match the oracle."""
    from oracle import blstm as oblstm
    # (a) learning rate 0: the weights cannot move, so the parity check at the backed-off scale is against the SAME weights
    model, batch, canon, inp = _build('av-blstm', 4, 9600, seed=52, starter_learning_rate=0.0)
    big = _scaled(canon, 'logits/weights', 4.0e4)
    model.assign_vars(big)
    theta0 = model.engine.theta.clone()
    model.train_op()                                   # overflows: skipped
    assert model.engine.guard_state() == (1, 0.5)
    assert float(model.engine.adam_m.abs().max()) == 0.0 and float(model.engine.adam_v.abs().max()) == 0.0
    for _ in range(12):
        model.train_op()
    skipped, scale = model.engine.guard_state()
    assert 1 <= skipped <= 10 and scale == 0.5 ** skipped
    assert float(model.engine.adam_v.abs().max()) > 0.0           # the later steps went through
    assert torch.equal(model.engine.theta, theta0)
    model.feed(dropout_rate=0.0)                       # same tensors, a fresh evaluation
    tsn, net_in = _oracle_inputs(batch, inp)
    outs, ograds = oblstm.loss_and_grads('si', dict(net_in=net_in, target=tsn, mask=batch['mask'], seq_len=batch['seq_len']),
                                         big, 3)
    # At these weights |prediction| ~ 1e4 and its fp16-level error ~ 8 covers 4e-4 of the bins' |prediction - target|
    # (6e-5 with ordinary weights): each such bin flips the sign of its +-1 L1 gradient, in ANY arithmetic narrower than
    # the oracle's.  The bound is therefore 2 x TOL here; the arithmetic of the backward pass at scale 1/2 is the same.
    _check_grads(model.canonical_gradients(), ograds, 'av-blstm at the backed-off scale %g' % scale, tol=2 * TOL)
    # (b) with a learning rate: the skipped step leaves the weights alone, training goes on afterwards
    model, batch, canon, inp = _build('av-blstm', 4, 9600, seed=52)
    model.assign_vars(big)
    theta0 = model.engine.theta.clone()
    model.train_op()
    assert model.engine.guard_state() == (1, 0.5) and torch.equal(model.engine.theta, theta0)
    for _ in range(12):
        model.train_op()
    assert bool(torch.isfinite(model.engine.theta).all()) and not torch.equal(model.engine.theta, theta0)


def test_frontend_bench_launch_2048_distinct_utterances_vs_cufft():
    """The front-end launch bench.py times (B = 2048 DISTINCT utterances, N = 48000, T = 250, with the landmark stream):
    all 131 M values of the normalised log-spectrogram against cuFFT in float64 on the same GPU (torch.fft.rfft of the
    frames of tf.signal.frame(pad_end=True) x the periodic Hann window, audio_processing.py:25-56 as restated in
    oracle/stft.py) within the 1e-5 relative tolerance, every utterance on its own; the mask application and the fp16
    time-major network input bit-exact given that spectrogram; the landmark columns against the oracle on a sample."""
    from oracle import video as ovideo
    B, N, T, F, fl, hop = 2048, 48000, 250, 257, 384, 192
    model, batch, canon, inp = _build('av-blstm', B, N, seed=21)
    dev = model.device
    got = model.target_spec_norm                                             # [B,T,F] fp32 (the training kernel)
    wav = torch.as_tensor(batch['wav'], device=dev).double()
    k = torch.arange(fl, device=dev, dtype=torch.float64)
    win = 0.5 - 0.5 * torch.cos(2.0 * np.pi * k / fl)
    mean = torch.as_tensor(np.asarray(batch['mean'], np.float64), device=dev)
    std = torch.as_tensor(np.asarray(batch['std'], np.float64), device=dev)
    xp = torch.nn.functional.pad(wav, (0, fl + hop * (T - 1) - N))
    worst, num, den = 0.0, 0.0, 0.0
    for b0 in range(0, B, 256):
        frames = xp[b0:b0 + 256].unfold(1, fl, hop)[:, :T] * win
        lin = torch.fft.rfft(frames, n=512, dim=-1).abs() + 1e-6
        ref = (torch.log(lin) - mean) / std
        g = got[b0:b0 + 256].double()
        d = g - ref
        num += float((d * d).sum())
        den += float((ref * ref).sum())
        # utterance by utterance on the linear magnitude (a single near-zero bin -- 131 M draws hold a few -- moves the LOG
        # of one value by 1e-2 whatever the transform's accuracy; the tolerance is on the spectra)
        dl = torch.exp(g * std + mean) - lin
        worst = max(worst, float((dl.flatten(1).norm(dim=1) / lin.flatten(1).norm(dim=1)).max()))
    assert (num / den) ** 0.5 < 1e-5 and worst < 1e-5, ((num / den) ** 0.5, worst)
    mask = torch.as_tensor(batch['mask'], device=dev).float()
    x = model.net_inputs                                                      # [B,T,393] fp32 view of the fp16 input
    assert torch.equal(x[:, :, :F].half(), (got * mask).half())               # mask application + fp16 rounding: exact
    assert bool((x[:, :, :F][mask == 0] == 0).all())
    for b in (0, 777, 2047):
        vid = ovideo.video_features(batch['landmarks'][b].astype(np.float64), T, batch['vmean'][b].astype(np.float64),
                                    batch['vstd'][b].astype(np.float64))
        assert rel_l2(x[b, :, F:].cpu().numpy(), vid) < 1e-3                  # fp16 copy of the z-normed motion vectors


def test_ctc_bench_launch_2048_utterances_vs_torch_ctc():
    """The CTC launch of the AV-MTL-SI bench step (B = 2048 utterances, T = 250 frames, 34 classes, labels padded to 50,
    blank = last class, unnormalised logits: models.py:1950-1953) against torch.nn.functional.ctc_loss in float64 on the
    same GPU (the implementation the float64 oracle itself is pinned to on small problems): every utterance's NLL and the
    whole gradient with respect to the logits."""
    from avsi_b200 import _lib
    lib = _lib.load()
    d = torch.device('cuda:0')
    T, B, C, Lmax, ldl, col0 = 250, 2048, 34, 50, 320, 257
    gen = torch.Generator(device='cpu').manual_seed(5)
    logits = (torch.randn(T * B, ldl, generator=gen) * 2).to(d)
    lab_len = torch.randint(12, 25, (B,), generator=gen, dtype=torch.int32)
    labels = torch.randint(0, C - 1, (B, Lmax), generator=gen, dtype=torch.int32)
    labels[0, 1] = labels[0, 0]                                      # a repeated label: the blank between them is mandatory
    seq = torch.full((B,), T, dtype=torch.int32)
    seq[1::7] = T - 9
    nll = torch.empty(B, device=d)
    dl = torch.zeros(T * B, ldl, dtype=torch.float16, device=d)
    ws = torch.empty(int(lib.avsi_ctc_workspace_bytes(B, T, Lmax)) // 4 + 4, device=d)
    tlab, tll, tsl = labels.to(d), lab_len.to(d), seq.to(d)
    _lib.check(lib.avsi_ctc_loss(_lib.ptr(logits), ldl, col0, C, _lib.ptr(tlab), Lmax, _lib.ptr(tll), _lib.ptr(tsl), B, T,
                                 8.0, None, _lib.ptr(nll), _lib.ptr(dl), ldl, col0, _lib.ptr(ws), _lib.stream_ptr()), 'ctc')
    torch.cuda.synchronize()
    lg = logits.view(T, B, ldl)[:, :, col0:col0 + C].double().clone().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(torch.log_softmax(lg, dim=2), tlab.long(), tsl.long(), tll.long(), blank=C - 1,
                                       reduction='none', zero_infinity=False)
    ref.sum().backward()
    assert torch.allclose(nll.double(), ref.detach(), rtol=2e-5, atol=1e-4)
    g = dl.view(T, B, ldl)[:, :, col0:col0 + C].double() / 8.0
    assert float((g - lg.grad).abs().max()) < 2e-3                   # fp16 storage of values in [-1, 1]
    assert rel_l2(g.cpu().numpy(), lg.grad.cpu().numpy()) < 1e-3
    assert bool((dl.view(T, B, ldl)[:, :, :col0] == 0).all())
