"""Bit-reproducible reductions (include/avsi_b200.h, avsi_set_reduce_scratch): with the reduction scratch registered the
split-K weight-gradient GEMMs, the bias-gradient column sums, the loss sums and the BPTT kernels' bias gradient add their
partial results in a fixed order.  Held here: (1) each of them still equals its float64 reference and its atomic form,
(2) repeated launches give the same BITS, (3) whole AV-SI training runs repeated from the same weights end on identical
weights, on the mma.sync (B = 8) and the tcgen05 (B = 256) kernels."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda:0')


@pytest.fixture
def scratch_state():
    """Every test leaves the registry as the engine expects it (registered, minimum size or larger)."""
    from avsi_b200 import blstm
    yield blstm
    blstm.release_reduce_scratch()
    blstm.ensure_reduce_scratch()


@pytest.mark.parametrize('M,N,K,split', [(2048, 512, 64000, 1),      # CTA-pair kernel, 256 x 512 tiles, library-chosen split
                                         (1024, 256, 40000, 1),      # CTA-pair kernel, 256 x 256 tiles
                                         (291, 500, 3200, 6),        # one-tile kernel, caller's split, ragged M and N
                                         (64, 40, 640, 32),          # more splits asked for than k blocks (10): empty splits
                                         (130, 37, 2048, 4)])        # odd row length of C: the scalar tail of the reduction
def test_split_k_gemm_ordered_sum(M, N, K, split, scratch_state):
    blstm = scratch_state
    from avsi_b200 import _lib
    d = dev()
    gen = torch.Generator(device='cpu').manual_seed(M + N)
    A = (torch.randn(K, -(-M // 8) * 8, generator=gen) * 0.5).half().to(d)      # trans = 1: A [K, M], B [K, N]
    Bm = (torch.randn(K, -(-N // 8) * 8, generator=gen) * 0.5).half().to(d)
    C0 = torch.randn(M, N, generator=gen).to(d)
    ref = C0.double() + A[:, :M].double().t() @ Bm[:, :N].double()

    def run():
        C = C0.clone()
        blstm.gemm(_lib.ptr(A), A.shape[1], _lib.ptr(Bm), Bm.shape[1], _lib.ptr(C), N, None, M, N, K, 1, 2, split)
        torch.cuda.synchronize()
        return C
    blstm.ensure_reduce_scratch()
    outs = [run() for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(outs[0], o)
    assert rel_l2(outs[0].double().cpu().numpy(), ref.cpu().numpy()) < 1e-4        # fp32 accumulation over K
    # the atomic form (nothing registered) agrees to fp32 summation-order noise
    blstm.release_reduce_scratch()
    lib = _lib.load()
    C = C0.clone()
    _lib.check(lib.avsi_gemm_f16(_lib.ptr(A), A.shape[1], _lib.ptr(Bm), Bm.shape[1], _lib.ptr(C), N, None, M, N, K, 1, 2, split, 0,
                                 _lib.stream_ptr()), 'gemm')
    torch.cuda.synchronize()
    assert rel_l2(C.double().cpu().numpy(), ref.cpu().numpy()) < 1e-4
    assert rel_l2(C.cpu().numpy(), outs[0].cpu().numpy()) < 1e-5


def test_too_small_scratch_is_an_error_not_a_fallback(scratch_state):
    blstm = scratch_state
    from avsi_b200 import _lib
    lib = _lib.load()
    d = dev()
    blstm.release_reduce_scratch()
    small = torch.empty(int(lib.avsi_reduce_scratch_min_bytes()), dtype=torch.uint8, device=d)
    _lib.check(lib.avsi_set_reduce_scratch(_lib.ptr(small), small.numel(), _lib.stream_ptr()), 'set')
    M, N, K = 2048, 512, 64000
    assert int(lib.avsi_gemm_f16_scratch_bytes(M, N, K, 1, 2, 1)) > small.numel()
    A = torch.zeros(K, M, dtype=torch.float16, device=d)
    Bm = torch.zeros(K, N, dtype=torch.float16, device=d)
    C = torch.zeros(M, N, device=d)
    rc = lib.avsi_gemm_f16(_lib.ptr(A), M, _lib.ptr(Bm), N, _lib.ptr(C), N, None, M, N, K, 1, 2, 1, 0, _lib.stream_ptr())
    assert rc != 0 and b'scratch too small' in lib.avsi_last_error()
    with pytest.raises(_lib.AvsiError):
        _lib.check(lib.avsi_set_reduce_scratch(_lib.ptr(small), 1024, _lib.stream_ptr()), 'set')      # below the minimum
    _lib.check(lib.avsi_set_reduce_scratch(None, 0, None), 'unset')


def test_colsum_and_loss_sums_ordered(scratch_state):
    blstm = scratch_state
    from avsi_b200 import _lib
    lib = _lib.load()
    d = dev()
    blstm.ensure_reduce_scratch()
    gen = torch.Generator(device='cpu').manual_seed(7)
    for rows, ld, c0, nc in ((50000, 320, 0, 257), (1, 64, 8, 17), (333, 2048, 1024, 1024), (200000, 320, 0, 291)):
        X = torch.randn(rows, ld, generator=gen).half().to(d)
        outs = []
        for _ in range(3):
            out = torch.ones(nc, device=d)
            _lib.check(lib.avsi_colsum_f16(_lib.ptr(X), ld, rows, c0, nc, _lib.ptr(out), _lib.stream_ptr()), 'colsum')
            torch.cuda.synchronize()
            outs.append(out)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
        ref = X[:, c0:c0 + nc].double().sum(0).cpu().numpy() + 1.0
        assert np.abs(outs[0].cpu().numpy() - ref).max() < 2e-3 * max(1.0, np.sqrt(rows) / 10), (rows, ld, c0, nc)
    # masked-L1 sums: B * T rows over more blocks than one wave
    B, T, F, ldl = 64, 250, 257, 320
    logits = torch.randn(T * B, ldl, generator=gen).to(d)
    target = torch.randn(B, T, F, generator=gen).to(d)
    mask = (torch.rand(B, T, F, generator=gen) > 0.3).float().to(d)
    seq = torch.full((B,), T, dtype=torch.int32, device=d)
    seq[3] = 17
    res = []
    for _ in range(3):
        sums = torch.zeros(8, dtype=torch.float64, device=d)
        pred = torch.empty(B, T, F, device=d)
        dl = torch.zeros(T * B, ldl, dtype=torch.float16, device=d)
        _lib.check(lib.avsi_masked_l1(_lib.ptr(logits), ldl, _lib.ptr(target), _lib.ptr(mask), _lib.ptr(seq), B, T, F, 1, 4.0, None,
                                      _lib.ptr(sums), _lib.ptr(pred), _lib.ptr(dl), ldl, _lib.stream_ptr()), 'l1')
        torch.cuda.synchronize()
        res.append((sums, pred, dl))
    for r in res[1:]:
        assert torch.equal(res[0][0], r[0]) and torch.equal(res[0][1], r[1]) and torch.equal(res[0][2], r[2])
    sm = (torch.arange(T, device=d)[None, :] < seq[:, None]).double()[:, :, None]
    x = logits.view(T, B, ldl)[:, :, :F].transpose(0, 1).double()
    p = (target.double() * mask.double() + x * (1 - mask.double())) * sm
    ad = (target.double() - p).abs()
    want = torch.stack([(ad * (1 - mask)).sum(), (1 - mask.double()).sum(), (ad * mask).sum(), mask.double().sum(), ad.sum(),
                        torch.tensor(float(B * T * F), dtype=torch.float64, device=d)])
    assert torch.allclose(res[0][0][:6], want, rtol=1e-6)
    # a second call accumulates (the contract is +=), and the tickets reset themselves
    sums = res[0][0].clone()
    _lib.check(lib.avsi_masked_l1(_lib.ptr(logits), ldl, _lib.ptr(target), _lib.ptr(mask), _lib.ptr(seq), B, T, F, 1, 4.0, None,
                                  _lib.ptr(sums), None, None, ldl, _lib.stream_ptr()), 'l1')
    torch.cuda.synchronize()
    assert torch.allclose(sums[:6], 2 * want, rtol=1e-6)


@pytest.mark.parametrize('T,B,kernels', [(9, 40, 'mma'), (12, 300, 'l4'), (5, 129, 'l4')])
def test_bptt_bias_gradient_ordered(T, B, kernels):
    """avsi_lstm_bwd with its scratch argument: per-tile sums + ordered reduction == the atomic form, and repeats bit for bit
    (the ragged last tile included)."""
    from avsi_b200 import _lib
    lib = _lib.load()
    d = dev()
    gen = torch.Generator(device='cpu').manual_seed(T * B)
    Mp = -(-T * B // 32) * 32
    g0 = torch.randn(Mp, 2048, generator=gen).half().to(d)
    whh = (torch.randn(2048, 256, generator=gen) * 0.05).half().to(d)
    whhT = whh.t().contiguous()
    bias = torch.zeros(2048, device=d)
    dy = torch.randn(Mp, 512, generator=gen).half().to(d)
    scratch = torch.empty(int(lib.avsi_lstm_bwd_scratch_bytes(B)) // 4, device=d)
    assert scratch.numel() >= -(-B // 16) * 2048
    _lib.set_env(AVSI_LSTM_FWD=kernels, AVSI_LSTM_BWD=kernels)
    try:
        outs = []
        for sc in (scratch, scratch, None):
            gates = g0.clone()
            y = torch.zeros(T * B, 512, dtype=torch.float16, device=d)
            cst = torch.zeros(Mp, 512, device=d)
            dbias = torch.full((2048,), 0.25, device=d)
            scratch.fill_(float('nan'))
            _lib.check(lib.avsi_lstm_fwd(_lib.ptr(gates), _lib.ptr(whh), _lib.ptr(bias), _lib.ptr(y), _lib.ptr(cst), T, B, 0,
                                         _lib.stream_ptr()), 'lstm_fwd')
            _lib.check(lib.avsi_lstm_bwd(_lib.ptr(gates), _lib.ptr(whhT), _lib.ptr(cst), _lib.ptr(dy), _lib.ptr(dbias),
                                         _lib.ptr(sc), T, B, _lib.stream_ptr()), 'lstm_bwd')
            torch.cuda.synchronize()
            outs.append((gates, dbias))
    finally:
        _lib.set_env(AVSI_LSTM_FWD=None, AVSI_LSTM_BWD=None)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[2][0])
    assert torch.isfinite(outs[0][1]).all() and float((outs[0][1] - 0.25).abs().max()) > 0
    assert rel_l2((outs[0][1] - 0.25).cpu().numpy(), (outs[2][1] - 0.25).cpu().numpy()) < 1e-5


@pytest.mark.parametrize('model_name,B,audio_len', [('av-blstm', 8, 9600), ('av-blstm', 256, 4800), ('a-blstm', 130, 2400),
                                                    ('av-blstm-ssnn-ctc', 64, 9600)])
def test_training_runs_repeat_bit_for_bit(model_name, B, audio_len):
    """Two training runs from the same weights over the same three batches: identical weights, Adam moments and
    losses, on the small-batch (mma.sync) and the large-batch (tcgen05, CTA-pair split-K GEMM) kernels, and for the
    multi-task model (ordered CTC posterior sums, exact hole count)."""
    from test_gpu_model import _build
    from avsi_b200 import av_sync, blstm, synth
    assert blstm.deterministic()
    batches = [synth.make_batch(B, audio_len=audio_len, seed=50 + i) for i in range(3)]
    runs = []
    for _ in range(2):
        model, _, _, _ = _build(model_name, B, audio_len, seed=4)
        losses = []
        for b in batches:
            video = av_sync.video_pipeline(b['landmarks'], b['T'], b['vmean'], b['vstd'])
            f = dict(target_sources=b['wav'], masks=b['mask'], sequence_lengths=b['seq_len'], video_features=video)
            if model.MTL:
                f.update(labels=b['labels'], labels_lengths=b['lab_len'])
            model.feed(**f)
            model.train_op()
            losses.append(float(model.loss))
        torch.cuda.synchronize()
        runs.append((model.engine.theta.clone(), model.engine.adam_m.clone(), model.engine.adam_v.clone(), losses))
        del model
    assert torch.equal(runs[0][0], runs[1][0]), float((runs[0][0] - runs[1][0]).abs().max())
    assert torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])
    assert runs[0][3] == runs[1][3] and all(np.isfinite(runs[0][3]))


def test_atomic_mode_of_the_host_still_trains(monkeypatch, scratch_state):
    """AVSI_DETERMINISTIC=0: nothing is registered, the BPTT kernels get no scratch, the step runs on the atomics of
    round 1 -- same gradient to summation-order noise."""
    blstm = scratch_state
    from test_gpu_model import _build
    model, _, _, _ = _build('av-blstm', 6, 4800, seed=8)
    g_ordered = {k: np.array(v) for k, v in model.canonical_gradients().items()}
    monkeypatch.setenv('AVSI_DETERMINISTIC', '0')
    blstm.release_reduce_scratch()
    other, _, _, _ = _build('av-blstm', 6, 4800, seed=8)
    assert other.engine.workspace(other._fed['masks'].shape[1], 6)['scratch'] is None
    g_atomic = other.canonical_gradients()
    assert torch.cuda.current_device() not in blstm._REDUCE
    a = np.concatenate([g_ordered[k].ravel() for k in sorted(g_ordered)])
    b = np.concatenate([g_atomic[k].ravel() for k in sorted(g_ordered)])
    assert np.linalg.norm(a) > 0 and rel_l2(b, a) < 1e-5


def test_bench_shape_steps_repeat_bit_for_bit():
    """The exact bench launch (B = 2048, T = 250: 32 recurrence clusters, 9- and 37-way split-K dW GEMMs, 16 bias-gradient
    tiles): two steps from the same weights, twice -- identical weights and losses."""
    from test_gpu_model import _build
    from avsi_b200 import blstm
    assert blstm.deterministic()
    runs = []
    for _ in range(2):
        model, _, _, _ = _build('av-blstm', 2048, 48000, seed=12)
        losses = []
        for _s in range(2):
            model.feed()
            model.train_op()
            losses.append(float(model.loss))
        torch.cuda.synchronize()
        runs.append((model.engine.theta.clone(), losses))
        del model
        torch.cuda.empty_cache()
    assert torch.equal(runs[0][0], runs[1][0]), float((runs[0][0] - runs[1][0]).abs().max())
    assert runs[0][1] == runs[1][1] and all(np.isfinite(runs[0][1]))
