"""The oracle against vectors produced by TensorFlow itself (tests/golden/make_tf_golden.py).

The golden file does not exist yet: this image has no TensorFlow and no network (DESIGN.md section 2 records the
attempt), so every test here SKIPS until someone runs the generator in an environment that has TF.  When the file is
present these tests turn the "parity unpinned" rows (LSTM gate order / bias, CTC conventions, Adam epsilon-hat, mel matrix,
MFCC scaling, checkpoint bundle, TFRecord) into pinned ones."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

PATH = os.path.join(GOLDEN, 'tf_golden.npz')
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason='tests/golden/tf_golden.npz not generated (TensorFlow unavailable here)')


@pytest.fixture(scope='module')
def gold():
    return np.load(PATH)


def test_stft_mel_mfcc(gold):
    from oracle import stft as ostft
    st = ostft.get_stft(gold['wav'].astype(np.float64), window_size=24, step_size=12)
    assert rel_l2(np.stack([st.real, st.imag]), np.stack([gold['stft'].real, gold['stft'].imag])) < 1e-5
    assert np.abs(ostft.linear_to_mel_weight_matrix(80, 257, 16000, 125.0, 7600.0) - gold['mel']).max() < 1e-5
    assert rel_l2(ostft.get_mfcc(gold['logmel'].astype(np.float64), 13), gold['mfcc']) < 1e-5
    inv = ostft.reconstruct_sources(gold['stft'].astype(np.complex128), window_size=24, step_size=12)
    n = min(inv.shape[1], gold['inv_stft'].shape[1])
    assert rel_l2(inv[:, :n], gold['inv_stft'][:, :n]) < 1e-5


def test_ctc(gold):
    import torch
    from oracle import ctc as octc
    x = torch.tensor(gold['ctc_logits'].astype(np.float64), requires_grad=True)
    nll = octc.ctc_nll_torch(x, torch.from_numpy(gold['ctc_labels']).long(), torch.from_numpy(gold['ctc_lab_len']),
                             torch.from_numpy(gold['ctc_seq_len']))
    nll.sum().backward()
    assert np.allclose(nll.detach().numpy(), gold['ctc_nll'], rtol=1e-5, atol=1e-5)
    assert rel_l2(x.grad.numpy(), gold['ctc_dlogits']) < 1e-5


def test_adam(gold):
    from oracle import adam as oadam
    th, m, v = gold['adam_theta0'].astype(np.float64), np.zeros(11), np.zeros(11)
    for i in range(3):
        th, m, v = oadam.adam_tf_step(th, gold['adam_grads'][i].astype(np.float64), m, v, i + 1)
        assert np.abs(th - gold['adam_thetas'][i]).max() < 1e-6


def test_lstm_cell(gold):
    import torch
    from oracle import blstm as oblstm
    k = torch.tensor(gold['lstm_kernel'].astype(np.float64), requires_grad=True)
    b = torch.tensor(gold['lstm_bias'].astype(np.float64), requires_grad=True)
    h = oblstm.lstm_direction(torch.tensor(gold['lstm_x'].astype(np.float64)), k, b, False)
    (h * h).sum().backward()
    assert rel_l2(h.detach().numpy(), gold['lstm_h']) < 1e-5
    assert rel_l2(k.grad.numpy(), gold['lstm_dkernel']) < 1e-5 and rel_l2(b.grad.numpy(), gold['lstm_dbias']) < 1e-5


def test_checkpoint_and_tfrecord(gold):
    from avsi_b200 import tf_bundle, tfrecord_io as tio
    prefix = os.path.join(GOLDEN, 'tf_golden_ckpt')
    if os.path.exists(prefix + '.index'):
        got = tf_bundle.read_bundle(prefix)
        names = bytes(gold['ckpt_names']).decode().split('\n')
        assert sorted(got) == sorted(names)
        for n in names:
            assert np.array_equal(got[n], gold['ckpt/' + n]), n
    rec = os.path.join(GOLDEN, 'tf_golden.tfrecord')
    if os.path.exists(rec):
        data = list(tio.read_records(rec, verify=True))
        ctx, seq = tio.parse_sequence_example(data[0])
        assert int(ctx['sequence_length'][0]) == 5 and ctx['sample_path'][0] == b's1/bbaf2n'
        assert np.array_equal(np.asarray(ctx['target_audio_wav'], np.float32), gold['rec_wav'])
        assert np.array_equal(np.stack(seq['video_features']).astype(np.float32), gold['rec_video'])
        assert np.array_equal(np.stack(seq['mask']).astype(np.float32), gold['rec_mask'])
