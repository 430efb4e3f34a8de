"""Oracle vs the reference's only known-answer fixtures and extracted-function golden vectors."""
import json
import os
import random

import numpy as np

from oracle import maskgen, stft, video


def test_stft_mask_istft_chain_matches_docs_fixtures(golden_dir):
    fx = np.load(os.path.join(golden_dir, 'docs_fixtures.npz'))
    for key in ('800ms_ex1', '800ms_ex2', '1600ms_ex1', '1600ms_ex2'):
        target, masked = fx[key + '_target'], fx[key + '_masked']
        a, b = fx[key + '_range']
        mask = np.ones((250, 257))
        mask[a:b] = 0
        rec = stft.mask_app_chain(target, mask, oracle_phase=True).astype(np.int16)   # masking.py:93-95
        err = np.abs(rec.astype(int) - masked.astype(int))
        assert err.max() <= 1, (key, err.max())
        # the fixture is sharp: an off-by-one mask boundary is off by >= 100 LSB
        mask2 = np.ones((250, 257))
        mask2[a + 1:b + 1] = 0
        rec2 = stft.mask_app_chain(target, mask2).astype(np.int16)
        assert np.abs(rec2.astype(int) - masked.astype(int)).max() > 100


def test_stft_basic_properties():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 48000))
    s = stft.get_stft(x, window_size=24, step_size=12, n_fft=512)
    assert s.shape == (2, 250, 257)
    # frame 3 by direct DFT
    w = stft.hann_periodic(384)
    fr = np.concatenate([x[0, 3 * 192:3 * 192 + 384] * w, np.zeros(128)])
    assert np.allclose(s[0, 3], np.fft.rfft(fr), atol=1e-9)
    # last frame is zero padded at the end (pad_end=True)
    last = np.zeros(384)
    last[:192] = x[1, 249 * 192:]
    assert np.allclose(s[1, 249], np.fft.rfft(np.concatenate([last * w, np.zeros(128)])), atol=1e-9)
    # STFT -> iSTFT reconstructs the interior exactly (inverse_stft_window_fn)
    rec = stft.reconstruct_sources(s, num_samples=48000, window_size=24, step_size=12)
    assert np.allclose(rec[:, 192:47808], x[:, 192:47808], atol=1e-8)


def test_mel_matrix_properties():
    m = stft.linear_to_mel_weight_matrix(80, 257, 16000, 125, 7600)
    assert m.shape == (257, 80)
    assert np.all(m[0] == 0) and np.all(m >= 0) and m.max() <= 1.0
    assert np.all((m > 0).sum(axis=1) <= 2)            # every bin belongs to at most two triangles
    peaks = m.argmax(axis=0)
    assert np.all(np.diff(peaks) >= 0)                 # band centres do not decrease


def test_maskgen_matches_reference_function(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, 'maskgen_cases.json')))
    gold = np.load(os.path.join(golden_dir, 'maskgen.npz'))
    last = None
    for c in cases:
        key = (c['seed'], c['n_max'], c['mean'], c['std'])
        if key != last:
            random.seed(c['seed'])
            last = key
        mask, cov, n_intr, iv = maskgen.get_intrusions_mask(257, 250, c['mean'], c['std'], c['n_max'])
        col = np.unpackbits(gold[c['key']])[:250]
        assert n_intr == c['n_intr'] and cov == c['cov']
        assert np.array_equal(mask[:, 0].astype(np.uint8), col)
        assert np.all(mask == mask[:, :1])


def test_motion_vector_matches_reference_function(golden_dir):
    g = np.load(os.path.join(golden_dir, 'motion_vector.npz'))
    assert np.array_equal(video.get_motion_vector(g['landmarks'], 1), g['delta1'])
    assert np.array_equal(video.get_motion_vector(g['landmarks'], 2), g['delta2'])


def test_inc_fps_is_clamped_lerp():
    rng = np.random.default_rng(1)
    lm = rng.uniform(0, 100, (75, 136))
    up = video.inc_fps(lm, 250)
    assert up.shape == (250, 136)
    assert np.allclose(up[0], lm[0]) and np.allclose(up[10], lm[3])      # 10 * 75/250 = 3
    assert np.allclose(up[5], 0.5 * (lm[1] + lm[2]))                      # 1.5
    assert np.allclose(up[249], 0.3 * lm[74] + 0.7 * lm[74])               # y = 74.7 clamps to the last frame
    short = video.sync_audio_visual_features(250, lm[:72], tot_frames=75, min_frames=70)
    assert np.allclose(short[0], lm[0]) and short.shape == (250, 136)
    assert video.sync_audio_visual_features(250, lm[:60], tot_frames=75, min_frames=70) is None
