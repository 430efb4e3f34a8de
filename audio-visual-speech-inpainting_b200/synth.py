"""Seeded synthetic GRID-shaped batches (BASELINE.md section 3 / SURVEY.md 8d).

Host numpy only; used by tests, smoke() and bench.py (there is no network for datasets).
  wav        [B,N]   round(clip(N(0, 3500^2), +-32767)) as float32 (int16-valued, like dataset_reader.py:78)
  intervals  one gap per utterance cycling 100/200/400/800/1600 ms, onset from random.seed(30)
  landmarks  [B,75,136] base face shape + random walk (sigma 1 px), float32 pixel coordinates
  labels     [B,50] phone ids U{0..32}, lengths U{12..24}, zero padded (tfrecord_utils.py:99-101)
  mean/std   6.0 / 2.0 per bin (the docs fixtures have log-spectrum mean ~5.96, std ~2.1-2.5)
"""
import random

import numpy as np

GAPS_MS = (100, 200, 400, 800, 1600)


def make_batch(B, audio_len=48000, hop=192, F=257, V=136, video_frames=75, seed=0, n_labels=33, label_pad=50):
    rng = np.random.default_rng(seed)
    T = -(-audio_len // hop)
    wav = np.round(np.clip(rng.normal(0.0, 3500.0, (B, audio_len)), -32767, 32767)).astype(np.float32)
    pyrng = random.Random(30 + seed)
    dur_ms = audio_len / 16.0
    intervals = []
    for b in range(B):
        bins = int(np.around(T * GAPS_MS[b % len(GAPS_MS)] / dur_ms))
        bins = max(1, min(bins, T - 1))
        intervals.append([(pyrng.randint(0, T - bins), bins)])
    mask = np.ones((B, T, F), np.float32)
    for b, iv in enumerate(intervals):
        for o, l in iv:
            mask[b, o:o + l] = 0.0
    base = rng.uniform(100.0, 300.0, (1, 1, V))
    walk = np.cumsum(rng.normal(0.0, 1.0, (B, video_frames, V)), axis=1)
    landmarks = np.round(base + walk).astype(np.float32)
    vmean = np.zeros((B, V), np.float32)
    vstd = np.full((B, V), 0.3, np.float32)
    lab_len = rng.integers(12, 25, B).astype(np.int32)
    labels = np.zeros((B, label_pad), np.int32)
    for b in range(B):
        labels[b, :lab_len[b]] = rng.integers(0, n_labels, lab_len[b])
    return {
        'wav': wav, 'intervals': intervals, 'mask': mask, 'landmarks': landmarks, 'vmean': vmean, 'vstd': vstd,
        'labels': labels, 'lab_len': lab_len, 'seq_len': np.full(B, T, np.int32),
        'mean': np.full(F, 6.0, np.float32), 'std': np.full(F, 2.0, np.float32), 'T': T,
    }


def default_config(model='av-blstm', batch_size=8, audio_len=48000, net_dim=(250, 250, 250), ctc_loss=0.001, **over):
    """The shipped hyper-parameters (scripts/config/blstm*.config) as a dict, post check_trainconfiguration."""
    cfg = {
        'model': model, 'audio_feat_dim': 257, 'video_feat_dim': 136, 'audio_len': audio_len,
        'batch_size': batch_size, 'net_dim': list(net_dim), 'dropout_rate': 0.0, 'optimizer_type': 'adam',
        'starter_learning_rate': 0.001, 'lr_updating_steps': 10000, 'lr_decay': 1.0, 'l2': 0.0,
        'num_asr_labels': 34, 'ctc_loss': ctc_loss, 'seed': 0,
    }
    cfg.update(over)
    return cfg
