"""Random intrusion masks (dataset_generator.py:11-48).

The integer part (how many intrusions, their lengths and onsets) is drawn on the host from Python's
``random`` module in the reference's call order -- randint, gauss, randint..., shuffle, randint... --
so a seeded run is bit-exact; the dense [T, F] mask is expanded on the GPU from the integer
(onset, length) intervals by the ``avsi_expand_mask`` kernel.
"""
import random

import numpy as np
import torch

from . import _lib


def draw_intrusions(spec_len, cov_mean, cov_std, n_max_intr, min_intr_len=3, rng=random):
    """Integer half of get_intrusions_mask: returns ([(onset, length), ...], true_cov, n_intr)."""
    n_intr = rng.randint(1, n_max_intr)
    cov = max(min_intr_len * n_intr / spec_len, min(rng.gauss(cov_mean, cov_std), 0.8))
    mask_bins = int(np.around(spec_len * cov))
    true_cov = mask_bins / spec_len
    decay = np.exp(-(n_intr - 1) / 6)
    lens = []
    for i in range(n_intr):
        if i == n_intr - 1:
            lens.append(mask_bins - sum(lens))
        else:
            room = mask_bins - sum(lens) - min_intr_len * (n_intr - i - 1)
            lens.append(rng.randint(min_intr_len, max(min_intr_len, int(room * decay))))
    rng.shuffle(lens)
    onsets = []
    for i in range(n_intr):
        if n_intr == 1:
            onsets.append(rng.randint(0, spec_len - mask_bins))
        elif i == 0:
            onsets.append(rng.randint(0, spec_len - mask_bins - (n_intr - 1)) // 2)
        else:
            after_prev = onsets[-1] + lens[i - 1] + 1
            if i == n_intr - 1:
                onsets.append(rng.randint(onsets[-1], after_prev + spec_len - lens[i]))
            else:
                onsets.append(rng.randint(after_prev, (after_prev + spec_len - sum(lens[i:]) - (n_intr - i - 1)) // 2))
    return list(zip(onsets, lens)), true_cov, n_intr


def expand_masks(intervals, spec_len, frame_dim, device='cuda'):
    """[[(onset, length), ...] per sample] -> dense mask [B, spec_len, frame_dim] f32 on the GPU."""
    lib = _lib.load()
    B = len(intervals)
    K = max(1, max(len(iv) for iv in intervals))
    arr = np.zeros((B, K, 2), np.int32)
    for b, iv in enumerate(intervals):
        for k, (o, l) in enumerate(iv):
            arr[b, k] = (o, l)
    iv_dev = torch.from_numpy(arr).to(device)
    mask = torch.empty(B, spec_len, frame_dim, dtype=torch.float32, device=device)
    _lib.check(lib.avsi_expand_mask(_lib.ptr(iv_dev), B, K, spec_len, frame_dim, _lib.ptr(mask), _lib.stream_ptr()),
               'avsi_expand_mask')
    return mask


def get_intrusions_mask(frame_dim, spec_len, cov_mean, cov_std, n_max_intr, min_intr_len=3):
    """Reference signature (dataset_generator.py:11): returns (mask [spec_len, frame_dim] numpy,
    true_mask_cov, n_intr).  Draws from the global ``random`` state like the reference."""
    iv, true_cov, n_intr = draw_intrusions(spec_len, cov_mean, cov_std, n_max_intr, min_intr_len)
    mask = expand_masks([iv], spec_len, frame_dim)[0].cpu().numpy().astype(np.float64)
    return mask, true_cov, n_intr
