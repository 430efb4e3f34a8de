"""Oracle (test infrastructure): random intrusion masks.

Restates dataset_generator.py:11-48 (get_intrusions_mask).  The draw order from
Python's ``random`` module (MT19937: randint, gauss, randint..., shuffle,
randint...) is part of the contract: with the same seed the mask is bit-exact.
Returns the dense mask like the reference plus the (onset, length) intervals.
"""
import random

import numpy as np


def get_intrusions_mask(frame_dim, spec_len, cov_mean, cov_std, n_max_intr, min_intr_len=3, rng=random):
    n_intr = rng.randint(1, n_max_intr)                                   # :13
    cov = max(min_intr_len * n_intr / spec_len, min(rng.gauss(cov_mean, cov_std), 0.8))   # :16
    mask_bins = int(np.around(spec_len * cov))                            # :17
    true_cov = mask_bins / spec_len

    decay = np.exp(-(n_intr - 1) / 6)
    lens = []
    for i in range(n_intr):                                               # :21-28
        if i == n_intr - 1:
            lens.append(mask_bins - sum(lens))
        else:
            room = mask_bins - sum(lens) - min_intr_len * (n_intr - i - 1)
            lens.append(rng.randint(min_intr_len, max(min_intr_len, int(room * decay))))
    rng.shuffle(lens)                                                     # :29

    onsets = []
    for i in range(n_intr):                                               # :32-41
        if i == 0 and n_intr == 1:
            onsets.append(rng.randint(0, spec_len - mask_bins))
        elif i == 0:
            onsets.append(rng.randint(0, spec_len - mask_bins - (n_intr - 1)) // 2)
        else:
            prev_end = onsets[-1] + lens[i - 1] + 1
            if i == n_intr - 1:
                onsets.append(rng.randint(onsets[-1], prev_end + spec_len - lens[i]))
            else:
                onsets.append(rng.randint(prev_end, (prev_end + spec_len - sum(lens[i:]) - (n_intr - i - 1)) // 2))

    mask = np.ones([spec_len, frame_dim])
    for o, l in zip(onsets, lens):                                        # :44-46
        mask[o:o + l] = 0
    return mask, true_cov, n_intr, list(zip(onsets, lens))
