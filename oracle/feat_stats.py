"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the statistics loop of compute_mean_std_features
(/root/reference/av_speech_inpainting/audio_feat_preprocessing.py:76-115): per file, features -> optional
`feat[:len(mask), :feat_dim] * mask` (:87-92) -> sum x, sum x^2 in float64 (:102-103), frame count =
sum(mask[:, 0]) or len(feat) (:104-107); mean = S/n, std = sqrt(S2/n - mean^2) (:113-114)."""
import numpy as np

from . import stft as ostft


def features_of(samples, ftype='spec', sample_rate=16000, window_size=25, step_size=10, num_mel_bins=80):
    st = ostft.get_stft(np.asarray(samples, np.float64)[None], sample_rate=sample_rate, window_size=window_size,
                        step_size=step_size)[0]
    if ftype == 'spec':
        return np.log(np.abs(st) + 1e-6)
    if ftype == 'fbanks':
        m = ostft.linear_to_mel_weight_matrix(num_mel_bins, 257, sample_rate, 125, 7600)
        return np.log((np.abs(st) ** 2) @ m + 1e-6)
    raise ValueError(ftype)


def mean_std(feature_list, masks=None):
    n = 0
    s = s2 = None
    for i, feat in enumerate(feature_list):
        feat = np.asarray(feat, np.float64)
        if masks is not None:
            mask = np.asarray(masks[i], np.float64)
            feat = feat[:len(mask), :mask.shape[1]] * mask
            n += int(mask[:, 0].sum())
        else:
            n += len(feat)
        s = feat.sum(0) if s is None else s + feat.sum(0)
        s2 = (feat ** 2).sum(0) if s2 is None else s2 + (feat ** 2).sum(0)
    mean = s / n
    return mean, np.sqrt(s2 / n - mean ** 2), n
