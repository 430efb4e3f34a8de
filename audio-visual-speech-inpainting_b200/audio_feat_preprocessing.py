"""Feature statistics / feature dumps: the reference's audio_feat_preprocessing.py on the B200 hot path.

compute_mean_std_features(...) keeps the reference signature (audio_feat_preprocessing.py:23-24) and on-disk
contract (`<out_prefix>_mean.npy`, `<out_prefix>_std.npy` in `audio_folder`, optional per-sample `.npy` dumps,
`mask.npy` per sample for apply_mask).  Instead of one TF session run per file, files of equal length are
batched through the fused front end and the float64 sums of :102-107 are accumulated on the GPU
(avsi_feature_stats).  `mfcc`, delta features and pre-emphasis are thin kernels on top (csrc/features_extra.cu).
"""
import os
from glob import glob

import numpy as np
import torch

from . import _lib
from . import audio_processing as ap


class FeatureStats(object):
    """Running sum x, sum x^2 (float64, device) and frame count over batches of features."""

    def __init__(self, feat_dim, device='cuda'):
        self.F = int(feat_dim)
        self.device = torch.device(device)
        self.sum = torch.zeros(self.F, dtype=torch.float64, device=self.device)
        self.sumsq = torch.zeros(self.F, dtype=torch.float64, device=self.device)
        self.count = torch.zeros(1, dtype=torch.float64, device=self.device)

    def update(self, feats, masks=None):
        """feats [..., F] f32 CUDA (any leading shape); masks like feats (feat * mask, count += mask[..., 0])."""
        lib = _lib.load()
        x = feats.reshape(-1, feats.shape[-1]).contiguous()
        if x.shape[1] < self.F:
            raise _lib.AvsiError('features have %d columns, statistics want %d' % (x.shape[1], self.F))
        m = None
        if masks is not None:
            m = masks.to(device=x.device, dtype=torch.float32).reshape(-1, masks.shape[-1]).contiguous()
            if m.shape[0] != x.shape[0] or m.shape[1] < self.F:
                raise _lib.AvsiError('mask shape does not match the features')
        with _lib.span('feature_stats', nbytes=x.shape[0] * self.F * 4 * (2 if m is not None else 1)):
            _lib.check(lib.avsi_feature_stats(_lib.ptr(x), x.shape[1], _lib.ptr(m), m.shape[1] if m is not None else 0,
                                              x.shape[0], self.F, _lib.ptr(self.sum), _lib.ptr(self.sumsq),
                                              _lib.ptr(self.count), _lib.stream_ptr()), 'avsi_feature_stats')
        return self

    def finalize(self):
        """(mean, std, frame_count) as float64 numpy: std = sqrt(E[x^2] - mean^2) (audio_feat_preprocessing.py:113-114)."""
        n = float(self.count.item())
        mean = (self.sum / n).cpu().numpy()
        std = np.sqrt((self.sumsq / n).cpu().numpy() - mean ** 2)
        return mean, std, int(round(n))


def _features(batch, ftype, sample_rate, window_size, step_size, num_mel_bins, num_mfcc=13, preemph=0, delta=0):
    """batch [B,N] f32 CUDA -> [B,T,F] features of the given type (audio_feat_preprocessing.py:36-62)."""
    frame_len, hop = ap.ms_to_samples(window_size, sample_rate), ap.ms_to_samples(step_size, sample_rate)
    if preemph > 0:
        batch = ap.preemphasis(batch, alpha=preemph)
    if ftype == 'spec':
        feats = ap.fused_features(batch, frame_len, hop, log=True, want_spec=True)['spec']
    else:
        feats = ap.log_mel_features(batch, sample_rate, window_size, step_size, num_mel_bins)
        if ftype == 'mfcc':
            feats = ap.get_mfcc(feats, num_mfcc)
    if delta > 0:
        feats = ap.add_delta_features(feats, n_delta=delta, N=2)
    return feats


def compute_mean_std_features(audio_folder, file_prefix, out_prefix, type='spec', sample_rate=16e3, n_fft=512,
                              window_size=25, step_size=10, preemph=0, num_mel_bins=80, num_mfcc=13, delta=0,
                              apply_mask=False, save_feat=False, file_ext='wav', batch_size=64, device='cuda'):
    from scipy.io import wavfile
    if type not in ('spec', 'fbanks', 'mfcc'):
        # 'stft' included: the reference falls through to this exit for it as well (audio_feat_preprocessing.py:44-61)
        print('Type must be "stft", "spec", "fbanks" or "mfcc". Closing...')
        raise SystemExit(1)
    if n_fft != 512:
        raise _lib.AvsiError('only n_fft = 512 is supported')
    sample_rate = int(sample_rate)
    audio_sample_dirs = sorted(d for d in glob(os.path.join(audio_folder, '*')) if os.path.isdir(d))
    feat_dim = {'spec': 257, 'fbanks': num_mel_bins, 'mfcc': num_mfcc}[type] * (delta + 1)
    stats = FeatureStats(feat_dim, device)
    print('Computing features...')
    # files of equal length share a batch (GRID: every utterance has 48000 samples)
    by_len = {}
    for d in audio_sample_dirs:
        rate, samples = wavfile.read(os.path.join(d, file_prefix + '.' + file_ext))
        samples = np.asarray(ap.downsampling(samples, rate, sample_rate), np.float32)
        by_len.setdefault(len(samples), []).append((d, samples))
    for n, items in sorted(by_len.items()):
        for i in range(0, len(items), batch_size):
            chunk = items[i:i + batch_size]
            wav = torch.from_numpy(np.stack([s for _, s in chunk])).to(device)
            feats = _features(wav, type, sample_rate, window_size, step_size, num_mel_bins, num_mfcc, preemph, delta)
            if apply_mask:
                for k, (d, _) in enumerate(chunk):
                    mask = torch.from_numpy(np.load(os.path.join(d, 'mask.npy')).astype(np.float32)).to(device)
                    f = feats[k, :mask.shape[0], :mask.shape[1]]       # discard last frequency bins and last frames
                    stats.update(f, mask)
                    if save_feat:
                        np.save(os.path.join(audio_folder, os.path.basename(d), file_prefix + '.npy'), (f * mask).cpu().numpy())
            else:
                stats.update(feats)
                if save_feat:
                    for k, (d, _) in enumerate(chunk):
                        np.save(os.path.join(audio_folder, os.path.basename(d), file_prefix + '.npy'), feats[k].cpu().numpy())
    print('done. Audio files processed:', len(audio_sample_dirs))
    print('Computing mean and standard deviation of features...')
    feat_mean, feat_std, frame_count = stats.finalize()
    print('Total number of frames:', frame_count)
    print('done.')
    np.save(os.path.join(audio_folder, out_prefix + '_mean.npy'), feat_mean)
    np.save(os.path.join(audio_folder, out_prefix + '_std.npy'), feat_std)
    print('Normalization data files saved.')
    return feat_mean, feat_std


def save_features(audio_folder, type='spec', sample_rate=16e3, n_fft=512, window_size=25, step_size=10, preemph=0,
                  num_mel_bins=80, num_mfcc=13, delta=0, file_ext='wav', device='cuda'):
    """audio_feat_preprocessing.py:130-196: features of every `<audio_folder>/*.<ext>` saved next to it as `.npy`."""
    from scipy.io import wavfile
    if type not in ('stft', 'spec', 'fbanks', 'mfcc'):
        print('Type must be "spec", "fbanks" or "mfcc". Closing...')
        raise SystemExit(1)
    if n_fft != 512:
        raise _lib.AvsiError('only n_fft = 512 is supported')
    sample_rate = int(sample_rate)
    files = sorted(glob(os.path.join(audio_folder, '*.' + file_ext)))
    print('Computing and saving features...')
    for audio_file in files:
        rate, samples = wavfile.read(audio_file)
        samples = np.asarray(ap.downsampling(samples, rate, sample_rate), np.float32)
        wav = torch.from_numpy(samples[None]).to(device)
        if type == 'stft':
            if preemph > 0:
                wav = ap.preemphasis(wav, alpha=preemph)
            feat = ap.get_stft(wav, sample_rate, window_size, step_size, n_fft)
            if delta > 0:
                raise _lib.AvsiError('delta features of a complex STFT are not defined')
        else:
            feat = _features(wav, type, sample_rate, window_size, step_size, num_mel_bins, num_mfcc, preemph, delta)
        np.save(os.path.splitext(audio_file)[0] + '.npy', feat[0].cpu().numpy())
    print('done. Audio files processed:', len(files))
