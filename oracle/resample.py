"""CPU oracle (test infrastructure only) for downsampling(samples, sample_rate, downsample_rate),
audio_processing.py:9-16.

The arithmetic lives in a third-party dependency, ``scipy.signal.resample`` (requirements.txt: scipy, unpinned;
scipy/signal/_signaltools.py ``resample``, real-input branch).  Its published algorithm, restated with numpy FFTs:
rfft of the n_in samples; keep the first N/2 + 1 bins, N = min(num, n_in); when N is even the bin N/2 is doubled
(downsampling: it absorbs the component at -N/2) or halved (upsampling: it is split over +-N/2); irfft of length num;
scale by num / n_in.  PINNED: tests/test_third_party_pins_cpu.py compares it with scipy.signal.resample itself (scipy
is in the image) on even / odd lengths in both directions.
"""
import numpy as np


def resample(x, num):
    x = np.asarray(x, dtype=np.float64)
    nx = x.shape[-1]
    X = np.fft.rfft(x, axis=-1)
    N = min(num, nx)
    Y = np.zeros(x.shape[:-1] + (num // 2 + 1,), dtype=np.complex128)
    Y[..., :N // 2 + 1] = X[..., :N // 2 + 1]
    if N % 2 == 0:
        if num < nx:
            Y[..., N // 2] *= 2.0
        elif nx < num:
            Y[..., N // 2] *= 0.5
    return np.fft.irfft(Y, num, axis=-1) * (float(num) / float(nx))


def downsampling(samples, sample_rate, downsample_rate):
    """audio_processing.py:9-16."""
    secs = len(samples) / float(sample_rate)
    num_samples = int(downsample_rate * secs)
    if sample_rate != downsample_rate:
        return resample(samples, num_samples)
    return samples
