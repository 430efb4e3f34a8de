"""CPU oracle (test infrastructure only) of phase_reconstruction.refine_phase: the post-processing of
inference.py:143-154 with the exact consistency projection X <- mag . exp(i angle(STFT(iSTFT(X)))) where the reference
calls lws.run_lws (the `lws` C extension is absent from the image and from the reference tree: PARITY UNPINNED against
it; this restatement pins the CUDA path to the same arithmetic in float64)."""
import numpy as np

from . import stft as ostft


def refine_phase(enhanced, masks, n_iter=100, window_size=24, step_size=12):
    x = np.asarray(enhanced, np.float64)
    B, N = x.shape
    stft = ostft.get_stft(x, window_size=window_size, step_size=step_size)
    keep = np.zeros(stft.shape)
    t, f = min(masks.shape[1], keep.shape[1]), min(masks.shape[2], keep.shape[2])
    keep[:, :t, :f] = masks[:, :t, :f]
    mag = np.abs(stft)
    ang = np.angle(stft) * keep
    for _ in range(n_iter):
        wav = ostft.get_sources(mag, ang, num_samples=N, window_size=window_size, step_size=step_size)
        again = ostft.get_stft(wav, window_size=window_size, step_size=step_size)
        ang = np.angle(stft) * keep + np.angle(again) * (1.0 - keep)
    return ostft.get_sources(mag, ang, num_samples=N, window_size=window_size, step_size=step_size)


def inconsistency(wav, mag, window_size=24, step_size=12):
    """|| |STFT(wav)| - mag || / || mag ||: 0 for a waveform whose spectrogram has exactly the wanted magnitudes."""
    s = np.abs(ostft.get_stft(np.asarray(wav, np.float64), window_size=window_size, step_size=step_size))
    return float(np.linalg.norm(s - mag) / np.linalg.norm(mag))
