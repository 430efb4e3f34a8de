"""Oracle (test infrastructure): STFT front end in numpy, TF-1.x semantics.

Follows audio_processing.py:25-72 (get_stft / get_spectrogram /
get_log_mel_spectrogram), :145-164 (reconstruct_sources / get_sources),
masking.py:41-45 (mask application on the complex STFT) and
models.py:30-45 (normalise, mask, concat).  The TF ops behind them
(tf.contrib.signal.stft, inverse_stft, inverse_stft_window_fn,
tf.signal.linear_to_mel_weight_matrix) are restated from their documented
behaviour; the chain is pinned by the docs/files fixtures (see __init__).
"""
import numpy as np


def ms_to_samples(ms, sample_rate):
    # audio_processing.py:27-28
    return int(round(ms / 1e3 * sample_rate))


def hann_periodic(n, dtype=np.float64):
    # tf.signal.hann_window(periodic=True), the stft default window_fn
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def num_frames(n_samples, hop):
    # tf.signal.frame(pad_end=True): ceil(N / hop)
    return -(-n_samples // hop)


def frame_signal(x, frame_len, hop):
    """tf.signal.frame(x, frame_len, hop, pad_end=True) -> [B, T, frame_len]."""
    x = np.asarray(x)
    B, N = x.shape
    T = num_frames(N, hop)
    need = frame_len + hop * (T - 1)
    if need > N:
        x = np.concatenate([x, np.zeros((B, need - N), x.dtype)], axis=1)
    idx = hop * np.arange(T)[:, None] + np.arange(frame_len)[None, :]
    return x[:, idx]


def get_stft(sources, sample_rate=16000, window_size=25, step_size=10, n_fft=512,
             out_shape=(0, 0, 0), dtype=np.float64):
    """audio_processing.py:25-42.  Returns complex [B, T, n_fft//2+1]."""
    frame_len = ms_to_samples(window_size, sample_rate)
    hop = ms_to_samples(step_size, sample_rate)
    x = np.asarray(sources, dtype=dtype)
    frames = frame_signal(x, frame_len, hop) * hann_periodic(frame_len, dtype)
    # tf.signal.rfft(fft_length=n_fft) zero-pads (or crops) the END of each frame
    stfts = np.fft.rfft(frames.astype(np.float64), n=n_fft, axis=-1)
    if dtype == np.float32:
        stfts = stfts.astype(np.complex64)
    if not all(int(s) == 0 for s in out_shape):
        stfts = stfts[:out_shape[0], :out_shape[1], :out_shape[2]]
    return stfts


def get_spectrogram(stfts, power=1, log=False, out_shape=(0, 0, 0)):
    """audio_processing.py:45-56."""
    spec = np.abs(stfts)
    if power != 1:
        spec = spec ** power
    if log:
        spec = np.log(spec + 1e-6)
    if not all(int(s) == 0 for s in out_shape):
        spec = spec[:out_shape[0], :out_shape[1], :out_shape[2]]
    return spec


def hz_to_mel(f):
    return 1127.0 * np.log1p(np.asarray(f, dtype=np.float64) / 700.0)


def linear_to_mel_weight_matrix(num_mel_bins=80, num_spec_bins=257, sample_rate=16000,
                                lower_edge_hertz=125.0, upper_edge_hertz=7600.0):
    """tf.signal.linear_to_mel_weight_matrix (HTK mel, triangles linear in mel,
    DC bin row forced to zero).  Returns [num_spec_bins, num_mel_bins] float64."""
    nyquist = sample_rate / 2.0
    lin = np.linspace(0.0, nyquist, num_spec_bins)[1:]
    spec_mel = hz_to_mel(lin)[:, None]
    edges = np.linspace(hz_to_mel(lower_edge_hertz), hz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    lower, center, upper = edges[:-2][None, :], edges[1:-1][None, :], edges[2:][None, :]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(0.0, np.minimum(lower_slopes, upper_slopes))
    return np.concatenate([np.zeros((1, num_mel_bins)), w], axis=0)


def get_log_mel_spectrogram(spectrograms, sample_rate=16000, num_spec_bins=257, num_mel_bins=80,
                            lower_edge_freq=125, upper_edge_freq=7600, eps=1e-6):
    """audio_processing.py:59-72 (the out_shape slice there is a no-op)."""
    if upper_edge_freq is None:
        upper_edge_freq = sample_rate / 2
    m = linear_to_mel_weight_matrix(num_mel_bins, num_spec_bins, sample_rate, lower_edge_freq, upper_edge_freq)
    return np.log(np.tensordot(spectrograms, m.astype(spectrograms.dtype), axes=1) + eps)


def preemphasis(sources, alpha=0.95):
    """audio_processing.py:19-22: sources - alpha * [0, sources[:, :-1]]."""
    x = np.asarray(sources, np.float64)
    return x - alpha * np.concatenate([np.zeros((x.shape[0], 1)), x[:, :-1]], axis=1)


def get_mfcc(log_mel_spectrograms, num_mfccs=13):
    """audio_processing.py:74-81: tf.signal.mfccs_from_log_mel_spectrograms = DCT-II (tf.signal.dct type 2, norm=None:
    X_k = 2 sum_n x_n cos(pi k (2n+1) / (2N))) scaled by rsqrt(2 N), first num_mfccs coefficients."""
    x = np.asarray(log_mel_spectrograms, np.float64)
    N = x.shape[-1]
    n = np.arange(N)
    k = np.arange(N)[:, None]
    dct = 2.0 * np.cos(np.pi * k * (2 * n + 1) / (2.0 * N))          # [k, n]
    return (x @ dct.T / np.sqrt(2.0 * N))[..., :num_mfccs]


def delta(features, N=2):
    """audio_processing.py:84-93, line by line: one frame of SYMMETRIC padding per i, on the already padded tensor."""
    f = np.asarray(features, np.float64)
    denominator = 2 * sum(i ** 2 for i in range(1, N + 1))
    acc = np.zeros_like(f)
    padded = f
    for i in range(1, N + 1):
        padded = np.pad(padded, [(0, 0), (1, 1), (0, 0)], mode='symmetric')
        acc += i * (padded[:, i * 2:, :] - padded[:, :-i * 2, :])
    return acc / denominator


def add_delta_features(features, n_delta=2, N=2):
    """audio_processing.py:96-103."""
    full, cur = [np.asarray(features, np.float64)], np.asarray(features, np.float64)
    for _ in range(n_delta):
        cur = delta(cur, N)
        full.append(cur)
    return np.concatenate(full, axis=2)


def inverse_stft_window(frame_len, hop, dtype=np.float64):
    """tf.contrib.signal.inverse_stft_window_fn(hop)(frame_len) with a Hann forward window."""
    w = hann_periodic(frame_len)
    overlaps = -(-frame_len // hop)
    denom = np.concatenate([w * w, np.zeros(overlaps * hop - frame_len)])
    denom = denom.reshape(overlaps, hop).sum(0, keepdims=True)
    denom = np.tile(denom, (overlaps, 1)).reshape(-1)
    return (w / denom[:frame_len]).astype(dtype)


def reconstruct_sources(stfts, num_samples=0, sample_rate=16000, window_size=16, step_size=8):
    """audio_processing.py:145-157: inverse STFT (fft_length = next pow2 of frame_len),
    inverse window, overlap-add, optional slice to num_samples."""
    frame_len = ms_to_samples(window_size, sample_rate)
    hop = ms_to_samples(step_size, sample_rate)
    n_fft = 1 << (frame_len - 1).bit_length()
    frames = np.fft.irfft(stfts, n=n_fft, axis=-1)[..., :frame_len]
    frames = frames * inverse_stft_window(frame_len, hop)
    B, T, _ = frames.shape
    out = np.zeros((B, (T - 1) * hop + frame_len))
    for t in range(T):
        out[:, t * hop:t * hop + frame_len] += frames[:, t]
    if num_samples > 0:
        out = out[:, :num_samples]
    return out


def get_sources(mag, phase, num_samples=48000, sample_rate=16000, window_size=24, step_size=12):
    """audio_processing.py:160-164."""
    stfts = mag * np.cos(phase) + 1j * (mag * np.sin(phase))
    return reconstruct_sources(stfts, num_samples, sample_rate, window_size, step_size)


def mask_app_chain(target_wav, mask, oracle_phase=True, num_samples=48000):
    """masking.py:41-45 + :93-95: STFT -> x mask -> |.|, angle -> iSTFT -> int16."""
    wav = np.asarray(target_wav, dtype=np.float64)[None, :]
    stft = get_stft(wav, window_size=24, step_size=12, n_fft=512,
                    out_shape=(1,) + tuple(mask.shape))
    masked = stft * mask[None].astype(np.complex128)
    mag = np.abs(masked)
    phase = np.angle(stft) if oracle_phase else np.angle(masked)
    rec = get_sources(mag, phase, num_samples=num_samples)[0]
    return rec


def frontend(wav, mask, mean, std, video=None, dtype=np.float64):
    """models.py:30-45: target_spec_norm and net_inputs for input 'a' / 'av'.

    wav [B,N]; mask [B,T,F]; mean/std [F]; video [B,T,V] or None.
    """
    B, T, F = mask.shape
    stft = get_stft(wav, window_size=24, step_size=12, n_fft=512, out_shape=(B, T, F), dtype=dtype)
    spec = get_spectrogram(stft, log=True)
    tsn = (spec - mean) / std
    audio = tsn * mask
    net_in = audio if video is None else np.concatenate([audio, video], axis=2)
    return stft, tsn, net_in


def mean_std_accumulate(feats_list, masks_list=None):
    """audio_feat_preprocessing.py:76-115: running sum / square-sum in float64."""
    tot = None
    tot2 = None
    count = 0
    for i, feat in enumerate(feats_list):
        feat = np.asarray(feat, dtype=np.float64)
        if masks_list is not None:
            m = masks_list[i]
            feat = feat[:len(m), :m.shape[1]] * m
            count += int(m[:, 0].sum())
        else:
            count += len(feat)
        s, s2 = feat.sum(0), (feat ** 2).sum(0)
        tot = s if tot is None else tot + s
        tot2 = s2 if tot2 is None else tot2 + s2
    mean = tot / count
    std = np.sqrt(tot2 / count - mean ** 2)
    return mean, std
