"""Audio feature ops with the reference's names and signatures (audio_processing.py:9-184),
computed by the fused sm_100a front-end kernels.

Tensors are torch CUDA tensors instead of TF graph tensors; shapes, argument meaning and the
millisecond -> sample conversion (``int(round(ms / 1e3 * sr))``, audio_processing.py:27-28) are
the reference's.  ``fused_features`` is the one-launch path the models use; the individual
functions exist for drop-in use (masking.py, audio_feat_preprocessing.py) and run the same
kernel with different outputs enabled.
"""

import numpy as np
import torch

from . import _lib

NFFT = 512
_tables = {}


def ms_to_samples(ms, sample_rate):
    return int(round(ms / 1e3 * sample_rate))


def _device_tables(device, frame_len, hop=None):
    key = (str(device), frame_len, hop)
    tab = _tables.get(key)
    if tab is None:
        k = np.arange(frame_len, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * k / frame_len)           # periodic Hann (tf.signal.hann_window)
        m = np.arange(NFFT, dtype=np.float64)
        tw = np.stack([np.cos(2.0 * np.pi * m / NFFT), -np.sin(2.0 * np.pi * m / NFFT)], axis=1)
        tab = {'window': torch.tensor(window, dtype=torch.float32, device=device),
               'twiddle': torch.tensor(tw, dtype=torch.float32, device=device).contiguous()}
        if hop is not None:
            # tf.contrib.signal.inverse_stft_window_fn(hop) for a Hann forward window
            overlaps = -(-frame_len // hop)
            denom = np.concatenate([window ** 2, np.zeros(overlaps * hop - frame_len)])
            denom = np.tile(denom.reshape(overlaps, hop).sum(0, keepdims=True), (overlaps, 1)).reshape(-1)
            tab['inv_window'] = torch.tensor(window / denom[:frame_len], dtype=torch.float32, device=device)
        _tables[key] = tab
    return tab


def linear_to_mel_weight_matrix(num_mel_bins=80, num_spec_bins=257, sample_rate=16000,
                                lower_edge_hertz=125.0, upper_edge_hertz=7600.0):
    """tf.signal.linear_to_mel_weight_matrix: HTK mel scale, triangles linear in mel, DC row zero.
    Host-side table construction in float64 (the projection itself runs in the kernel)."""
    def hz_to_mel(f):
        return 1127.0 * np.log1p(np.asarray(f, dtype=np.float64) / 700.0)
    nyquist = sample_rate / 2.0
    lin = np.linspace(0.0, nyquist, num_spec_bins)[1:]
    spec_mel = hz_to_mel(lin)[:, None]
    edges = np.linspace(hz_to_mel(lower_edge_hertz), hz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    lower, center, upper = edges[:-2][None], edges[1:-1][None], edges[2:][None]
    w = np.maximum(0.0, np.minimum((spec_mel - lower) / (center - lower), (upper - spec_mel) / (upper - center)))
    return np.concatenate([np.zeros((1, num_mel_bins)), w], axis=0)


_BANDS = {}
_MEL = {}


def _mel_bands(mel):
    """Band form of a mel matrix [n_bins, n_mel] for the fused log-mel kernel (avsi_frontend_args.mel_bands): every
    filter of linear_to_mel_weight_matrix is non-zero on one contiguous bin range.  None when it does not fit."""
    key = (mel.data_ptr(), tuple(mel.shape), mel._version, str(mel.device))
    if key not in _BANDS:
        m = mel.detach().cpu().numpy().astype(np.float32)
        n_bins, n_mel = m.shape
        packed = None
        if n_mel < 128 and n_bins <= 257:
            hdr = np.zeros(256, np.int32)
            ws = []
            n = 0
            for j in range(n_mel):
                nz = np.nonzero(m[:, j])[0]
                lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 0)
                hdr[j], hdr[128 + j] = lo, n
                ws.append(m[lo:hi, j])
                n += hi - lo
            hdr[128 + n_mel] = n
            if n <= 1024:
                buf = np.concatenate([hdr.view(np.float32)] + ws) if ws else hdr.view(np.float32)
                packed = torch.from_numpy(np.ascontiguousarray(buf)).to(mel.device)
        if len(_BANDS) > 16:
            _BANDS.clear()
        _BANDS[key] = packed
    return _BANDS[key]


def _f32(t, device):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=device, dtype=torch.float32).contiguous()


def fused_features(sources, frame_len, hop, T=None, F=257, mean=None, std=None, mask=None, video=None,
                   power=1.0, log=True, want_stft=False, stft_masked=False, want_spec=True, want_feat=False,
                   xh_out=None, ldx=0, mel=None, mel_eps=1e-6, hole_count=None, xh_video_only=False, xh_skip_pad=False,
                   mel_masked=False):
    """One launch of the fused front end.  Returns dict(stft, spec, feat, logmel) (missing = None).

    sources [B,N] f32 CUDA; mask [B,T,F]; video [B,T,V]; xh_out optional preallocated
    time-major fp16 [T*B, ldx] network-input buffer; mel = (matrix [257,n_mel] CUDA f32)."""
    if not sources.is_cuda:
        raise _lib.AvsiError('fused_features needs CUDA tensors (no CPU fallback)')
    lib = _lib.load()
    dev = sources.device
    sources = _f32(sources, dev)
    B, N = sources.shape
    if T is None:
        T = -(-N // hop)
    tab = _device_tables(dev, frame_len)
    mean, std, mask, video = _f32(mean, dev), _f32(std, dev), _f32(mask, dev), _f32(video, dev)
    V = 0 if video is None else video.shape[2]
    out = {'stft': None, 'spec': None, 'feat': None, 'logmel': None}
    if want_stft:
        out['stft'] = torch.empty(B, T, F, 2, dtype=torch.float32, device=dev)
    if want_spec:
        out['spec'] = torch.empty(B, T, F, dtype=torch.float32, device=dev)
    if want_feat:
        out['feat'] = torch.empty(B, T, F + V, dtype=torch.float32, device=dev)
    n_mel = 0
    if mel is not None:
        n_mel = mel.shape[1]
        out['logmel'] = torch.empty(B, T, n_mel, dtype=torch.float32, device=dev)
    a = _lib.FrontendArgs()
    a.wav, a.B, a.N = sources.data_ptr(), B, N
    a.frame_len, a.hop, a.nfft, a.T, a.F = frame_len, hop, NFFT, T, F
    a.window, a.twiddle = tab['window'].data_ptr(), tab['twiddle'].data_ptr()
    a.mean = mean.data_ptr() if mean is not None else None
    a.stdev = std.data_ptr() if std is not None else None
    a.mask = mask.data_ptr() if mask is not None else None
    a.video = video.data_ptr() if video is not None else None
    a.V = V
    a.power, a.log_flag, a.stft_masked = float(power), int(bool(log)), int(bool(stft_masked))
    a.stft_out = out['stft'].data_ptr() if want_stft else None
    a.spec_out = out['spec'].data_ptr() if want_spec else None
    a.feat_out = out['feat'].data_ptr() if want_feat else None
    a.xh_out = xh_out.data_ptr() if xh_out is not None else None
    a.ldx = int(ldx)
    a.logmel_out = out['logmel'].data_ptr() if mel is not None else None
    a.mel_w = mel.data_ptr() if mel is not None else None
    a.n_mel, a.mel_eps = n_mel, float(mel_eps)
    a.hole_count = hole_count.data_ptr() if hole_count is not None else None
    a.xh_video_only = int(bool(xh_video_only))
    a.mel_masked = int(bool(mel_masked))
    bands = _mel_bands(mel) if mel is not None else None
    a.mel_bands = bands.data_ptr() if bands is not None else None
    a.xh_skip_pad = int(bool(xh_skip_pad))          # the engine's x0 workspace is zero-initialised once
    # algorithmic bytes of this launch: every requested input / output once
    nbytes = 4 * B * N + (4 * B * T * F if mask is not None else 0) + 4 * B * T * V
    nbytes += (8 * B * T * F if want_stft else 0) + (4 * B * T * F if want_spec else 0)
    # net_inputs counts as the reference's fp32 [B,T,I] tensor (4*T*I per utterance) even though the kernel
    # writes the fp16 time-major copy: the actual traffic (2*ldx per row) is lower and is reported separately
    I_cols = (0 if xh_video_only else F) + V
    nbytes += (4 * B * T * (F + V) if want_feat else 0)
    nbytes += (4 * B * T * I_cols if (xh_out is not None and not want_feat) else 0)
    nbytes += 4 * B * T * n_mel
    with _lib.span('frontend', nbytes=nbytes):
        _lib.check(lib.avsi_frontend_fwd(a, _lib.stream_ptr()), 'avsi_frontend_fwd')
    if want_stft:
        out['stft'] = torch.view_as_complex(out['stft'])
    return out


def _sliced(x, out_shape):
    if all(int(s) == 0 for s in out_shape):
        return x
    return x[:out_shape[0], :out_shape[1], :out_shape[2]]


def get_stft(sources, sample_rate=16000, window_size=25, step_size=10, n_fft=512, out_shape=[0, 0, 0]):
    """Compute STFT (audio_processing.py:25-42) -> complex64 [B, T, 257] (sliced to out_shape)."""
    if n_fft != NFFT:
        raise _lib.AvsiError('only n_fft = 512 is supported (the reference never uses another)')
    frame_len, hop = ms_to_samples(window_size, sample_rate), ms_to_samples(step_size, sample_rate)
    res = fused_features(sources, frame_len, hop, log=False, want_stft=True, want_spec=False)
    return _sliced(res['stft'], out_shape)


def get_spectrogram(stfts, power=1, log=False, out_shape=[0, 0, 0]):
    """audio_processing.py:45-56 on an already computed complex STFT tensor: one element-wise kernel."""
    if not (torch.is_tensor(stfts) and stfts.is_cuda and stfts.is_complex()):
        raise _lib.AvsiError('get_spectrogram needs a complex CUDA tensor (no CPU fallback)')
    z = torch.view_as_real(stfts.to(torch.complex64).contiguous())
    spec = torch.empty(stfts.shape, dtype=torch.float32, device=stfts.device)
    _lib.check(_lib.load().avsi_spectrogram(_lib.ptr(z), spec.numel(), float(power), int(bool(log)), _lib.ptr(spec),
                                            _lib.stream_ptr()), 'avsi_spectrogram')
    return _sliced(spec, out_shape)


def get_log_mel_spectrogram(spectrograms, sample_rate=16000, num_spec_bins=257, num_mel_bins=80,
                            lower_edge_freq=125, upper_edge_freq=7600, eps=1e-6, out_shape=[0, 0, 0]):
    """audio_processing.py:59-72 on a spectrogram tensor (out_shape is ignored there too)."""
    if upper_edge_freq is None:
        upper_edge_freq = sample_rate / 2
    if not (torch.is_tensor(spectrograms) and spectrograms.is_cuda):
        raise _lib.AvsiError('get_log_mel_spectrogram needs a CUDA tensor (no CPU fallback)')
    key = (str(spectrograms.device), num_mel_bins, num_spec_bins, sample_rate, lower_edge_freq, upper_edge_freq)
    if key not in _MEL:
        m = linear_to_mel_weight_matrix(num_mel_bins, num_spec_bins, sample_rate, lower_edge_freq, upper_edge_freq)
        _MEL[key] = torch.tensor(m, dtype=torch.float32, device=spectrograms.device).contiguous()
    x = spectrograms.float().contiguous()
    out = torch.empty(x.shape[:-1] + (num_mel_bins,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().avsi_log_mel(_lib.ptr(x), _lib.ptr(_MEL[key]), x.numel() // x.shape[-1], x.shape[-1], num_mel_bins,
                                        float(eps), _lib.ptr(out), _lib.stream_ptr()), 'avsi_log_mel')
    return out


def preemphasis(sources, alpha=0.95):
    """audio_processing.py:19-22 -> CUDA f32 [B, N]."""
    x = _f32(sources, sources.device if torch.is_tensor(sources) and sources.is_cuda else torch.device('cuda'))
    if not x.is_cuda:
        raise _lib.AvsiError('preemphasis needs CUDA tensors (no CPU fallback)')
    y = torch.empty_like(x)
    _lib.check(_lib.load().avsi_preemphasis(_lib.ptr(x), x.shape[0], x.shape[1], float(alpha), _lib.ptr(y), _lib.stream_ptr()),
               'avsi_preemphasis')
    return y


def get_mfcc(log_mel_spectrograms, num_mfccs=13, out_shape=[0, 0, 0]):
    """audio_processing.py:74-81: DCT-II of the log-mel rows (tf.signal.mfccs_from_log_mel_spectrograms), first num_mfccs."""
    x = log_mel_spectrograms
    if not (torch.is_tensor(x) and x.is_cuda):
        raise _lib.AvsiError('get_mfcc needs a CUDA tensor (no CPU fallback)')
    x = x.float().contiguous()
    out = torch.empty(x.shape[:-1] + (num_mfccs,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().avsi_mfcc(_lib.ptr(x), x.numel() // x.shape[-1], x.shape[-1], num_mfccs, _lib.ptr(out),
                                     _lib.stream_ptr()), 'avsi_mfcc')
    return _sliced(out, out_shape)


def delta(features, N=2):
    """audio_processing.py:84-93 on [B, T, F]."""
    return add_delta_features(features, n_delta=1, N=N)[:, :, features.shape[2]:]


def add_delta_features(features, n_delta=2, N=2):
    """audio_processing.py:96-103: [B, T, F] -> [B, T, F * (n_delta + 1)] (features, delta, delta-delta, ...)."""
    x = features
    if not (torch.is_tensor(x) and x.is_cuda):
        raise _lib.AvsiError('add_delta_features needs a CUDA tensor (no CPU fallback)')
    x = x.float().contiguous()
    B, T, F = x.shape
    ld = F * (n_delta + 1)
    out = torch.empty(B, T, ld, dtype=torch.float32, device=x.device)
    out[:, :, :F] = x
    lib = _lib.load()
    for i in range(n_delta):
        src = out.data_ptr() + 4 * F * i
        _lib.check(lib.avsi_delta_features(src, ld, src + 4 * F, ld, B, T, F, N, _lib.stream_ptr()), 'avsi_delta_features')
    return out


def log_mel_features(sources, sample_rate=16000, window_size=25, step_size=10, num_mel_bins=80,
                     lower_edge_freq=125, upper_edge_freq=7600, eps=1e-6):
    """Fused `fbanks` path of audio_feat_preprocessing.py:49-50 / models_asr.py:31-37:
    power-2 spectrogram -> mel projection -> log, one kernel."""
    frame_len, hop = ms_to_samples(window_size, sample_rate), ms_to_samples(step_size, sample_rate)
    key = (str(sources.device), num_mel_bins, sample_rate, lower_edge_freq, upper_edge_freq)
    if key not in _MEL:
        m = linear_to_mel_weight_matrix(num_mel_bins, 257, sample_rate, lower_edge_freq, upper_edge_freq)
        _MEL[key] = torch.tensor(m, dtype=torch.float32, device=sources.device).contiguous()
    mel = _MEL[key]
    res = fused_features(sources, frame_len, hop, power=2.0, log=False, want_spec=False, mel=mel, mel_eps=eps)
    return res['logmel']


def reconstruct_from(mag, phase_src, mask=None, mean=None, std=None, num_samples=48000, sample_rate=16000,
                     window_size=24, step_size=12):
    """Fused inverse path (audio_processing.py:145-164 + models.py:185-187): mag (or normalised
    prediction when mean/std given) x unit phase of ``phase_src`` (x mask) -> iSTFT -> [B, num_samples]."""
    lib = _lib.load()
    dev = mag.device
    frame_len, hop = ms_to_samples(window_size, sample_rate), ms_to_samples(step_size, sample_rate)
    tab = _device_tables(dev, frame_len, hop)
    mag = _f32(mag, dev)
    B, T, F = mag.shape
    ph = torch.view_as_real(phase_src.to(torch.complex64).contiguous()).contiguous()
    mask, mean, std = _f32(mask, dev), _f32(mean, dev), _f32(std, dev)
    out_len = num_samples if num_samples > 0 else (T - 1) * hop + frame_len
    out = torch.empty(B, out_len, dtype=torch.float32, device=dev)
    a = _lib.IstftArgs()
    a.mag, a.phase_src = mag.data_ptr(), ph.data_ptr()
    a.mask = mask.data_ptr() if mask is not None else None
    a.mean = mean.data_ptr() if mean is not None else None
    a.stdev = std.data_ptr() if std is not None else None
    a.inv_window, a.twiddle = tab['inv_window'].data_ptr(), tab['twiddle'].data_ptr()
    a.B, a.T, a.F, a.frame_len, a.hop, a.nfft, a.num_samples = B, T, F, frame_len, hop, NFFT, int(num_samples)
    a.out = out.data_ptr()
    _lib.check(lib.avsi_istft_fwd(a, _lib.stream_ptr()), 'avsi_istft_fwd')
    return out


def reconstruct_sources(stfts, num_samples=0, sample_rate=16000, window_size=16, step_size=8):
    """Compute inverse STFT (audio_processing.py:145-157)."""
    return reconstruct_from(torch.abs(stfts), stfts, num_samples=num_samples, sample_rate=sample_rate,
                            window_size=window_size, step_size=step_size)


def get_sources(mag_spectrograms, rec_ang_spectrograms, num_samples=48000, sample_rate=16000, window_size=24,
                step_size=12):
    """Get waveform from magnitude and phase of STFT (audio_processing.py:160-164)."""
    unit = torch.polar(torch.ones_like(rec_ang_spectrograms), rec_ang_spectrograms)
    return reconstruct_from(mag_spectrograms, unit, num_samples=num_samples, sample_rate=sample_rate,
                            window_size=window_size, step_size=step_size)


def downsampling(samples, sample_rate, downsample_rate):
    """audio_processing.py:9-16: Fourier-domain resampling of a whole recording (scipy.signal.resample semantics) on the
    GPU -- `avsi_resample_fft`, chirp-z DFTs of any length in complex double.  `samples`: 1-D array (or [batch, n] of equal
    lengths); returns float64 like scipy."""
    import numpy as np
    x = np.asarray(samples)
    n_in = x.shape[-1]
    secs = n_in / float(sample_rate)
    num_samples = int(downsample_rate * secs)
    if sample_rate == downsample_rate:
        return samples
    return resample(x, num_samples)


def resample(x, num):
    """scipy.signal.resample(x, num) along the last axis for real x, on the GPU (float64 result)."""
    import numpy as np
    x = np.asarray(x)
    if np.iscomplexobj(x):
        raise ValueError('resample: real input only (the reference resamples wav samples)')
    lead = x.shape[:-1]
    n_in = x.shape[-1]
    if n_in < 1 or num < 1:
        raise ValueError('resample: empty input or output')
    lib = _lib.load()
    xd = torch.as_tensor(np.ascontiguousarray(x.reshape(-1, n_in), dtype=np.float64), device='cuda')
    batch = xd.shape[0]
    need = lib.avsi_resample_workspace_bytes(batch, n_in, num)
    if need < 0:
        raise ValueError('resample: unsupported sizes %d -> %d' % (n_in, num))
    work = torch.empty(need, dtype=torch.uint8, device='cuda')
    y = torch.empty(batch, num, dtype=torch.float64, device='cuda')
    _lib.check(lib.avsi_resample_fft(_lib.ptr(xd), batch, n_in, _lib.ptr(y), num, _lib.ptr(work), need, _lib.stream_ptr()),
               'avsi_resample_fft')
    return y.cpu().numpy().reshape(lead + (num,))
