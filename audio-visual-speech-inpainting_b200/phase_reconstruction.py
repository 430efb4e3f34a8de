"""Phase refinement of the inpainted frames: the post-processing of inference.py:143-154.

The reference hands the network's waveform (masked phase: the recording's phase in the reliable frames, zero in the
holes) to the `lws` package -- `lws.lws(384, 192, fftsize=512, mode='speech').run_lws(...)`, Le Roux et al.'s "local
weighted sums" -- and keeps the refined phase in the holes only.  `lws` is a C extension that is neither listed in the
reference's requirements, nor vendored, nor present in this image, so its output cannot be reproduced bit for bit and
there is nothing to hold a port against.  What LWS approximates (with truncated kernels and magnitude thresholds, for
speed on a CPU) is the projection onto consistent spectrograms, i.e. the Griffin-Lim iteration
    X <- mag . exp(i angle(STFT(iSTFT(X))));
on a B200 the exact projection is three kernel launches, so that is what runs here, with the reference's
bookkeeping around it: magnitudes of the enhanced waveform's STFT, phases fixed to the recording's wherever the mask is 1,
refined wherever it is 0, inverse STFT with the reference's window / hop (24 ms / 12 ms, audio_processing.py:160).
NOT comparable sample by sample with the lws package; pinned to its own float64 restatement (oracle/phase.py).
"""
import torch

from . import _lib
from . import audio_processing as ap


def _select(a, b, keep):
    out = torch.empty_like(a)
    ar = torch.view_as_real(a)
    br = torch.view_as_real(b) if b is not None else None
    _lib.check(_lib.load().avsi_select_c64(_lib.ptr(ar), _lib.ptr(br), _lib.ptr(keep), a.numel(),
                                           _lib.ptr(torch.view_as_real(out)), _lib.stream_ptr()), 'avsi_select_c64')
    return out


def refine_phase(enhanced, masks, n_iter=100, sample_rate=16000, window_size=24, step_size=12):
    """enhanced [B, N] f32 CUDA (model.enhanced_sources), masks [B, T, F] (1 = reliable) -> refined waveform [B, N]."""
    if not (torch.is_tensor(enhanced) and enhanced.is_cuda):
        raise _lib.AvsiError('refine_phase needs CUDA tensors (no CPU fallback)')
    enhanced = enhanced.to(torch.float32).contiguous()
    B, N = enhanced.shape
    stft = ap.get_stft(enhanced, sample_rate=sample_rate, window_size=window_size, step_size=step_size).contiguous()
    keep = torch.zeros(stft.shape, dtype=torch.float32, device=enhanced.device)      # mask_adj of inference.py:145-146
    m = torch.as_tensor(masks, dtype=torch.float32, device=enhanced.device)
    t, f = min(m.shape[1], keep.shape[1]), min(m.shape[2], keep.shape[2])
    keep[:, :t, :f] = m[:, :t, :f]
    mag = ap.get_spectrogram(stft)
    phase_src = _select(stft, None, keep)               # zero phase in the holes (ang_spec = angle * mask_adj)
    for _ in range(int(n_iter)):
        wav = ap.reconstruct_from(mag, phase_src, num_samples=N, sample_rate=sample_rate, window_size=window_size,
                                  step_size=step_size)
        again = ap.get_stft(wav, sample_rate=sample_rate, window_size=window_size, step_size=step_size).contiguous()
        phase_src = _select(stft, again, keep)          # rec_ang_adj = ang_spec + rec_ang * (1 - mask_adj)
    return ap.reconstruct_from(mag, phase_src, num_samples=N, sample_rate=sample_rate, window_size=window_size,
                               step_size=step_size)
