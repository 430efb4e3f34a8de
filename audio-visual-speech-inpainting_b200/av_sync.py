"""Audio-visual synchronisation with the reference's names (av_sync.py:7-40), GPU-backed.

``sync_audio_visual_features`` keeps the reference's host-side acceptance logic (reject < min_frames,
start/end padding by repeating frame 0) and hands the interpolation to the
``avsi_video_features`` kernel through ``video_pipeline`` (upsample -> motion vector -> z-norm in
one launch).  ``inc_fps`` alone (upsampling only) is available for drop-in use.
"""
import numpy as np
import torch

from . import _lib


def pad_landmarks(video_features, tot_frames=None, min_frames=None, pad='start'):
    """Host half of av_sync.py:15-33: returns the padded [L,D] array or None if rejected."""
    video_features = np.asarray(video_features)
    if video_features.ndim != 2 or (min_frames is not None and video_features.shape[0] < min_frames):
        return None
    if tot_frames is not None and video_features.shape[0] < tot_frames:
        n_rep = tot_frames - video_features.shape[0]
        rep = np.tile(video_features[0], (n_rep, 1))
        video_features = np.vstack((rep, video_features)) if pad == 'start' else (
            np.vstack((video_features, rep)) if pad == 'end' else video_features)
    return video_features


def video_pipeline(landmarks, target_len, vmean, vstd, device='cuda'):
    """landmarks [B,L,D] (padded), vmean/vstd [B,D] or [D] -> z-normed motion vectors [B,T,D] f32 CUDA.
    tfrecord_utils.py:90-107 order: upsample -> first difference -> normalise."""
    lib = _lib.load()
    def dev32(x):
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x))
        return x.to(device=device, dtype=torch.float32)
    lm = dev32(landmarks).contiguous()
    B, L, D = lm.shape
    vm, vs = dev32(vmean), dev32(vstd)
    vm = vm.expand(B, D).contiguous() if vm.dim() == 1 else vm.contiguous()
    vs = vs.expand(B, D).contiguous() if vs.dim() == 1 else vs.contiguous()
    out = torch.empty(B, target_len, D, dtype=torch.float32, device=device)
    _lib.check(lib.avsi_video_features(_lib.ptr(lm), _lib.ptr(vm), _lib.ptr(vs), B, L, D, target_len, _lib.ptr(out),
                                       _lib.stream_ptr()), 'avsi_video_features')
    return out


def sync_audio_visual_features(mask, video_features, tot_frames=None, min_frames=None, pad='start'):
    """av_sync.py:15-40.  Returns the upsampled landmarks [len(mask), D] (numpy float64) or None."""
    padded = pad_landmarks(video_features, tot_frames, min_frames, pad)
    if padded is None:
        return None
    return inc_fps(padded, len(mask))


def inc_fps(frames, target_len):
    """av_sync.py:7-12: clamped linear interpolation in time (what interp2d(kind='linear') does on a
    regular grid).  Host numpy (offline data preparation; the training path uses video_pipeline)."""
    frames = np.asarray(frames, dtype=np.float64)
    L = frames.shape[0]
    y = np.minimum(np.linspace(0, L * (1 - 1 / target_len), target_len), L - 1)
    i0 = np.minimum(np.floor(y).astype(np.int64), L - 1)
    i1 = np.minimum(i0 + 1, L - 1)
    w = (y - i0)[:, None]
    return frames[i0] * (1.0 - w) + frames[i1] * w
