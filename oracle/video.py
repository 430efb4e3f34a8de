"""Oracle (test infrastructure): landmark upsampling and motion vectors.

Follows av_sync.py:7-40 (inc_fps / sync_audio_visual_features),
face_landmarks.py:30-39 (get_motion_vector, delta=1) and the order of
operations in tfrecord_utils.py:86-107 (pad -> upsample -> diff -> z-norm).
``scipy.interpolate.interp2d`` (removed from SciPy >= 1.14) is restated as the
clamped linear interpolation it performs on a regular grid.
"""
import numpy as np


def inc_fps(frames, target_len):
    """av_sync.py:7-12.  frames [L, D] -> [target_len, D] (float64)."""
    frames = np.asarray(frames, dtype=np.float64)
    L = frames.shape[0]
    y = np.linspace(0, L * (1 - 1 / target_len), target_len)
    y = np.minimum(y, L - 1)                      # interp2d clamps outside the grid
    i0 = np.floor(y).astype(np.int64)
    i0 = np.minimum(i0, L - 1)
    i1 = np.minimum(i0 + 1, L - 1)
    w = (y - i0)[:, None]
    return frames[i0] * (1.0 - w) + frames[i1] * w


def sync_audio_visual_features(mask_len, video_features, tot_frames=None, min_frames=None, pad='start'):
    """av_sync.py:15-40 (``mask`` is only used for its length there)."""
    video_features = np.asarray(video_features)
    if video_features.ndim != 2 or (min_frames is not None and video_features.shape[0] < min_frames):
        return None
    if tot_frames is not None and video_features.shape[0] < tot_frames:
        n_rep = tot_frames - video_features.shape[0]
        rep = np.tile(video_features[0], (n_rep, 1))
        if pad == 'start':
            video_features = np.vstack((rep, video_features))
        elif pad == 'end':
            video_features = np.vstack((video_features, rep))
    out = inc_fps(video_features, mask_len)
    return out if len(out) == mask_len else None


def get_motion_vector(landmarks, delta=1):
    """face_landmarks.py:30-39 (anchor_landmark < 0)."""
    landmarks = np.asarray(landmarks)
    feats = np.zeros_like(landmarks)
    feats[1:] = landmarks[1:] - landmarks[:-1]
    if delta == 2:
        feats = feats[1:] - feats[:-1]
    return feats


def video_features(landmarks, target_len, vmean, vstd, tot_frames=75, min_frames=70):
    """tfrecord_utils.py:86-107: landmarks [L,136] -> z-normed motion vectors [T,136]."""
    up = sync_audio_visual_features(target_len, landmarks, tot_frames=tot_frames, min_frames=min_frames)
    if up is None:
        return None
    mv = get_motion_vector(up, delta=1)
    return (mv - vmean) / vstd
