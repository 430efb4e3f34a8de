"""Landmark motion vectors (face_landmarks.py:30-39).  Landmark extraction (dlib / OpenCV,
face_landmarks.py:42-238) is vision pre-processing and out of scope."""
import numpy as np


def get_motion_vector(landmarks, delta=1, anchor_landmark=-1):
    """First-order frame difference, row 0 = 0 (face_landmarks.py:30-39).  Host numpy for drop-in
    use; the training path computes it inside the ``avsi_video_features`` kernel."""
    if anchor_landmark >= 0:
        raise NotImplementedError('anchor_landmark >= 0 is unused by the reference pipeline')
    landmarks = np.asarray(landmarks)
    features = landmarks
    if delta > 0:
        features = np.zeros_like(landmarks)
        features[1:] = landmarks[1:] - landmarks[:-1]
        if delta == 2:
            features = features[1:] - features[:-1]
    return features
