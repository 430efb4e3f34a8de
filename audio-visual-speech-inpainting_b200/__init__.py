"""B200-native hot path of dr-pato/audio-visual-speech-inpainting.

Host-side mirror of the reference's Python interfaces (audio_processing, models, config_utils,
av_sync, face_landmarks, dataset_generator, masking, training, inference) over hand-written
sm_100a CUDA kernels reached through the C ABI in include/avsi_b200.h.  Import as ``avsi_b200``.
"""
__version__ = '0.1.0'
