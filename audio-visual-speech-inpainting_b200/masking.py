"""mask_app(...): the reference's mask-application job (masking.py:18-103): STFT -> x mask -> magnitude, phase of
the masked (or target) STFT -> iSTFT -> `<audio_path>/<sample>/masked.wav` (int16, first seq_len * hop samples).
One fused front-end launch and one fused iSTFT launch per batch; the four docs/files fixtures pin this chain to
+-1 LSB (tests/test_gpu_kernels.py::test_mask_app_chain_matches_docs_fixtures)."""
import os
from glob import glob

import numpy as np
import torch

from . import audio_processing as ap
from .dataset_reader import DataManager


def masked_waveforms(wav, mask, oracle_phase=False, window_size=24, step_size=12):
    """wav [B,N], mask [B,T,F] (CUDA or numpy) -> masked waveforms [B,N] f32 CUDA (masking.py:41-45)."""
    dev = torch.device('cuda')
    wav = torch.as_tensor(np.asarray(wav) if not torch.is_tensor(wav) else wav).to(dev, torch.float32)
    mask = torch.as_tensor(np.asarray(mask) if not torch.is_tensor(mask) else mask).to(dev, torch.float32)
    frame_len, hop = ap.ms_to_samples(window_size, 16000), ap.ms_to_samples(step_size, 16000)
    T, F = mask.shape[1], mask.shape[2]
    res = ap.fused_features(wav, frame_len, hop, T=T, F=F, mask=mask, log=False, want_stft=True, stft_masked=True,
                            want_spec=False)
    masked_stft = res['stft']
    phase_src = masked_stft
    if oracle_phase:
        phase_src = ap.fused_features(wav, frame_len, hop, T=T, F=F, log=False, want_stft=True, want_spec=False)['stft']
    return ap.reconstruct_from(ap.get_spectrogram(masked_stft), phase_src, num_samples=wav.shape[1], window_size=window_size,
                               step_size=step_size)


def mask_app(data_path, audio_path, tfrecord_mode='fixed', oracle_phase=True, audio_feat_dim=257, video_feat_dim=136,
             num_audio_samples=48000, batch_size=1):
    from scipy.io import wavfile
    dm = DataManager(num_audio_samples=num_audio_samples, audio_feat_size=audio_feat_dim, video_feat_size=video_feat_dim,
                     mode=tfrecord_mode)
    files = sorted(glob(os.path.join(data_path, '*.tfrecord')))
    _, it = dm.get_iterator(dm.get_dataset(files, shuffle=False), batch_size=batch_size, n_epochs=1)
    hop = ap.ms_to_samples(12, 16000)
    n_done = 0
    for seq, _, wav, paths, _, _, mask in it:
        out = masked_waveforms(wav.astype(np.float32), mask, oracle_phase).cpu().numpy()
        for b, name in enumerate(paths):
            name = name.decode() if isinstance(name, bytes) else str(name)
            os.makedirs(os.path.join(audio_path, name), exist_ok=True)
            wavfile.write(os.path.join(audio_path, name, 'masked.wav'), 16000, out[b, :int(seq[b]) * hop].astype(np.int16))
            n_done += 1
    print('done. {:d} masked files written.'.format(n_done))
    return n_done
