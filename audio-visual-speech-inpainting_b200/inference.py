"""infer(...): the reference's inference job (inference.py:20, working twin inference_siasr_ctc.py:22) on the B200
hot path: restore `netmodel/sinet`, run the model over the test TFRecords, reconstruct the waveform with the masked
(or oracle) phase through the fused iSTFT kernel and write `<audio_path>/<sample>/enhanced/<prefix>.wav` as int16
(first seq_len * hop samples, inference_siasr_ctc.py:241-243).  Without the oracle phase the reference refines the
phase of the inpainted frames with the `lws` C extension (inference.py:143-154); that package is not available here, so
the same bookkeeping runs around the exact consistency projection on the GPU (phase_reconstruction.refine_phase,
`phase_iterations` of them; 0 writes the model's masked-phase `enhanced_sources` tensor as it is)."""
import os
from glob import glob

import numpy as np

from . import checkpoint
from .config_utils import load_configfile
from .dataset_reader import DataManager
from .training import build_model, feed_batch


def infer(model_path, data_path_test, audio_path, out_file_prefix, norm=True, oracle_phase=False, batch_size=1,
          phase_iterations=100):
    from scipy.io import wavfile
    config = load_configfile(os.path.join(model_path, 'config.txt'))
    for key, val in (('audio_feat_dim', 257), ('video_feat_dim', 136), ('num_asr_labels', 33), ('ctc_loss', 1),
                     ('optimizer_type', 'adam'), ('l2', 0.0)):
        config.setdefault(key, val)
    config['num_asr_labels'] += 1                           # + blank, as check_trainconfiguration did at training time
    config['batch_size'] = batch_size
    if norm:
        mean = np.load(os.path.join(model_path, 'audio_features_mean.npy')).astype(np.float32)
        std = np.load(os.path.join(model_path, 'audio_features_std.npy')).astype(np.float32)
    else:
        mean = np.zeros(config.get('audio_feat_dim', 257), np.float32)
        std = np.ones(config.get('audio_feat_dim', 257), np.float32)
    dm = DataManager(num_audio_samples=config['audio_len'], audio_feat_size=config.get('audio_feat_dim', 257),
                     video_feat_size=config.get('video_feat_dim', 136), mode='fixed')
    files = sorted(glob(os.path.join(data_path_test, '*.tfrecord')))
    _, it = dm.get_iterator(dm.get_dataset(files, shuffle=False), batch_size=batch_size, n_epochs=1)
    model, hop = None, None
    loss_sum, n_done = 0.0, 0
    for batch in it:
        if model is None:
            model = build_model(config, batch, mean, std, is_training=False)
            try:
                checkpoint.restore(model, os.path.join(model_path, 'sinet'), train_vars_only=True)
                print('Model variables restored.')
            except ValueError:
                print('{:s} is not a valid checkpoint. Closing...'.format(os.path.join(model_path, 'sinet')))
                raise SystemExit(2)
            hop = model.hop
        feed_batch(model, batch)
        enhanced = model.enhanced_sources_oracle_phase if oracle_phase else model.enhanced_sources
        if not oracle_phase and phase_iterations > 0:
            from .phase_reconstruction import refine_phase
            enhanced = refine_phase(enhanced, model._fed['masks'], n_iter=phase_iterations)
        enhanced = enhanced.cpu().numpy()
        loss_sum += float(model.loss_hole) * len(batch[0])
        for b, name in enumerate(batch[3]):
            name = name.decode() if isinstance(name, bytes) else str(name)
            out_dir = os.path.join(audio_path, name, 'enhanced')
            os.makedirs(out_dir, exist_ok=True)
            n = int(batch[0][b]) * hop
            wavfile.write(os.path.join(out_dir, out_file_prefix + '.wav'), 16000, enhanced[b, :n].astype(np.int16))
            n_done += 1
        if n_done % 100 < len(batch[0]):
            print('{:d} samples processed. Inpainting loss: {:.5f}'.format(n_done, loss_sum / max(n_done, 1)))
    print('done. {:d} samples, mean inpainting (hole) loss {:.5f}'.format(n_done, loss_sum / max(n_done, 1)))
    return loss_sum / max(n_done, 1)
