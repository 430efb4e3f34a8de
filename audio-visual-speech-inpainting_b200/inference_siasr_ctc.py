"""infer(...) of the reference's inference_siasr_ctc.py:22 -- what the CLI's `inference_siasr` sub-command runs -- on
the B200 hot path: per batch of test TFRecords

  1. the speech-inpainting model (restored from `<model_path>/sinet`) produces `enhanced_sources` (masked phase) or
     `enhanced_sources_oracle_phase` and the hole loss (inference_siasr_ctc.py:190-203);
  2. the phone-recognition model (restored from `<model_path_asr>/asrnet`) runs ON THE ENHANCED AUDIO with its own
     normalisation statistics and yields `decoding` (beam search), the CTC loss and the PER (:206-218) -- the waveform
     goes from the iSTFT kernel to the ASR front end on the device, without the reference's host round trip;
  3. `<audio_path>/<sample>/enhanced/<prefix>.wav` (int16, first seq_len * 192 samples, :240-243) and
     `<audio_path>/<sample>/transcriptions/<prefix>.lbl` (comma-separated phonemes, :251-259) are written.

The LWS phase refinement of :222-235 (applied to the WRITTEN waveform only; the recogniser hears the un-refined one, as in
the reference) needs the `lws` C extension, which is neither in the image nor vendored by the reference: its place is taken by
phase_reconstruction.refine_phase (exact consistency projection, `phase_iterations` of them; 0 = write the model's
masked-phase `enhanced_sources` tensor as it is)."""
import os
from glob import glob

import numpy as np

from . import checkpoint
from . import training as _training
from . import training_asr as _training_asr
from .config_utils import check_trainconfiguration, load_configfile
from .dataset_reader import DataManager
from .transcription2phonemes import get_phonemes_from_labels, load_dictionary


def _restore(model, folder, names):
    for n in names:
        path = os.path.join(folder, n)
        if os.path.exists(path + '.npz') or os.path.exists(path + '.index'):
            checkpoint.restore(model, path, train_vars_only=True)
            return path
    print('{:s} is not a valid checkpoint. Closing...'.format(os.path.join(folder, names[0])))
    raise SystemExit(2)


def infer(model_path, model_path_asr, data_path_test, audio_path, out_file_prefix, dictionary_file, norm=True,
          oracle_phase=False, batch_size=1, phase_iterations=100):
    from scipy.io import wavfile
    config = check_trainconfiguration(load_configfile(os.path.join(model_path, 'config.txt')))
    config_asr = check_trainconfiguration(load_configfile(os.path.join(model_path_asr, 'config.txt')))
    config['batch_size'] = config_asr['batch_size'] = batch_size
    F = config['audio_feat_dim']
    if norm:
        mean = np.load(os.path.join(model_path, 'audio_features_mean.npy')).astype(np.float32)
        std = np.load(os.path.join(model_path, 'audio_features_std.npy')).astype(np.float32)
    else:
        mean, std = np.zeros(F, np.float32), np.ones(F, np.float32)
    mean_asr = np.load(os.path.join(model_path_asr, 'audio_features_mean.npy')).astype(np.float32)
    std_asr = np.load(os.path.join(model_path_asr, 'audio_features_std.npy')).astype(np.float32)
    ph_dict = load_dictionary(dictionary_file)
    dm = DataManager(num_audio_samples=config['audio_len'], audio_feat_size=F, video_feat_size=config['video_feat_dim'],
                     buffer_size=4000, mode='fixed', embedding_size=512 if str(config['model']).endswith('-emb') else 0)
    files = sorted(glob(os.path.join(data_path_test, '*.tfrecord')))
    _, it = dm.get_iterator(dm.get_dataset(files, shuffle=False), batch_size=batch_size, n_epochs=1)
    model = model_asr = None
    loss_hole_list, loss_asr_list, per_list, total = [], [], [], 0
    print('Starting inference on dataset: {:s}'.format(data_path_test))
    for batch in it:
        seq, lab_len, wav, emb, paths, labels, video, mask = _training._unpack(batch)
        if model is None:
            print('Building speech inpainting inference model..')
            model = _training.build_model(config, batch, mean, std, is_training=False)
            print('done.')
            print('Building ASR inference model:')
            asr_batch = (seq, lab_len, wav, paths, labels, video, mask)
            model_asr = _training_asr.build_model(config_asr, asr_batch, mean_asr, std_asr, is_training=False)
            print('done.')
            print('Restore weigths:')
            _restore(model, model_path, ['sinet'])
            _restore(model_asr, model_path_asr, ['asrnet', 'sinet'])
            print('done.\n')
        # speech inpainting inference
        _training.feed_batch(model, batch)
        enhanced = model.enhanced_sources_oracle_phase if oracle_phase else model.enhanced_sources      # [B, audio_len] CUDA
        loss_hole = float(model.loss_hole)
        # ASR inference on the enhanced audio (a device tensor: no host round trip)
        model_asr.feed(sequence_lengths=seq, labels_lengths=lab_len, target_sources=enhanced, masks=mask, labels=labels,
                       video_features=video if model_asr.input_type == 'av' else None, dropout_rate=0.0)
        decoded, loss_asr, per = model_asr.decoding, float(model_asr.loss), model_asr.per
        if not oracle_phase and phase_iterations > 0:
            from .phase_reconstruction import refine_phase
            enhanced = refine_phase(enhanced, mask, n_iter=phase_iterations)
        enhanced = enhanced.cpu().numpy()
        for b, name in enumerate(paths):
            sample_dir = name.decode() if isinstance(name, bytes) else str(name)
            os.makedirs(os.path.join(audio_path, sample_dir, 'enhanced'), exist_ok=True)
            n = int(seq[b]) * 192
            wavfile.write(os.path.join(audio_path, sample_dir, 'enhanced', out_file_prefix + '.wav'), 16000,
                          enhanced[b, :n].astype(np.int16))
            dec = decoded[b]
            pad = np.where(dec == -1)[0]
            dec = dec[:len(dec) if len(pad) == 0 else pad.min()]
            os.makedirs(os.path.join(audio_path, sample_dir, 'transcriptions'), exist_ok=True)
            with open(os.path.join(audio_path, sample_dir, 'transcriptions', out_file_prefix + '.lbl'), 'w') as f:
                f.write(','.join(get_phonemes_from_labels(dec, ph_dict)))
        loss_hole_list.append(loss_hole)
        loss_asr_list.append(loss_asr)
        per_list += list(per)
        total += len(seq)
        print('Processed {:d} utterances. Total samples processed so far {:d}.'.format(len(seq), total))
    print('done.')
    res = (float(np.mean(loss_hole_list)), float(np.mean(loss_asr_list)), float(np.mean(per_list)))
    print('Loss hole: {:.5}'.format(res[0]))
    print('Loss ASR: {:.5}'.format(res[1]))
    print('PER: {:.5}'.format(res[2]))
    return res
