"""train(config_file) of the reference's training_asr.py:23 for the phone-recognition model (models_asr.py) on the B200
hot path: same experiment-folder layout, checkpoints and epoch loop as training.train (the loop is shared), the
monitored figures being the CTC loss and the PER; the best validation checkpoint is the one with the lowest CTC loss.
`config['model']` is "a-blstm" or "av-blstm" (training_asr.py:81-92; the video-only variant is not on the hot path),
the normalisation files are the log-mel statistics written by compute_mean_std_features(type='fbanks'), and the optional
config key `apply_mask` multiplies the power spectrogram by the mask before the mel projection (models_asr.py:33-36)."""
import sys

import numpy as np

from . import models_asr
from . import training as _training


def build_model(config, batch, mean, std, is_training=True, device='cuda', process_group=None):
    name = config['model']
    inp = {'a-blstm': 'a', 'av-blstm': 'av'}.get(name)
    if inp is None:
        print('Model selection must be "a-blstm", "v-blstm", "av-blstm". Closing...')
        sys.exit(1)
    seq, lab_len, wav, _, labels, video, mask = batch
    model = models_asr.StackedBLSTMModel(seq, lab_len, wav.astype(np.float32), mask, labels, mean, std, 0.0, config,
                                         video_features=video if inp == 'av' else None, input=inp,
                                         apply_mask=bool(config.get('apply_mask', False)), is_training=is_training,
                                         device=device, process_group=process_group)
    model.build_graph('asr/' + name)
    return model


def feed_batch(model, batch, dropout_rate=0.0):
    """Returns the number of labels of the batch: the running averages are weighted by it (training_asr.py:228-239)."""
    seq, lab_len, wav, _, labels, video, mask = batch
    model.feed(sequence_lengths=seq, labels_lengths=lab_len, target_sources=wav.astype(np.float32), masks=mask,
               labels=labels, video_features=video if model.input_type == 'av' else None, dropout_rate=dropout_rate)
    return int(np.sum(lab_len))


def _losses(model, want_per):
    loss, ctc = float(model.loss), float(model.ctc_loss)
    per = float(model.per.mean()) if want_per else 0.0
    return loss, ctc, ctc, per


def train(config_file, max_steps=None):
    """Train the phone-recognition model."""
    return _training._train(config_file, max_steps, build_model, feed_batch, _losses, best_name='asrnet')   # training_asr.py:308
