"""train(config_file): the reference's training job (training.py:23 / training_ctc.py:23, the loop the CLI's
`training` sub-command runs) on the B200 hot path.

Same config file, same experiment folder layout (`netmodel/{ckpt,sinet,config.txt,audio_features_*.npy}`,
`training_log.txt`), same console / log-file lines, same schedule (epochs over one pass of the shuffled training
TFRecords, validation pass, best-validation checkpoint `sinet`, early stopping, `ckpt` every 1000 steps, NaN / Inf
exit code 1).  The session run per step becomes model.feed(...) + model.train_op().  Differences, all host side:
checkpoints are `.npz` files keyed by the TF variable names (checkpoint.py); PER is computed from the best-path
decoding at the logging steps and on the validation set instead of a beam search on every step; TensorBoard
scalars are written when torch.utils.tensorboard is importable.  Under torchrun every rank trains on
files[rank::world] (of a per-epoch shuffle every rank draws with the same seed) in lock step -- the ranks agree every
step that all of them still hold a full batch, the ragged tail of an epoch is dropped (parallel.lockstep) -- with the
gradient all-reduce of parallel.py; rank 0 logs and saves."""
import os
import random
import shutil
import sys
from glob import glob
from time import time

import numpy as np

from . import checkpoint, parallel
from .config_utils import check_trainconfiguration, load_configfile
from .dataset_reader import DataManager
from .models import MODEL_REGISTRY


class _RunningAverage(object):
    """Averages weighted by the number of masked frames of the batch (training_ctc.py:285-297)."""

    def __init__(self):
        self.n, self.vals = 0, None

    def add(self, frames, vals):
        vals = np.asarray(vals, np.float64)
        if self.vals is None:
            self.n, self.vals = frames, vals
        else:
            prev = self.n
            self.n += frames
            self.vals = (self.vals * prev + vals * frames) / max(self.n, 1)
        return self.vals


def build_model(config, batch, mean, std, is_training=True, device='cuda', process_group=None):
    """Instantiate the model class `config['model']` names on the tensors of a first batch."""
    name = config['model']
    if name not in MODEL_REGISTRY:
        print('Model selection must be "a-blstm", "v-blstm", "av-blstm" or a "-ctc" variant of them. Closing...')
        sys.exit(1)
    cls, inp = MODEL_REGISTRY[name]
    seq, lab_len, wav, emb, _, labels, video, mask = _unpack(batch)
    kw = dict(video_features=video, input=inp, is_training=is_training, device=device, process_group=process_group)
    if name.endswith('-emb'):                                   # training_emb.py:98-110
        if emb is None:
            print('Model "{:s}" needs TFRecords with an `embedding` feature (tfrecord_emb_utils.py). Closing...'.format(name))
            sys.exit(1)
        model = cls(seq, wav.astype(np.float32), mask, mean, std, 0.0, config, embeddings=emb, **kw)
    elif name == 'av-blstm-twosteps':                           # training.py:99-101
        model = cls(seq, wav.astype(np.float32), mask, mean, std, 0.0, config, video, is_training=is_training,
                    device=device, process_group=process_group)
    elif cls.MTL:
        model = cls(seq, lab_len, wav.astype(np.float32), mask, labels, mean, std, 0.0, config, **kw)
    else:
        model = cls(seq, wav.astype(np.float32), mask, mean, std, 0.0, config, **kw)
    model.build_graph(name)
    return model


def _unpack(batch):
    """(seq_len, lab_len, wav, embedding | None, paths, labels, video, mask) from a 7-tuple (dataset_reader.py:77-79) or the
    8-tuple with embeddings (dataset_reader_emb.py:79-81)."""
    if len(batch) == 8:
        return batch
    seq, lab_len, wav, paths, labels, video, mask = batch
    return seq, lab_len, wav, None, paths, labels, video, mask


def feed_batch(model, batch, dropout_rate=0.0):
    """dropout_rate: config['dropout_rate'] on training steps, 0.0 on validation / inference (training.py:241, 303, 314)."""
    seq, lab_len, wav, emb, _, labels, video, mask = _unpack(batch)
    kw = dict(sequence_lengths=seq, target_sources=wav.astype(np.float32), masks=mask, video_features=video,
              dropout_rate=dropout_rate)
    if model.MTL:
        kw.update(labels_lengths=lab_len, labels=labels)
    if getattr(model, 'embedding_dim', 0):
        kw.update(embeddings=emb)
    model.feed(**kw)
    return np.count_nonzero(mask[:, :, 0] == 0)                 # masked frames of the batch


def _save(model, path, config):
    """saver.save(sess, path) (training.py:267,335).  `checkpoint_format` in the config file picks the container: 'npz'
    (default), 'tf' = TensorFlow's own tensor bundle (`<path>.index` + `.data-00000-of-00001`, loadable by the
    reference's saver.restore), or 'both'."""
    fmt = str(config.get('checkpoint_format', 'npz')).lower()
    if fmt not in ('npz', 'tf', 'both'):
        print('checkpoint_format must be "npz", "tf" or "both". Closing...')
        sys.exit(1)
    out = path
    if fmt in ('npz', 'both'):
        out = checkpoint.save(model, path)
    if fmt in ('tf', 'both'):
        out = checkpoint.save(model, path, fmt='tf')
    return out


def _losses(model, want_per):
    loss, hole = float(model.loss), float(model.loss_hole)
    ctc = float(model.ctc_loss) if model.MTL else 0.0
    per = float(model.per.mean()) if (model.MTL and want_per) else 0.0
    return loss, hole, ctc, per


def train(config_file, max_steps=None):
    """Train the speech inpainting model."""
    return _train(config_file, max_steps, build_model, feed_batch, _losses)


def _train(config_file, max_steps, build_model, feed_batch, _losses, best_name='sinet'):
    """The training job of training.py / training_ctc.py; `training_asr.train` runs it with the phone-recognition model's
    hooks (model construction, feed, the monitored loss tuple whose entry 1 selects the best validation checkpoint)."""
    import torch
    import torch.distributed as dist
    config = check_trainconfiguration(load_configfile(config_file))
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    pg = dist.group.WORLD if world > 1 else None
    if world > 1:
        parallel.bind_to_gpu_numa(torch.cuda.current_device())
    data_path_train = os.path.join(config['root_folder'], 'training-set')
    data_path_val = os.path.join(config['root_folder'], 'validation-set')
    exp_path = config['exp_folder']
    exp_name = os.path.basename(exp_path)
    checkpoints_dir = os.path.join(exp_path, 'netmodel')
    log_path = os.path.join(exp_path, 'training_log.txt')

    def manager():
        return DataManager(num_audio_samples=config['audio_len'], audio_feat_size=config['audio_feat_dim'],
                           video_feat_size=config['video_feat_dim'], buffer_size=4000, mode='fixed', rank=rank, world=world,
                           embedding_size=512 if str(config['model']).endswith('-emb') else 0)    # training_emb.py:46
    train_files = sorted(glob(os.path.join(data_path_train, '*.tfrecord')))
    val_files = sorted(glob(os.path.join(data_path_val, '*.tfrecord')))
    audio_feat_mean = np.load(config['audio_feat_mean']).astype(np.float32)
    audio_feat_std = np.load(config['audio_feat_std']).astype(np.float32)
    per_rank_batch = max(1, config['batch_size'] // world)

    if rank == 0:
        os.makedirs(checkpoints_dir, exist_ok=True)
        os.makedirs(os.path.join(exp_path, 'tfboard'), exist_ok=True)
        dest_config_file = os.path.join(checkpoints_dir, 'config.txt')
        if os.path.abspath(dest_config_file) != os.path.abspath(config_file):
            shutil.copy(config_file, dest_config_file)
        shutil.copy(config['audio_feat_mean'], os.path.join(checkpoints_dir, 'audio_features_mean.npy'))
        shutil.copy(config['audio_feat_std'], os.path.join(checkpoints_dir, 'audio_features_std.npy'))
    log = open(log_path, 'a') if rank == 0 else None
    tb = None
    if rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            tb = SummaryWriter(os.path.join(exp_path, 'tfboard'))
        except Exception:
            tb = None

    seq_lengths_train = np.load(os.path.join(data_path_train, 'seq_lengths.npy'))
    seq_lengths_val = np.load(os.path.join(data_path_val, 'seq_lengths.npy'))
    train_size, val_size = len(seq_lengths_train), len(seq_lengths_val)
    n_steps_epoch = max(1, int(train_size / config['batch_size']))
    n_steps = n_steps_epoch * config['max_n_epochs']

    # model on the shape of a first batch
    _, probe = manager().get_iterator(manager().get_dataset(train_files, shuffle=False), batch_size=per_rank_batch, n_epochs=1)
    model = build_model(config, next(probe), audio_feat_mean, audio_feat_std, True, process_group=pg)
    print('Model building done.')
    if config.get('model_ckp_vnet') and hasattr(model, 'video_model'):      # training.py:115-116,153-159 (two-step model)
        try:
            checkpoint.restore(model.video_model, config['model_ckp_vnet'], train_vars_only=True)
            print('Visual model variables restored.')
        except ValueError:
            print('{:s} is not a valid checkpoint. Closing...'.format(config['model_ckp_vnet']))
            sys.exit(2)
    if config.get('model_ckp'):
        try:
            checkpoint.restore(model, config['model_ckp'])
            print('Model variables restored.')
        except ValueError:
            print('{:s} is not a valid checkpoint. Closing...'.format(config['model_ckp']))
            sys.exit(2)
    header = ['+-- EXPERIMENT NAME - {:s} --+'.format(exp_name), '## Model type: {:s}'.format(config['model']),
              '## Network dimensions: {:s}'.format(str(config['net_dim'])), '## Optimizer: {:s}'.format(config['optimizer_type']),
              '## Starter learning rate: {:.6f}'.format(config['starter_learning_rate']),
              '## Learning rate update steps: {:d}'.format(config['lr_updating_steps']),
              '## Learning rate decay: {:.6f}'.format(config['lr_decay']),
              '## CTC-loss coefficient: {:.6f}'.format(float(config.get('ctc_loss', 0))),
              '## L2 regularization coefficient: {:.6f}'.format(config['l2']),
              '## Dropout rate (no dropout if 0): {:.6f}'.format(config['dropout_rate']),
              '## Training dataset: {:s}'.format(data_path_train), '## Training size: {:d}'.format(train_size),
              '## Validation dataset: {:s}'.format(data_path_val), '## Validation size: {:d}'.format(val_size),
              '## Batch size: {:d}'.format(config['batch_size']),
              '## Approximated number of steps per epoch: {:d}'.format(n_steps_epoch),
              '## Number of training epochs: {:d}'.format(config['max_n_epochs']),
              '## Approximated total number of steps: {:d}'.format(n_steps)]
    if rank == 0:
        if not config.get('model_ckp'):
            log.write('\n'.join(header) + '\n')
            log.write('\nEpoch\tLR\tTraining loss\tTraining PER \tValidation loss\tValidation PER[TIME]\n')
        print('\n' + '\n'.join(header[:-1]) + '\n')

    tot_step = model.global_step
    epoch_counter = int(tot_step / n_steps_epoch)
    best_val_checkpoint, best_val_loss, cneg_epochs = (0, 0), -1.0, 0
    train_start_time = time()
    stop = False
    for n_epoch in range(config['max_n_epochs']):
        epoch_counter += 1
        epoch_start_time = time()
        files = list(train_files)
        if world > 1:
            # the same permutation on every rank (training.py:212 shuffles with the process-global generator): the
            # shards files[rank::world] of it are disjoint and change from epoch to epoch
            random.Random(int(config.get('seed', 0)) * 1000003 + epoch_counter).shuffle(files)
        else:
            random.shuffle(files)
        _, train_it = manager().get_iterator(manager().get_dataset(files, shuffle=True, keep_order=True),
                                             batch_size=per_rank_batch, n_epochs=1)
        train_it = parallel.lockstep(train_it, per_rank_batch, pg, model.device)
        if rank == 0:
            print('-> Epoch {:d}'.format(epoch_counter))
        avg, n_step, lr = _RunningAverage(), 0, model.learning_rate
        for batch in train_it:
            n_step += 1
            tot_step += 1
            frames = feed_batch(model, batch, dropout_rate=config['dropout_rate'])
            lr = model.learning_rate
            model.train_op()
            log_now = (n_step % 200 == 0 or n_step == 1)
            loss, loss_ipt, loss_ctc, per = _losses(model, log_now)
            if np.isnan(loss):
                print('GOT INSTABILITY: loss is NaN. Leaving...')
                sys.exit(1)
            if np.isinf(loss):
                print('GOT INSTABILITY: loss is inf. Leaving...')
                sys.exit(1)
            tr = avg.add(frames, [loss, loss_ipt, loss_ctc, per if log_now else (avg.vals[3] if avg.vals is not None else 0.0)])
            if rank == 0 and log_now:
                print('Step[{:7d}] Loss[{:3.5f}|{:3.5f}|{:3.5f}] PER[{:.5f}] LR[{:.6f}] Epoch training time[{:.2f}]'
                      .format(tot_step, tr[0], tr[1], tr[2], tr[3], lr, time() - epoch_start_time))
            if rank == 0 and n_step % 1000 == 0:
                print('Model checkpoint saved in file %s' % _save(model, os.path.join(checkpoints_dir, 'ckpt'), config))
            if max_steps is not None and tot_step >= max_steps:
                stop = True
                break
        tr = avg.vals if avg.vals is not None else np.zeros(4)
        epoch_duration = time() - epoch_start_time
        if rank == 0:
            print('Completed epoch {:d} at step {:d} --> Training loss: {:3.5f} - {:3.5f} - {:3.5f}; PER: {:3.5f}'
                  .format(epoch_counter, tot_step, tr[0], tr[1], tr[2], tr[3]))
            print('Epoch training time (seconds) = {:.6f}'.format(epoch_duration))
            print('Start validation set evaluation...')
        # ---- validation: same model, no update ----------------------------------------------------------------
        _, val_it = manager().get_iterator(manager().get_dataset(val_files, shuffle=False), batch_size=per_rank_batch, n_epochs=1)
        val_it = parallel.lockstep(val_it, per_rank_batch, pg, model.device)   # the MTL loss all-reduces its hole count
        vavg, n_vstep, summaries = _RunningAverage(), 0, None
        for batch in val_it:
            n_vstep += 1
            frames = feed_batch(model, batch)
            va = vavg.add(frames, _losses(model, True))
            if n_vstep == 1 and rank == 0 and tb is not None and hasattr(model, 'summaries'):
                # model.summaries of the first validation batch (training.py:295-304): images + audio, once per epoch
                try:
                    summaries = {k: (kind, t.detach().float().cpu()) for k, (kind, t) in model.summaries.items()}
                except Exception as e:                           # noqa: BLE001 -- monitoring must not stop the job
                    print('summaries skipped: %s' % e)
            if rank == 0 and (n_vstep % 200 == 0 or n_vstep == 1):
                print('Step[{:7d}] Loss[{:3.5f}]'.format(n_vstep, va[1]))
        va = vavg.vals if vavg.vals is not None else np.zeros(4)
        if world > 1:                                            # average the validation figures over the ranks
            t = torch.tensor(np.concatenate([va * vavg.n, [vavg.n]]), dtype=torch.float64, device=model.device)
            dist.all_reduce(t)
            va = (t[:4] / t[4].clamp(min=1)).cpu().numpy()
        if rank == 0:
            print('done.')
            print('Validation loss: {:3.5f}; PER: {:3.5f}. Best loss so far {:2.5f} [Epoch {:d} (step {:d})]'
                  .format(va[1], va[3], best_val_loss, best_val_checkpoint[0], best_val_checkpoint[1]))
        if best_val_checkpoint == (0, 0) or va[1] < best_val_loss:
            if rank == 0:
                print('Model saved in file %s' % _save(model, os.path.join(checkpoints_dir, best_name), config))
            best_val_checkpoint, best_val_loss, cneg_epochs = (epoch_counter, tot_step), va[1], 0
        else:
            cneg_epochs += 1
        if rank == 0:
            if tb is not None:
                for tag, val in (('Training loss full', tr[0]), ('Training loss inpainting', tr[1]), ('Training loss CTC', tr[2]),
                                 ('Training loss PER', tr[3]), ('Validation loss', va[0]), ('Validation loss inpainting', va[1]),
                                 ('Validation loss CTC', va[2]), ('Validation loss PER', va[3])):
                    tb.add_scalar(tag, float(val), epoch_counter)
                for tag, (kind, t) in (summaries or {}).items():       # tb_writer.add_summary(summaries, epoch) (training.py:352)
                    for i in range(t.shape[0]):
                        if kind == 'image':
                            img = t[i, :, :, 0]
                            img = (img - img.min()) / (img.max() - img.min()).clamp(min=1e-12)   # tf.summary.image scales to [0, 255]
                            tb.add_image('%s/image/%d' % (tag, i), img.unsqueeze(0), epoch_counter)
                        else:
                            tb.add_audio('%s/audio/%d' % (tag, i), t[i].clamp(-1, 1).unsqueeze(0), epoch_counter, sample_rate=16000)
                tb.flush()
            print('')
            log.write('{:d}\t{:.6f}\t{:.6f}|{:.6f}|{:.6f}\t{:.6f}\t{:.6f}|{:.6f}|{:.6f}\t{:.6f}\t[{:.2f}]\n'
                      .format(epoch_counter, lr, tr[0], tr[1], tr[2], tr[3], va[0], va[1], va[2], va[3], epoch_duration))
            log.flush()
        if stop or cneg_epochs >= config['n_earlystop_epochs']:
            break
    if rank == 0:
        if cneg_epochs >= config['n_earlystop_epochs']:
            print('+---- Done training: early stopped ----+')
        else:
            print('+---- Done training: epoch limit reached ----+')
        print('Total training time: {:.2f} s'.format(time() - train_start_time))
        print('{:d} epochs, {:d} steps.'.format(epoch_counter, tot_step))
        print('Best validation checkpoint: {:d} ({:d}) - Loss: {:.5f}'.format(best_val_checkpoint[0], best_val_checkpoint[1],
                                                                               best_val_loss))
        log.close()
    return model
