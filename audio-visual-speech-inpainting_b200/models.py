"""Speech-inpainting BLSTM models with the reference's class names, constructor signatures and
attribute names (models.py:11-237 StackedBLSTMModel = A-SI / V-SI / AV-SI; models.py:1741-2047
StackedBLSTMSSNNCTCLossModel = the runnable multi-task model).

The reference builds a TF graph once from placeholders and feeds it every step; here the model
object is built once from the config and fed every step with ``feed(...)`` (same names as the
reference's placeholders, training_ctc.py:67-77).  Attributes (``inference``, ``prediction``,
``loss``, ``loss_hole``, ``enhanced_sources`` ...) evaluate lazily on the fed batch and are
cached until the next ``feed``; ``train_op()`` runs forward + backward + optimiser update.
All arithmetic runs in the sm_100a kernels of libavsi_b200.so; there is no autograd and no
CPU fallback.
"""
import os
import sys

import numpy as np
import torch

from . import _lib
from . import audio_processing as ap
from .blstm import BLSTMEngine, NY
from .layout import init_canonical

_p = _lib.ptr


def edit_distance(hyp, ref):
    """Levenshtein distance of two label sequences (tf.edit_distance, normalize=False, on dense sequences)."""
    prev = list(range(len(ref) + 1))
    for i, h in enumerate(hyp, 1):
        cur = [i] + [0] * len(ref)
        for jx, r in enumerate(ref, 1):
            cur[jx] = min(prev[jx] + 1, cur[jx - 1] + 1, prev[jx - 1] + (h != r))
        prev = cur
    return prev[-1]


class StackedBLSTMModel(object):
    """
    Speech inpainting BLSTM model
    Input: log-compressed linear spectrogram of corrupted audio (+ landmark motion vectors).
    Model: stacked BLSTM.  Output: log-compressed linear spectrogram of restored audio.
    Loss: L1 (target_spectrogram - reconstructed_spectrogram)       [models.py:11-18]
    """
    MTL = False

    def __init__(self, sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate,
                 config, audio_features=None, video_features=None, input='a', is_training=True, device='cuda',
                 process_group=None, _out_dim=None, embeddings=None):
        if input not in ('a', 'v', 'av'):
            raise ValueError("input must be 'a', 'v' or 'av'")
        self.config = config
        self.audio_feat_dim = config['audio_feat_dim']
        self.video_feat_dim = config.get('video_feat_dim', 136)
        self.audio_len = config['audio_len']
        self.input_type = input
        self.net_dim = config['net_dim']
        if len(set(self.net_dim)) != 1:
            # the reference passes net_dim[0] to every layer (models.py:96-97)
            raise ValueError('net_dim must be uniform (the reference uses net_dim[0] for all layers)')
        self.num_layers = len(self.net_dim)
        self.optimizer_choice = config.get('optimizer_type', 'adam')
        self.starter_learning_rate = config.get('starter_learning_rate', 1e-3)
        self.updating_step = config.get('lr_updating_steps', 10000)
        self.learning_decay = config.get('lr_decay', 1.0)
        self.regularization = config.get('l2', 0.0)
        self.is_training = is_training
        self.batch_size = config.get('batch_size', 1)
        self.device = torch.device(device)
        self.process_group = process_group
        self.var_scope = ''
        self.frame_len = ap.ms_to_samples(24, 16000)      # models.py:31 window_size=24, step_size=12
        self.hop = ap.ms_to_samples(12, 16000)
        in_dim = {'a': self.audio_feat_dim, 'v': self.video_feat_dim,
                  'av': self.audio_feat_dim + self.video_feat_dim}[input]
        # audio_features given (models.py:34-37): the network's audio input is that tensor instead of the masked target
        # spectrogram -- the two-step model feeds the video-only model's prediction (models.py:258-262)
        self.external_audio_features = audio_features is not None
        # a per-utterance vector replicated over the frames and appended to the network input (models.py:1204-1206)
        self.base_in_dim = in_dim
        self.embedding_dim = 0 if embeddings is None else int(np.asarray(embeddings.shape)[-1])
        in_dim += self.embedding_dim
        self.num_classes = config['num_asr_labels'] if self.MTL else 0
        self.engine = BLSTMEngine(in_dim, self.net_dim[0], self.num_layers,
                                  self.audio_feat_dim if _out_dim is None else _out_dim, self.num_classes,
                                  device=self.device, dense=self._dense_layers())
        self.engine.load_canonical(init_canonical(self.engine.layout, seed=config.get('seed', 0)))
        self.global_step = 0
        self.dropout_seed = int(config.get('seed', 0))
        self._feeds = 0
        self._feed_step, self._feeds_at_step = -1, 0
        self._stale = False
        self._sums = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._cache = {}
        self._fed = {}
        self._widen = {}
        self.feed(sequence_lengths=sequence_lengths, target_sources=target_sources, masks=masks,
                  audio_features_mean=audio_feat_mean, audio_features_std=audio_feat_std,
                  dropout_rate=dropout_rate, video_features=video_features, audio_features=audio_features,
                  embeddings=embeddings)

    def _dense_layers(self):
        """Extra fully connected layers living in the engine's flat parameter buffer (the SSNN models' MLP)."""
        return ()

    # ---- feed contract (training_ctc.py:67-77, 264-275) -------------------------------------------
    _COMPACT = {torch.int16: 0, torch.uint8: 1, torch.bool: 1, torch.int32: 2}

    def _to_dev(self, x, dtype, name=None):
        if x is None:
            return None
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x))
        if dtype == torch.float32 and x.dtype in self._COMPACT:
            # storage dtypes (int16 / int32 samples, uint8 masks) cross PCIe as they are and are widened on the device
            src = x.view(torch.uint8) if x.dtype == torch.bool else x
            src = src.to(device=self.device, non_blocking=True).contiguous()
            buf = self._widen.get(name)
            if buf is None or buf.shape != src.shape:
                buf = self._widen[name] = torch.empty(src.shape, dtype=torch.float32, device=self.device)
            _lib.check(_lib.load().avsi_cast_to_f32(_p(src), self._COMPACT[x.dtype], src.numel(), _p(buf), _lib.stream_ptr()),
                       'avsi_cast_to_f32')
            return buf
        return x.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def feed(self, **kw):
        """Set the tensors of the next batch; unknown names raise.  None values are ignored."""
        f32 = ('target_sources', 'masks', 'audio_features_mean', 'audio_features_std', 'video_features', 'audio_features',
               'embeddings')
        i32 = ('sequence_lengths', 'labels_lengths', 'labels')
        for k, v in kw.items():
            if v is None:
                continue
            if k in f32:
                self._fed[k] = self._to_dev(v, torch.float32, k)
            elif k in i32:
                self._fed[k] = self._to_dev(v, torch.int32)
            elif k == 'dropout_rate':
                if not 0.0 <= float(v) < 1.0:
                    raise ValueError('dropout_rate must be in [0, 1)')
                self._fed[k] = float(v)
            else:
                raise KeyError('unknown feed name %r' % k)
        self._cache = {}
        self._stale = False
        self._feeds += 1                     # a new sess.run: a new dropout mask (tf.nn.dropout draws per run)
        return self

    def _dropout(self):
        """(rate, seed, offset) of the current feed, or None: tf.nn.dropout(rnn_outputs, rate) of models.py:117."""
        rate = self._fed.get('dropout_rate', 0.0)
        if rate <= 0.0:
            return None
        # Philox offset = (rank, global step, feeds since that step): distinct masks on every rank, and the stream
        # continues where it was after a checkpoint resume (global_step is restored, the feed counter is not)
        rank = 0
        if self.process_group is not None:
            import torch.distributed as dist
            rank = dist.get_rank(self.process_group)
        if self._feed_step != self.global_step:
            self._feed_step, self._feeds_at_step = self.global_step, self._feeds
        sub = (self._feeds - self._feeds_at_step) & 0xFF
        return (rate, self.dropout_seed, (rank << 48) | ((self.global_step & 0xFFFFFFFFFF) << 8) | sub)

    def dropout_keep_mask(self):
        """The keep mask of the current feed as bool [B, T, 2H] (fw units, then bw units), or None.  For parity tests."""
        d = self._dropout()
        if d is None:
            return None
        fr = self._front()
        from .blstm import HP, NY
        M, H = fr['T'] * fr['B'], self.net_dim[-1]
        ones = torch.ones(M, NY, dtype=torch.float16, device=self.device)
        keep = torch.empty(M, NY, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.load().avsi_dropout_f16(_p(ones), NY, _p(ones), NY, M, NY, d[0], d[1], d[2], _p(keep), 0,
                                                _lib.stream_ptr()), 'avsi_dropout_f16')
        k = keep.view(fr['T'], fr['B'], 2, HP)[:, :, :, :H].reshape(fr['T'], fr['B'], 2 * H)
        return k.permute(1, 0, 2).bool()

    def _need(self, *names):
        for n in names:
            if n not in self._fed:
                raise _lib.AvsiError('tensor %r has not been fed' % n)
        return [self._fed[n] for n in names]

    def _check_shapes(self, wav, masks, mean, std, seq, video):
        """What TensorFlow's shape inference / concat checks catch in the reference graph; here the kernels take raw
        pointers, so a mismatch would be an out-of-bounds read.  Values (labels, lengths) are clamped in the kernels."""
        B, F = wav.shape[0], self.audio_feat_dim
        if wav.dim() != 2 or masks.dim() != 3 or masks.shape[0] != B or masks.shape[2] != F:
            raise ValueError('target_sources must be [B,N] and masks [B,T,%d]; got %s and %s'
                             % (F, tuple(wav.shape), tuple(masks.shape)))
        if mean.numel() != F or std.numel() != F:
            raise ValueError('audio_features_mean / _std must hold %d values' % F)
        if seq.numel() != B:
            raise ValueError('sequence_lengths must hold one entry per utterance')
        if video is not None and (video.dim() != 3 or tuple(video.shape[:2]) != tuple(masks.shape[:2])
                                  or video.shape[2] != self.video_feat_dim):
            raise ValueError('video_features must be [B,T,%d] with the B, T of masks; got %s vs masks %s'
                             % (self.video_feat_dim, tuple(video.shape), tuple(masks.shape)))
        if self.MTL or 'labels' in self._fed:
            labels, lab_len = self._fed.get('labels'), self._fed.get('labels_lengths')
            if labels is not None and (labels.dim() != 2 or labels.shape[0] != B):
                raise ValueError('labels must be [B,Lmax]')
            if lab_len is not None and lab_len.numel() != B:
                raise ValueError('labels_lengths must hold one entry per utterance')

    # ---- front end (models.py:30-45) -----------------------------------------------------------------
    def _front(self, want_stft=False):
        key = 'front_stft' if want_stft else 'front'
        if key in self._cache:
            return self._cache[key]
        wav, masks, mean, std, seq = self._need('target_sources', 'masks', 'audio_features_mean',
                                                'audio_features_std', 'sequence_lengths')
        B = wav.shape[0]
        T = masks.shape[1]                   # = max(sequence_lengths) in the reference's feed
        video = self._fed.get('video_features') if self.input_type in ('v', 'av') else None
        if self.input_type != 'a' and video is None:
            raise _lib.AvsiError("video_features must be fed for input='%s'" % self.input_type)
        self._check_shapes(wav, masks, mean, std, seq, video)
        ws = self.engine.workspace(T, B, self.is_training)
        hole = None
        if self.MTL:
            hole = self._cache.setdefault('hole', torch.zeros(1, dtype=torch.float64, device=self.device))
            hole.zero_()
        L = self.engine.layout
        ext = self._fed.get('audio_features') if self.external_audio_features else None
        if self.external_audio_features and ext is None:
            raise _lib.AvsiError('audio_features must be fed (the model was built with precomputed audio features)')
        res = ap.fused_features(wav, self.frame_len, self.hop, T=T, F=self.audio_feat_dim, mean=mean, std=std,
                                mask=masks, video=video, power=1.0, log=True, want_stft=want_stft, want_spec=True,
                                xh_out=None if ext is not None else ws['x0'], ldx=L.k0p, hole_count=hole,
                                xh_video_only=(self.input_type == 'v'), xh_skip_pad=True)
        if ext is not None:
            # network input = [given audio features ; video] (models.py:34-45), no normalisation, no mask
            if tuple(ext.shape) != (B, T, self.audio_feat_dim):
                raise ValueError('audio_features must be [B,T,%d]; got %s' % (self.audio_feat_dim, tuple(ext.shape)))
            if self.input_type == 'v':
                raise ValueError("precomputed audio_features make no sense with input='v'")
            _lib.check(_lib.load().avsi_features_to_x0(_p(ext), None, None, _p(video), B, T, self.audio_feat_dim,
                                                       0 if video is None else video.shape[2], _p(ws['x0']), L.k0p,
                                                       _lib.stream_ptr()), 'avsi_features_to_x0')
        if self.embedding_dim:
            emb = self._embedding_vectors(res, B, T)
            _lib.check(_lib.load().avsi_tile_embedding(_p(emb), B, T, self.embedding_dim, _p(ws['x0']), L.k0p,
                                                       self.base_in_dim, _lib.stream_ptr()), 'avsi_tile_embedding')
        out = {'ws': ws, 'B': B, 'T': T, 'target_spec_norm': res['spec'], 'target_stft': res['stft'], 'hole': hole}
        self._cache[key] = out
        if want_stft:
            self._cache['front'] = out
        return out

    def _embedding_vectors(self, res, B, T):
        """[B, embedding_dim] f32 vectors to replicate over the frames: here the fed `embeddings` placeholder."""
        emb = self._fed.get('embeddings')
        if emb is None or tuple(emb.shape) != (B, self.embedding_dim):
            raise ValueError('embeddings must be fed as [B,%d]' % self.embedding_dim)
        return emb

    @property
    def target_spec_norm(self):
        return self._front()['target_spec_norm']

    @property
    def target_stft(self):
        return self._front(want_stft=True)['target_stft']

    @property
    def net_inputs(self):
        """fp32 [B,T,I] view of the network input (the kernels consume the fp16 time-major copy)."""
        fr = self._front()
        L = self.engine.layout
        x = fr['ws']['x0'].view(fr['T'], fr['B'], L.k0p)[:, :, :L.in_dim]
        return x.permute(1, 0, 2).float()

    # ---- network (models.py:89-125) ----------------------------------------------------------------
    def _logits(self):
        if 'logits' not in self._cache:
            fr = self._front()
            self._cache['logits'] = self.engine.forward(fr['ws'], dropout=self._dropout())
        return self._cache['logits']

    def _bt(self, lo, n):
        fr = self._front()
        lg = self._logits().view(fr['T'], fr['B'], -1)[:, :, lo:lo + n]
        return lg.permute(1, 0, 2).contiguous()

    @property
    def inference(self):
        return self._bt(0, self.audio_feat_dim)

    # ---- loss (models.py:127-159) --------------------------------------------------------------------
    def _loss_pass(self, want_grad, want_pred=True):
        """One launch of the loss kernel(s).  The training step asks for gradients only (want_pred=False): the
        [B,T,F] prediction tensor is then not written (a fifth of the kernel's HBM traffic)."""
        for key in (('loss_grad', 'loss_grad_nopred') if want_grad else ('loss', 'loss_grad')):
            c = self._cache.get(key)
            if c is not None and (c['prediction'] is not None or not want_pred):
                return c
        key = ('loss_grad' if want_pred else 'loss_grad_nopred') if want_grad else 'loss'
        if not want_grad and not want_pred and 'loss_grad_nopred' in self._cache:
            return self._cache['loss_grad_nopred']
        lib = _lib.load()
        fr = self._front()
        ws, B, T = fr['ws'], fr['B'], fr['T']
        F = self.audio_feat_dim
        # AVSI_FUSE_HEAD_L1=1 (experiment, off): in the training step the inpainting head and the masked-L1 loss run as ONE
        # kernel (avsi_head_l1) -- the fp32 logits are consumed in the GEMM's epilogue instead of being written and read back
        # (1.3 GB per step at B = 2048); only the phone head's columns are kept for the CTC kernel.  Measured SLOWER than the
        # two launches (1.04 vs 0.35 + 0.40 ms): the one-tile-per-CTA kernel has 8 epilogue warps per SM to hide the loss's
        # loads behind, the stand-alone loss kernel 64 (profiles/README.md).  Parity-tested all the same.
        fused = (want_grad and not want_pred and 'logits' not in self._cache and 'dlogits' in ws
                 and os.environ.get('AVSI_FUSE_HEAD_L1', '0') == '1')
        logits = ws['logits'] if fused else self._logits()
        masks, seq = self._need('masks', 'sequence_lengths')
        L = self.engine.layout
        self._sums.zero_()
        pred = torch.empty(B, T, F, dtype=torch.float32, device=self.device) if want_pred else None
        out = {'prediction': pred}
        scales = None
        if self.MTL:
            scales = torch.empty(4, dtype=torch.float32, device=self.device)
            # the hole normaliser sum(1-m) is GLOBAL (models.py:1947 on the whole batch): the local count of the front end
            # is cloned and all-reduced ONCE per feed (the reduced value is cached beside the local one, so re-evaluating
            # the loss on the same feed neither reduces twice nor issues a collective the other ranks do not)
            hole = fr['hole']
            world = 1
            if self.process_group is not None:
                import torch.distributed as dist
                world = dist.get_world_size(self.process_group)
                if fr.get('hole_global') is None:
                    fr['hole_global'] = hole.clone()
                    dist.all_reduce(fr['hole_global'], group=self.process_group)
                hole = fr['hole_global']
            _lib.check(lib.avsi_mtl_scales(_p(hole), B * world, float(self.ctc_loss_weight), _p(scales),
                                           _p(self.engine.guard), _lib.stream_ptr()), 'avsi_mtl_scales')
        dl = ws['dlogits'] if (want_grad and 'dlogits' in ws) else None
        if fused:
            x, yil = self.engine.forward(ws, dropout=self._dropout(), head=False)
            self.engine.head_l1(ws, x, yil, fr['target_spec_norm'], masks, seq, F, 1 if self.MTL else 0,
                                _p(scales) if self.MTL else self.engine.guard_scale_ptr, self._sums,
                                logits_from_col=(F // 4) * 4 if self.MTL else 1 << 30)
        else:
            with _lib.span('masked_l1'):
                _lib.check(lib.avsi_masked_l1(_p(logits), L.nop, _p(fr['target_spec_norm']), _p(masks), _p(seq), B, T, F,
                                              1 if self.MTL else 0, 1.0,
                                              _p(scales) if self.MTL else self.engine.guard_scale_ptr, _p(self._sums),
                                              _p(pred), _p(dl), L.nop, _lib.stream_ptr()), 'avsi_masked_l1')
        if self.MTL:
            labels, lab_len = self._need('labels', 'labels_lengths')
            Lmax = labels.shape[1]
            nbytes = int(lib.avsi_ctc_workspace_bytes(B, T, Lmax))
            wsc = ws.get('ctc_ws')
            if wsc is None or wsc.numel() * 4 < nbytes:
                wsc = ws['ctc_ws'] = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=self.device)
            nll = torch.empty(B, dtype=torch.float32, device=self.device)
            scale_ptr = (scales.data_ptr() + 4) if dl is not None else None
            with _lib.span('ctc'):
                _lib.check(lib.avsi_ctc_loss(_p(logits), L.nop, F, self.num_classes, _p(labels), Lmax, _p(lab_len),
                                             _p(seq), B, T, 1.0, scale_ptr, _p(nll), _p(dl), L.nop, F, _p(wsc),
                                             _lib.stream_ptr()), 'avsi_ctc_loss')
            out['ctc_nll'] = nll
            out['scales'] = scales
        out['sums'] = self._sums
        self._cache[key] = out
        return out

    @property
    def prediction(self):
        return self._loss_pass(False)['prediction']

    def _sum(self, i):
        return self._loss_pass(False, want_pred=False)['sums'][i]

    @property
    def loss_hole(self):
        return (self._sum(0) / self._sum(1)).float()

    @property
    def loss_valid(self):
        return (self._sum(2) / self._sum(3)).float()

    @property
    def loss_func(self):
        return (self._sum(4) / self._sum(5)).float()                     # models.py:151

    @property
    def reg_loss(self):
        if not self.regularization:
            return torch.zeros((), device=self.device)
        n = self.engine.layout.n_params_padded
        return 0.5 * (self.engine.theta[:n] ** 2).sum()                  # models.py:153-154 (padding is zero)

    @property
    def loss(self):
        return self.loss_func + self.regularization * self.reg_loss      # models.py:158

    @property
    def learning_rate(self):
        # exponential_decay(staircase=True); Adam is given the constant starter rate (models.py:165-168)
        return self.starter_learning_rate * self.learning_decay ** (self.global_step // self.updating_step)

    # ---- training step (models.py:161-179) ---------------------------------------------------------------
    def _grad_unscale(self, out, world):
        """(host factor, device factor) turning the scaled gradient sum into d loss / d theta."""
        fr = self._front()
        return 1.0 / (fr['B'] * world * fr['T'] * self.audio_feat_dim), None

    def compute_gradients(self):
        """forward + loss + backward; leaves the (scaled) gradient in engine.grad."""
        if not self.is_training:
            raise _lib.AvsiError('model was built with is_training=False')
        if self._stale or self._cache.get('bwd_done'):
            # the weights moved since these activations were computed (train_op without a new feed), or a previous
            # backward pass on this feed has already overwritten the stashed gates in place with dG: run again.
            # Reads between the two calls keep returning the values of the run that produced the update, as
            # sess.run([train_op, loss]) does
            self._cache = {k: v for k, v in self._cache.items() if k.startswith('front')}
            self._stale = False
        out = self._loss_pass(True, want_pred=False)
        self.engine.backward(self._front()['ws'])
        self._backward_extra(self._front())
        self._cache['bwd_done'] = True
        return out

    def _backward_extra(self, fr):
        """Gradients of variables outside the BLSTM stack (the SSNN models' MLP); engine.grad already holds the stack's."""
        return None

    def _reduce_gradients(self, out):
        """Data parallel: ONE all-reduce per step over the flat fp32 gradient with the loss scalars riding in its tail.
        Returns the world size."""
        if self.process_group is None:
            return 1
        import torch.distributed as dist
        from . import parallel
        world = dist.get_world_size(self.process_group)
        parallel.pack_loss_tail(self.engine.grad, self.engine.layout.n_params_padded, out['sums'])
        with _lib.span('grad_allreduce', nbytes=self.engine.grad.numel() * 4):
            parallel.all_reduce_flat(self.engine.grad, self.process_group)
        return world

    def train_op(self, _device_step=False):
        if self.optimizer_choice not in ('adam', 'sgd', 'momentum'):
            print('Optimizer must be either sgd, momentum or adam. Closing...')         # models.py:175-176
            sys.exit(1)
        out = self.compute_gradients()
        world = self._reduce_gradients(out)
        host, dev = self._grad_unscale(out, world)
        if self.optimizer_choice == 'adam':
            # Adam is given the constant starter rate (models.py:168); the decayed rate applies to sgd / momentum only
            self.engine.adam_step(lr=self.starter_learning_rate, grad_unscale=host, unscale_dev=dev,
                                  l2=self.regularization, device_step=_device_step)
        else:
            self.engine.sgd_step(self.learning_rate, 0.9 if self.optimizer_choice == 'momentum' else None,
                                 grad_unscale=host, unscale_dev=dev, l2=self.regularization)
        self.global_step += 1
        self._stale = True

    def capture_train_step(self):
        """The whole training step of the CURRENT feed's shapes -- front end, BLSTM stack, loss, BPTT, optimiser: ~40 kernel
        launches through ctypes -- captured once into a CUDA graph and replayed with one launch per step.  At the
        reference's own batch sizes (8 ... 32 utterances, training_ctc.py) the step is a 1500-link dependent chain of
        microsecond kernels and the host cannot always keep the launch queue ahead of it; a graph takes the host out.

        Returns `step(**feeds)`: copies the named tensors into the captured input buffers (same shapes), replays, and
        leaves `loss`, `loss_hole`, ... readable as after `train_op()`.  Requirements (raised, never silently eager):
        Adam (the decayed learning rate of sgd / momentum is a host value per step), dropout_rate == 0 (the dropout
        offset is a host counter), a single process, and at least one eager `train_op()` on these shapes beforehand
        (workspaces, tensor maps and kernel attributes are created on first use)."""
        if not self.is_training:
            raise _lib.AvsiError('model was built with is_training=False')
        if self.optimizer_choice != 'adam' or float(self._fed.get('dropout_rate', 0.0)) != 0.0 or self.process_group is not None:
            raise _lib.AvsiError('capture_train_step needs optimizer adam, dropout_rate 0 and no process group')
        if self.engine.step_count < 1:
            raise _lib.AvsiError('run one eager train_op() on a feed of these shapes before capturing')
        static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._fed.items()}
        self._fed = static
        self._widen = {}
        eng = self.engine
        eng.guard[3] = eng.step_count                       # the device-resident update count takes over from the host's
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        steps_before, host_count = self.global_step, eng.step_count
        with torch.cuda.stream(side):
            self._cache, self._stale = {}, False
            with torch.cuda.graph(graph, stream=side):
                self.train_op(_device_step=True)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.global_step, eng.step_count = steps_before, host_count   # capturing executes nothing
        captured_cache = self._cache

        def step(**feeds):
            for k, v in feeds.items():
                if v is None:
                    continue
                dst = static.get(k)
                if not torch.is_tensor(dst):
                    raise KeyError('feed %r is not part of the captured step' % k)
                src = v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError('captured step: %s must keep shape %s, got %s' % (k, tuple(dst.shape), tuple(src.shape)))
                dst.copy_(src, non_blocking=True)
            self._fed, self._cache = static, captured_cache
            graph.replay()
            self._feeds += 1
            self.global_step += 1
            eng.step_count += 1
            self._stale = True
        step.graph = graph
        return step

    def canonical_gradients(self, reduce=False):
        """d loss / d variable in the reference's canonical layout (float64 numpy), for parity tests.  reduce=True (data
        parallel): the gradient of the GLOBAL batch's loss, after the all-reduce train_op would issue."""
        out = self.compute_gradients()
        world = self._reduce_gradients(out) if reduce else 1
        host, dev = self._grad_unscale(out, world)
        if dev is not None:
            host = host * float(dev.item())
        return self.engine.export_canonical_grads(host / self.engine.guard_state()[1])

    # ---- waveform reconstruction (models.py:181-197) -----------------------------------------------------
    def _enhanced(self, oracle_phase):
        mean, std, masks = self._need('audio_features_mean', 'audio_features_std', 'masks')
        return ap.reconstruct_from(self.prediction, self.target_stft, mask=None if oracle_phase else masks,
                                   mean=mean, std=std, num_samples=self.audio_len)

    @property
    def enhanced_sources(self):
        return self._enhanced(False)

    @property
    def enhanced_sources_oracle_phase(self):
        return self._enhanced(True)

    # ---- TensorBoard summaries (models.py:199-219) -----------------------------------------------------------
    N_SUMMARY_SAMPLES = 10

    @staticmethod
    def _spec_image(x, n):
        """tf.map_fn(tf.image.flip_up_down, expand_dims(transpose(x, [0, 2, 1]), 3)): [B,T,F] -> [n,F,T,1], low bins at the bottom."""
        return x[:n].transpose(1, 2).flip(1).unsqueeze(3).contiguous()

    @property
    def summaries(self):
        """The tensors behind tf.summary.merge_all() of models.py:199-219, under the reference's tags: three spectrogram
        images (target, enhanced, mask; frequency axis flipped, at most 10 samples) and the target / enhanced waveforms
        normalised to their peak, at 16 kHz.  {tag: (kind, CUDA tensor)}; training.py writes them to the event file."""
        n = self.N_SUMMARY_SAMPLES
        masks, wav = self._need('masks', 'target_sources')
        enh = self.enhanced_sources
        out = {
            'summary/Target_spectrogram': ('image', self._spec_image(self.target_spec_norm, n)),
            'summary/Enhanced_spectrogram': ('image', self._spec_image(self.prediction, n)),
            'summary/Mask': ('image', self._spec_image(masks, n)),
            'summary/Target_audio': ('audio', (wav / wav.abs().amax(dim=1, keepdim=True))[:n]),
            'summary/Enhanced_audio': ('audio', (enh / enh.abs().amax(dim=1, keepdim=True))[:n]),
        }
        return out

    # ---- variables / checkpoints --------------------------------------------------------------
    def build_graph(self, var_scope=''):
        self.var_scope = var_scope

    def _scoped(self, name):
        return (self.var_scope + '/' + name) if self.var_scope else name

    @property
    def train_vars(self):
        return {self._scoped(k): v for k, v in self.engine.export_canonical().items()}

    @property
    def all_vars(self):
        v = self.train_vars
        v[self._scoped('Variable')] = np.asarray(self.global_step, np.int32)
        return v

    def assign_vars(self, variables):
        """Load canonical variables (names with or without the scope prefix)."""
        want = self.engine.layout.canonical_shapes()
        got = {}
        for k in want:
            for cand in (self._scoped(k), k):
                if cand in variables:
                    got[k] = np.asarray(variables[cand])
                    break
        self.engine.load_canonical(got)
        gs = variables.get(self._scoped('Variable'), variables.get('Variable'))
        if gs is not None:
            self.global_step = int(gs)
        self._cache = {}
        self._stale = False


class StackedBLSTMSSNNCTCLossModel(StackedBLSTMModel):
    """Multi-task model: inpainting head + phone-recognition head with CTC loss
    (models.py:1741-2047).  The speaker-embedding MLP of that class is dead w.r.t. the loss in
    the reference and is not built."""
    MTL = True

    def __init__(self, sequence_lengths, labels_lengths, target_sources, masks, labels, audio_feat_mean,
                 audio_feat_std, dropout_rate, config, audio_features=None, video_features=None, input='a',
                 apply_mask=False, is_training=True, device='cuda', process_group=None):
        self.ctc_loss_weight = config.get('ctc_loss', 1)
        super(StackedBLSTMSSNNCTCLossModel, self).__init__(
            sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate, config,
            audio_features=audio_features, video_features=video_features, input=input, is_training=is_training,
            device=device, process_group=process_group)
        self.feed(labels_lengths=labels_lengths, labels=labels)

    @property
    def inference(self):
        return self._bt(0, self.audio_feat_dim), self._bt(self.audio_feat_dim, self.num_classes)

    @property
    def ctc_loss(self):
        return self._loss_pass(False, want_pred=False)['ctc_nll'].mean()   # models.py:1950-1953

    @property
    def loss_func(self):
        return self.loss_hole + self.ctc_loss_weight * self.ctc_loss      # models.py:1955

    def _grad_unscale(self, out, world):
        return 1.0, out['scales'][2:3]

    @property
    def per(self):
        """Phone error rate per utterance: edit distance(decoded, labels) / label length (models.py:1718 uses
        tf.edit_distance on the beam-search output), host side, off the hot loop."""
        dec = self.decoding
        labels = self._fed['labels'].cpu().numpy()
        lens = self._fed['labels_lengths'].cpu().numpy()
        out = np.zeros(len(dec), np.float32)
        for b in range(len(dec)):
            hyp = [int(x) for x in dec[b] if x >= 0]
            ref = [int(x) for x in labels[b, :int(lens[b])]]
            out[b] = edit_distance(hyp, ref) / max(1, len(ref))
        return out

    BEAM_WIDTH = 20                          # models.py:1627, :2027
    decoder = 'beam'                         # 'greedy' = best-path decoding (cheaper monitoring)

    def _decode(self, col0):
        """`decoding` of the reference: tf.nn.ctc_beam_search_decoder(tm_logits, sequence_lengths, beam_width)
        (top path, merge_repeated=True) -> dense int32 [B, max decoded length] padded with -1.  Host side, as in
        TensorFlow; it runs when the driver asks for `decoding` / `per`, never inside train_op."""
        import ctypes
        fr = self._front()
        T, B, C = fr['T'], fr['B'], self.num_classes
        seq = np.ascontiguousarray(self._fed['sequence_lengths'].cpu().numpy(), np.int32)
        logits = self._logits()
        if self.decoder == 'greedy':
            best = logits.view(T, B, -1)[:, :, col0:col0 + C].argmax(dim=2).t().cpu().numpy()
            outs = []
            for b in range(B):
                prev, seqo = -1, []
                for k in best[b, :int(seq[b])]:
                    if k != prev and k != C - 1:
                        seqo.append(int(k))
                    prev = k
                outs.append(seqo)
        else:
            host = np.ascontiguousarray(logits.detach().cpu().numpy(), np.float32)       # [T*B, nop] time-major
            out = np.empty((B, T), np.int32)
            n = np.zeros(B, np.int32)
            vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            _lib.check(_lib.load().avsi_ctc_beam_search_host(vp(host), T, B, host.shape[1], col0, C, vp(seq), self.BEAM_WIDTH,
                                                             1, T, vp(out), vp(n), None, 0), 'avsi_ctc_beam_search_host')
            outs = [out[b, :n[b]].tolist() for b in range(B)]
        width = max([len(o) for o in outs] + [1])
        dense = -np.ones((B, width), np.int32)
        for b, o in enumerate(outs):
            dense[b, :len(o)] = o
        return dense

    @property
    def decoding(self):
        return self._decode(self.audio_feat_dim)


class StackedBLSTMEmbeddingModel(StackedBLSTMModel):
    """Inpainting BLSTM with an external speaker embedding (models.py:1120-1472, `integration_layer` = 0, the shipped
    default): the 512-d vector of the utterance (dataset_reader_emb.py:63-81) is replicated over the frames and
    concatenated to the network input (models.py:1204-1206).  No extra variables: the layer-0 kernel simply has
    I_0 = in_dim + 512 input rows.  integration_layer >= 1 (embedding injected after the first BLSTM layer,
    models.py:1228-1290) is not built."""

    def __init__(self, sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate, config,
                 audio_features=None, video_features=None, embeddings=None, input='a', is_training=True, device='cuda',
                 process_group=None):
        if embeddings is None:
            raise ValueError('StackedBLSTMEmbeddingModel needs `embeddings` [B,E]')
        if config.get('integration_layer', 0):
            raise NotImplementedError('integration_layer >= 1 (models.py:1228-1290) is not built')
        super(StackedBLSTMEmbeddingModel, self).__init__(
            sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate, config,
            audio_features=audio_features, video_features=video_features, input=input, is_training=is_training,
            device=device, process_group=process_group, embeddings=embeddings)


class StackedBLSTMSSNNModel(StackedBLSTMModel):
    """Speech inpainting BLSTM model with SSNN (models.py:718-1117, `integration_layer` = 0, the shipped default): a
    speaker embedding is computed FROM THE CORRUPTED INPUT by a small trainable network and appended to every frame of
    the BLSTM input.

        inp   = add_delta_features(audio_features, n_delta=1, N=2)          [B,T,2F]                 (models.py:819)
        l1    = leaky_relu(inp . W1 + b1, 0.3) ; l2 = leaky_relu(l1 . W2 + b2, 0.3) ; l3 = l2 . W3 + b3     (:821-825)
        emb_b = sum_t l3[b,t] m[b,t] / (sum_t m[b,t] + 1),  m = masks[:, :, 0]                      (:832-836)
        BLSTM input = concat(net_inputs, tile(emb))                                                 (:846-849)

    Variables `speaker_embedding/weights_{1,2,3}`, `biases_{1,2,3}` ([2F,200], [200,200], [200,200]) train with the rest:
    the three dense layers run on the tcgen05 GEMM forward and backward, the gradient of emb is the time-sum of
    dG_0 . W_ih0[embedding columns]."""
    EMB = 200

    def __init__(self, sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate, config,
                 audio_features=None, video_features=None, input='a', is_training=True, device='cuda', process_group=None):
        if config.get('integration_layer', 0):
            raise NotImplementedError('integration_layer >= 1 (models.py:879-925) is not built')
        self._F2 = 2 * config['audio_feat_dim']

        class _Shape(object):                          # the base constructor only reads embeddings.shape[-1]
            shape = (0, self.EMB)
        super(StackedBLSTMSSNNModel, self).__init__(
            sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate, config,
            audio_features=audio_features, video_features=video_features, input=input, is_training=is_training,
            device=device, process_group=process_group, embeddings=_Shape())

    def feed(self, **kw):
        kw.pop('embeddings', None)                    # the embedding is computed, never fed
        return super(StackedBLSTMSSNNModel, self).feed(**kw)

    def _dense_layers(self):
        E = self.EMB
        return (('speaker_embedding/weights_1', 'speaker_embedding/biases_1', self._F2, E),
                ('speaker_embedding/weights_2', 'speaker_embedding/biases_2', E, E),
                ('speaker_embedding/weights_3', 'speaker_embedding/biases_3', E, E))

    def _embedding_vectors(self, res, B, T):
        from .blstm import gemm
        lib, st = _lib.load(), _lib.stream_ptr
        eng, E, F = self.engine, self.EMB, self.audio_feat_dim
        M = B * T
        masks = self._need('masks')[0]
        ext = self._fed.get('audio_features') if self.external_audio_features else None
        # audio_features = target_spec_norm * masks (models.py:733-734) unless given
        af = ext if ext is not None else ap.fused_features(
            self._need('target_sources')[0], self.frame_len, self.hop, T=T, F=F, mean=self._need('audio_features_mean')[0],
            std=self._need('audio_features_std')[0], mask=masks, power=1.0, log=True, want_spec=False, want_feat=True)['feat']
        inp = ap.add_delta_features(af[:, :, :F], n_delta=1, N=2)                   # [B,T,2F] f32
        kp = eng.layout.index['dw0'][1][1]
        s = self._ssnn = {'B': B, 'T': T}
        new = lambda *shape, **kw: torch.empty(*shape, device=self.device, **kw)
        s['a0'] = new(M, kp, dtype=torch.float16)
        _lib.check(lib.avsi_cast_pad_f16(_p(inp), self._F2, self._F2, _p(s['a0']), kp, M, st()), 'avsi_cast_pad_f16')
        x, ldx, K = s['a0'], kp, self._F2
        for k in range(3):
            z = s['z%d' % k] = new(M, E, dtype=torch.float32)
            gemm(_p(x), ldx, _p(eng.half['dw%d' % k]), eng.layout.index['dw%d' % k][1][1], _p(z), E,
                 _p(eng.view(eng.theta, 'db%d' % k)), M, E, K, 0, 1, tag='gemm_ssnn')
            if k < 2:
                a = s['a%d' % (k + 1)] = new(M, E, dtype=torch.float16)
                _lib.check(lib.avsi_leaky_relu(_p(z), M * E, 0.3, _p(a), st()), 'avsi_leaky_relu')
                x, ldx, K = a, E, E
        emb = new(B, E, dtype=torch.float32)
        s['inv_den'] = new(B, dtype=torch.float32)
        _lib.check(lib.avsi_masked_time_mean(_p(s['z2']), _p(masks), masks.shape[2], B, T, E, _p(emb), _p(s['inv_den']), st()),
                   'avsi_masked_time_mean')
        s['emb'] = emb
        return emb

    @property
    def speaker_embedding(self):
        """[B,200] (models.py:836); `speaker_embedding_ext` = the per-frame outputs times the mask (:833)."""
        self._front()
        return self._ssnn['emb']

    @property
    def speaker_embedding_ext(self):
        self._front()
        s = self._ssnn
        m = self._need('masks')[0][:, :, :1]
        return s['z2'].view(s['B'], s['T'], self.EMB) * m

    def _backward_extra(self, fr):
        from .blstm import A_IL, NG, gemm, pick_split_k
        lib, st = _lib.load(), _lib.stream_ptr
        eng, E, s = self.engine, self.EMB, self._ssnn
        B, T = s['B'], s['T']
        M = B * T
        L = eng.layout
        ws = fr['ws']
        masks = self._need('masks')[0]
        new = lambda *shape, **kw: torch.empty(*shape, device=self.device, **kw)
        # d emb: columns [base_in_dim, +E) of dX_0 = dG_0 . W_ih0, summed over the frames (the embedding was replicated)
        dxe = new(M, E, dtype=torch.float32)
        wT = eng.half['wihT0'].data_ptr() + self.base_in_dim * NG * 2               # rows = input columns of layer 0
        gemm(_p(ws['G'][0]), NG, wT, NG, _p(dxe), E, None, M, E, NG, 0, 1, tag='gemm_ssnn', layout=A_IL)
        demb = new(B, E, dtype=torch.float32)
        _lib.check(lib.avsi_time_sum(_p(dxe), E, T, B, E, _p(demb), st()), 'avsi_time_sum')
        d = new(M, E, dtype=torch.float16)                                          # d l3 (batch-major rows)
        _lib.check(lib.avsi_masked_time_mean_bwd(_p(demb), _p(masks), masks.shape[2], _p(s['inv_den']), B, T, E, _p(d), st()),
                   'avsi_masked_time_mean_bwd')
        g = eng.grad
        for k in (2, 1, 0):
            a_in = s['a%d' % k]
            kp = L.index['dw%d' % k][1][1]
            lda = a_in.shape[1]
            # dW_k^T [E, kp] += d^T . a_in ; db_k = colsum(d) ; d a_in = d . W_k^T
            gemm(_p(d), E, _p(a_in), lda, _p(eng.view(g, 'dw%d' % k)), kp, None, E, lda if k else kp, M, 1, 2,
                 pick_split_k(E, kp, M), tag='gemm_ssnn')
            _lib.check(lib.avsi_colsum_f16(_p(d), E, M, 0, E, _p(eng.view(g, 'db%d' % k)), st()), 'avsi_colsum_f16')
            if k > 0:
                da = new(M, E, dtype=torch.float32)
                gemm(_p(d), E, _p(eng.half['dwT%d' % k]), E, _p(da), E, None, M, E, E, 0, 1, tag='gemm_ssnn')
                d = new(M, E, dtype=torch.float16)
                _lib.check(lib.avsi_leaky_relu_bwd(_p(s['z%d' % (k - 1)]), _p(da), M * E, 0.3, _p(d), st()), 'avsi_leaky_relu_bwd')


class StackedBLSTM2StepsModel(object):
    """2-steps speech inpainting BLSTM model (models.py:240-317): a video-only model `v-blstm` predicts the spectrogram
    from the landmark motion vectors; its prediction replaces the masked spectrogram as the audio input of the
    audio-visual model `av-blstm-twosteps`, which is the one that is trained (train_vars = the av model's variables
    only, models.py:295; the video model is restored from `model_ckp_vnet`, training.py:115-116,153-156)."""
    MTL = False

    def __init__(self, sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std, dropout_rate, config,
                 video_features, is_training=True, device='cuda', process_group=None):
        self.config = config
        # the video model only ever runs forward here (its variables are not in train_vars): inference workspace
        self.video_model = StackedBLSTMModel(sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std,
                                             0.0, config, video_features=video_features, input='v', is_training=False,
                                             device=device)
        self.video_model.build_graph('v-blstm')
        self.av_model = StackedBLSTMModel(sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std,
                                          dropout_rate, config, audio_features=self.video_model.prediction,
                                          video_features=video_features, input='av', is_training=is_training, device=device,
                                          process_group=process_group)
        self.av_model.build_graph('av-blstm-twosteps')
        self.is_training = is_training
        self.var_scope = ''
        self.device = self.av_model.device
        self.optimizer_choice = self.av_model.optimizer_choice

    def feed(self, **kw):
        vkw = {k: v for k, v in kw.items() if k != 'dropout_rate'}
        self.video_model.feed(dropout_rate=0.0, **vkw)
        self.av_model.feed(audio_features=self.video_model.prediction, **kw)
        return self

    def build_graph(self, var_scope=''):
        self.var_scope = var_scope

    # models.py:281-296: everything but video_prediction is the av model's
    video_prediction = property(lambda self: self.video_model.prediction)
    engine = property(lambda self: self.av_model.engine)

    def __getattr__(self, name):
        if name in ('target_spec_norm', 'inference', 'prediction', 'loss', 'loss_func', 'loss_hole', 'loss_valid',
                    'learning_rate', 'global_step', 'enhanced_sources', 'enhanced_sources_oracle_phase', 'train_vars',
                    'train_op', 'compute_gradients', 'canonical_gradients', 'dropout_keep_mask', 'summaries', '_fed',
                    '_loss_pass'):
            return getattr(self.av_model, name)
        raise AttributeError(name)

    @property
    def all_vars(self):
        v = dict(self.av_model.all_vars)
        v.update(self.video_model.train_vars)
        return v

    def assign_vars(self, variables):
        """Both sub-models when the checkpoint holds them; a checkpoint of the video model alone (`model_ckp_vnet`)
        or of the av model alone restores that part."""
        done = 0
        for m in (self.video_model, self.av_model):
            if any(k.startswith(m.var_scope + '/') for k in variables):
                m.assign_vars(variables)
                done += 1
        if not done:
            raise KeyError('no variable of scope v-blstm/ or av-blstm-twosteps/')
        self.feed()                                   # the av model's audio input is the video model's (new) prediction


# In the reference, StackedBLSTMCTCLossModel.inference is broken (models.py:1565-1566 uses an undefined
# attribute); its intended semantics are those of the SSNN-CTC class without the embedding.
StackedBLSTMCTCLossModel = StackedBLSTMSSNNCTCLossModel

MODEL_REGISTRY = {
    'a-blstm': (StackedBLSTMModel, 'a'), 'v-blstm': (StackedBLSTMModel, 'v'), 'av-blstm': (StackedBLSTMModel, 'av'),
    'a-blstm-ctc': (StackedBLSTMCTCLossModel, 'a'), 'v-blstm-ctc': (StackedBLSTMCTCLossModel, 'v'),
    'av-blstm-ctc': (StackedBLSTMCTCLossModel, 'av'),
    'a-blstm-ssnn-ctc': (StackedBLSTMSSNNCTCLossModel, 'a'), 'v-blstm-ssnn-ctc': (StackedBLSTMSSNNCTCLossModel, 'v'),
    'av-blstm-ssnn-ctc': (StackedBLSTMSSNNCTCLossModel, 'av'),
    # training_emb.py:98-110 / training_ctc.py:80-136
    'a-blstm-emb': (StackedBLSTMEmbeddingModel, 'a'), 'v-blstm-emb': (StackedBLSTMEmbeddingModel, 'v'),
    'av-blstm-emb': (StackedBLSTMEmbeddingModel, 'av'),
    'av-blstm-twosteps': (StackedBLSTM2StepsModel, 'av'),
    'a-blstm-ssnn': (StackedBLSTMSSNNModel, 'a'), 'v-blstm-ssnn': (StackedBLSTMSSNNModel, 'v'),
    'av-blstm-ssnn': (StackedBLSTMSSNNModel, 'av'),
}
