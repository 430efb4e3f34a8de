"""Phone-recognition BLSTM (models_asr.StackedBLSTMModel, models_asr.py:18-184) on the B200 hot path.

Same constructor signature and attribute names as the reference class: power-2 spectrogram (x mask when
apply_mask) -> log-mel-80 -> normalisation -> (concat video) -> stacked BLSTM -> `logits` head -> mean CTC loss,
Adam / SGD / Momentum.  Every stage reuses the kernels of the inpainting path: the fused front end's `fbanks`
variant, avsi_features_to_x0, the projection GEMMs, the recurrence kernels, the CTC kernel.  The beam-search
decoder of models_asr.py:132-135 runs on the host (avsi_ctc_beam_search_host), off the training step."""
import torch

from . import _lib
from . import audio_processing as ap
from .models import StackedBLSTMModel as _InpaintingModel
from .models import StackedBLSTMSSNNCTCLossModel as _MTLModel

_p = _lib.ptr


class StackedBLSTMModel(_InpaintingModel):
    MTL = False
    GRAD_SCALE = 256.0

    def __init__(self, sequence_lengths, labels_lengths, target_sources, masks, labels, audio_feat_mean, audio_feat_std,
                 dropout_rate, config, audio_features=None, video_features=None, input='a', apply_mask=False,
                 is_training=True, device='cuda', process_group=None):
        self.apply_mask = bool(apply_mask)
        self.num_mel_bins = 80
        cfg = dict(config)
        self._asr_classes = cfg['num_asr_labels']
        # the engine's head is the `logits` variable of models_asr.py:118-124: [2H, num_classes]
        cfg['audio_feat_dim'] = self.num_mel_bins
        super(StackedBLSTMModel, self).__init__(sequence_lengths, target_sources, masks, audio_feat_mean, audio_feat_std,
                                                dropout_rate, cfg, audio_features=audio_features,
                                                video_features=video_features, input=input, is_training=is_training,
                                                device=device, process_group=process_group,
                                                _out_dim=self._asr_classes)
        self.num_classes = self._asr_classes
        self.feed(labels_lengths=labels_lengths, labels=labels)
        m = ap.linear_to_mel_weight_matrix(self.num_mel_bins, 257, 16000, 125, 7600)
        self._mel = torch.tensor(m, dtype=torch.float32, device=self.device).contiguous()

    # ---- front end: models_asr.py:30-37 -----------------------------------------------------------------------
    def _front(self, want_stft=False):
        if 'front' in self._cache:
            return self._cache['front']
        wav, masks, mean, std = self._need('target_sources', 'masks', 'audio_features_mean', 'audio_features_std')
        B, T = wav.shape[0], masks.shape[1]
        video = self._fed.get('video_features') if self.input_type in ('v', 'av') else None
        ws = self.engine.workspace(T, B, self.is_training)
        L = self.engine.layout
        res = ap.fused_features(wav, self.frame_len, self.hop, T=T, F=257, mask=masks if self.apply_mask else None,
                                power=2.0, log=False, want_spec=False, mel=self._mel, mel_masked=self.apply_mask)
        fb = res['logmel']
        lib = _lib.load()
        if self.input_type == 'v':
            raise NotImplementedError("input='v' of the ASR model is not on the hot path")
        _lib.check(lib.avsi_features_to_x0(_p(fb), _p(mean), _p(std), _p(video), B, T, self.num_mel_bins,
                                           0 if video is None else video.shape[2], _p(ws['x0']), L.k0p,
                                           _lib.stream_ptr()), 'avsi_features_to_x0')
        out = {'ws': ws, 'B': B, 'T': T, 'target_fbanks': fb, 'target_spec_norm': None, 'target_stft': None, 'hole': None}
        self._cache['front'] = out
        return out

    @property
    def target_fbanks_norm(self):
        fr = self._front()
        mean, std = self._need('audio_features_mean', 'audio_features_std')
        return (fr['target_fbanks'] - mean) / std

    @property
    def inference(self):
        return self._bt(0, self.num_classes)

    # ---- loss: models_asr.py:140-160 ----------------------------------------------------------------------------
    def _loss_pass(self, want_grad, want_pred=False):
        key = 'loss_grad' if want_grad else 'loss'
        if key in self._cache:
            return self._cache[key]
        if 'loss_grad' in self._cache:
            return self._cache['loss_grad']
        lib = _lib.load()
        fr = self._front()
        ws, B, T = fr['ws'], fr['B'], fr['T']
        logits = self._logits()
        labels, lab_len, seq = self._need('labels', 'labels_lengths', 'sequence_lengths')
        L = self.engine.layout
        Lmax = labels.shape[1]
        nbytes = int(lib.avsi_ctc_workspace_bytes(B, T, Lmax))
        wsc = ws.get('ctc_ws')
        if wsc is None or wsc.numel() * 4 < nbytes:
            wsc = ws['ctc_ws'] = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=self.device)
        nll = torch.empty(B, dtype=torch.float32, device=self.device)
        dl = ws['dlogits'] if (want_grad and 'dlogits' in ws) else None
        with _lib.span('ctc'):
            # dlogits = GRAD_SCALE * d nll_b / d logits: softmax - posterior entries are ~1/C, the loss scale keeps
            # the back-propagated fp16 activations gradients out of the subnormal range; 1 / (GRAD_SCALE * B) is
            # applied by the optimiser's grad_unscale
            _lib.check(lib.avsi_ctc_loss(_p(logits), L.nop, 0, self.num_classes, _p(labels), Lmax, _p(lab_len), _p(seq),
                                         B, T, self.GRAD_SCALE, None, _p(nll), _p(dl), L.nop, 0, _p(wsc), _lib.stream_ptr()),
                       'avsi_ctc_loss')
        self._sums.zero_()
        self._sums[4] = nll.double().sum()
        self._sums[5] = float(B)
        out = {'ctc_nll': nll, 'sums': self._sums, 'prediction': None}
        self._cache[key] = out
        return out

    @property
    def ctc_loss(self):
        return self._loss_pass(False)['ctc_nll'].mean()

    @property
    def loss_func(self):
        return self.ctc_loss

    @property
    def loss_hole(self):
        return torch.zeros((), device=self.device)

    def _grad_unscale(self, out, world):
        return 1.0 / (self.GRAD_SCALE * self._front()['B'] * world), None

    per = _MTLModel.per                      # edit distance of the decoding against the labels (models_asr.py:162-166)
    decoder = 'beam'
    BEAM_WIDTH = 100                         # tf.nn.ctc_beam_search_decoder default (models_asr.py:132-135)
    _decode = _MTLModel._decode

    @property
    def decoding(self):
        return self._decode(0)
