"""Oracle (test infrastructure): stacked-BLSTM inpainting models on the CPU.

Restates models.py:89-159 (StackedBLSTMModel: inference / prediction / loss)
and models.py:1873-1963 (StackedBLSTMSSNNCTCLossModel: two heads, hole-only
prediction, loss_hole + w * CTC).  LSTM cell = tf.contrib.cudnn_rnn
CudnnCompatibleLSTMCell (LSTMBlockCell, forget_bias 0, no peephole):

    [i, j, f, o] = [x_t ; h_{t-1}] @ kernel + bias      (TF column order i, j, f, o)
    c_t = sigmoid(f) * c_{t-1} + sigmoid(i) * tanh(j)
    h_t = sigmoid(o) * tanh(c_t)

stack_bidirectional_dynamic_rnn without sequence_length (models.py:111-115):
every layer runs fw over t=0..T-1 and bw over t=T-1..0 on the full padded
sequence and the next layer's input is concat(h_fw, h_bw).

Written with torch float64 CPU ops so gradients come from autograd (checked
against finite differences and torch.nn.LSTM in tests/test_oracle_blstm.py).
Parameters use the reference's checkpoint names and layouts.
"""
import math

import numpy as np
import torch

from . import ctc as octc


def cell_prefix(layer, direction):
    return ('cudnn_lstm/stack_bidirectional_rnn/cell_%d/bidirectional_rnn/%s/cudnn_compatible_lstm_cell'
            % (layer, direction))


def param_shapes(in_dim, hidden, n_layers, out_dim=257, n_classes=0):
    """Ordered {name: shape} in the reference's canonical checkpoint layout."""
    shapes = {}
    for l in range(n_layers):
        i_l = in_dim if l == 0 else 2 * hidden
        for d in ('fw', 'bw'):
            shapes[cell_prefix(l, d) + '/kernel'] = (i_l + hidden, 4 * hidden)
            shapes[cell_prefix(l, d) + '/bias'] = (4 * hidden,)
    if n_classes:
        shapes['inpainting/weights'] = (2 * hidden, out_dim)
        shapes['inpainting/biases'] = (out_dim,)
        shapes['asr/weights'] = (2 * hidden, n_classes)
        shapes['asr/biases'] = (n_classes,)
    else:
        shapes['logits/weights'] = (2 * hidden, out_dim)
        shapes['logits/biases'] = (out_dim,)
    return shapes


def init_params(in_dim, hidden, n_layers, out_dim=257, n_classes=0, seed=1, bias_scale=0.0):
    """Seeded synthetic weights: Glorot-uniform kernels,
    truncated-normal heads (sigma = 1/sqrt(2H)), zero (or small random) biases."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shp in param_shapes(in_dim, hidden, n_layers, out_dim, n_classes).items():
        if name.endswith('/kernel'):
            lim = math.sqrt(6.0 / (shp[0] + shp[1]))
            params[name] = rng.uniform(-lim, lim, shp)
        elif name.endswith('/weights'):
            sd = 1.0 / math.sqrt(shp[0])
            w = rng.standard_normal(shp)
            bad = np.abs(w) > 2.0
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            params[name] = w * sd
        else:
            params[name] = bias_scale * rng.standard_normal(shp)
    return params


def lstm_direction(x, kernel, bias, reverse):
    """x [B,T,I] -> h [B,T,H]."""
    B, T, _ = x.shape
    H = kernel.shape[1] // 4
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        z = torch.cat([x[:, t], h], dim=1) @ kernel + bias
        i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(j)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def blstm_stack(x, params, n_layers):
    """models.py:106-115.  x [B,T,I] -> [B,T,2H]."""
    for l in range(n_layers):
        hf = lstm_direction(x, params[cell_prefix(l, 'fw') + '/kernel'], params[cell_prefix(l, 'fw') + '/bias'], False)
        hb = lstm_direction(x, params[cell_prefix(l, 'bw') + '/kernel'], params[cell_prefix(l, 'bw') + '/bias'], True)
        x = torch.cat([hf, hb], dim=2)
    return x


def to_torch(params, dtype=torch.float64, requires_grad=False):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in params.items()}


def sequence_mask(seq_len, T, dtype):
    return (torch.arange(T)[None, :] < torch.as_tensor(seq_len)[:, None]).to(dtype)


def _dropout(rnn, drop):
    """tf.nn.dropout(rnn_outputs, rate) (models.py:117, :1901; models_asr.py:120) with a GIVEN keep mask:
    drop = (keep [B,T,2H] of {0,1}, rate) -> rnn * keep / (1 - rate).  TF's own random stream is not reproducible
    outside TF; parity is on the semantics, with the mask the CUDA kernel drew."""
    if drop is None:
        return rnn
    keep, rate = drop
    return rnn * torch.as_tensor(np.asarray(keep, np.float64), dtype=rnn.dtype) / (1.0 - float(rate))


def forward_si(net_in, target, mask, seq_len, params, n_layers, l2=0.0, drop=None):
    """StackedBLSTMModel (models.py:89-159).  All inputs torch tensors.

    Returns dict(inference, prediction, loss, loss_func, loss_hole, loss_valid)."""
    B, T, F = target.shape
    rnn = _dropout(blstm_stack(net_in, params, n_layers), drop)
    inference = (rnn.reshape(B * T, -1) @ params['logits/weights'] + params['logits/biases']).reshape(B, T, F)
    prediction = sequence_mask(seq_len, T, target.dtype)[:, :, None] * inference           # :135-137
    ad = (target - prediction).abs()
    loss_hole = (ad * (1 - mask)).sum() / (1 - mask).sum()                                  # :144
    loss_valid = (ad * mask).sum() / mask.sum()                                             # :145
    loss_func = ad.mean()                                                                   # :151
    reg = sum(0.5 * (p ** 2).sum() for p in params.values()) if l2 else 0.0                 # :153-154
    return dict(inference=inference, prediction=prediction, loss=loss_func + l2 * reg,
                loss_func=loss_func, loss_hole=loss_hole, loss_valid=loss_valid, rnn=rnn)


def speaker_embedding(delta_inp, mask, params, slopes=None):
    """StackedBLSTMSSNNModel.speaker_embedding (models.py:800-842).  delta_inp = add_delta_features(audio_features, 1, 2)
    [B,T,2F] (a constant of the parameters), mask [B,T,F].  Returns (embedding [B,200], per-frame masked outputs [B,T,200]).

    slopes = (s1, s2), each [B*T,200] of {1, 0.3}: evaluate the two leaky ReLUs with a GIVEN activation pattern
    (z * slope) instead of the sign of this evaluation's own pre-activations.  Identical wherever the signs agree; lets a
    test separate the kink of the activation (a pre-activation within rounding error of 0 switches the derivative between
    1 and 0.3) from the arithmetic of an implementation."""
    B, T, _ = delta_inp.shape
    x = delta_inp.reshape(B * T, -1)

    def lrelu(z, k):
        if slopes is None:
            return torch.nn.functional.leaky_relu(z, 0.3)
        return z * torch.as_tensor(np.asarray(slopes[k], np.float64), dtype=z.dtype)
    l1 = lrelu(x @ params['speaker_embedding/weights_1'] + params['speaker_embedding/biases_1'], 0)        # :821-822
    l2 = lrelu(l1 @ params['speaker_embedding/weights_2'] + params['speaker_embedding/biases_2'], 1)       # :823-824
    l3 = l2 @ params['speaker_embedding/weights_3'] + params['speaker_embedding/biases_3']                 # :825
    emb_mask = mask[:, :, 0]                                                                               # :832
    ext = l3.reshape(B, T, -1) * emb_mask[:, :, None]                                                      # :833
    return ext.sum(1) / (emb_mask.sum(1) + 1)[:, None], ext                                                # :834-835


def forward_si_ssnn(net_in, delta_inp, target, mask, seq_len, params, n_layers, l2=0.0, slopes=None):
    """StackedBLSTMSSNNModel with integration_layer 0 (models.py:844-849): the embedding tiled over the frames and
    concatenated to the network input; everything after is forward_si."""
    emb, ext = speaker_embedding(delta_inp, mask, params, slopes)
    x = torch.cat([net_in, emb[:, None, :].expand(-1, net_in.shape[1], -1)], dim=2)
    out = forward_si(x, target, mask, seq_len, params, n_layers, l2=l2)
    out['speaker_embedding'] = emb
    out['speaker_embedding_ext'] = ext
    return out


def forward_mtl(net_in, target, mask, seq_len, labels, lab_len, params, n_layers, ctc_weight, l2=0.0, drop=None):
    """StackedBLSTMSSNNCTCLossModel (models.py:1873-1963), the runnable MTL model.
    Blank label = n_classes - 1 (tf.nn.ctc_loss convention)."""
    B, T, F = target.shape
    rnn = _dropout(blstm_stack(net_in, params, n_layers), drop)
    flat = rnn.reshape(B * T, -1)
    logits_ipt = (flat @ params['inpainting/weights'] + params['inpainting/biases']).reshape(B, T, F)
    logits_asr = (flat @ params['asr/weights'] + params['asr/biases']).reshape(B, T, -1)
    prediction = target * mask + logits_ipt * (1 - mask)                                    # :1927
    prediction = sequence_mask(seq_len, T, target.dtype)[:, :, None] * prediction           # :1929
    ad = (target - prediction).abs()
    loss_hole = (ad * (1 - mask)).sum() / (1 - mask).sum()                                  # :1947
    nll = octc.ctc_nll_torch(logits_asr.transpose(0, 1), labels, lab_len, seq_len)          # :1950-1953
    ctc_loss = nll.mean()
    loss_func = loss_hole + ctc_weight * ctc_loss                                           # :1955
    reg = sum(0.5 * (p ** 2).sum() for p in params.values()) if l2 else 0.0
    return dict(inference=logits_ipt, logits_asr=logits_asr, prediction=prediction,
                loss=loss_func + l2 * reg, loss_func=loss_func, loss_hole=loss_hole,
                ctc_loss=ctc_loss, ctc_nll=nll, rnn=rnn)


def forward_asr(net_in, seq_len, labels, lab_len, params, n_layers, l2=0.0, drop=None):
    """models_asr.StackedBLSTMModel (models_asr.py:87-160): BLSTM stack -> logits head -> mean CTC NLL."""
    B, T, _ = net_in.shape
    rnn = _dropout(blstm_stack(net_in, params, n_layers), drop)
    logits = (rnn.reshape(B * T, -1) @ params['logits/weights'] + params['logits/biases']).reshape(B, T, -1)
    nll = octc.ctc_nll_torch(logits.transpose(0, 1), labels, lab_len, seq_len)                # models_asr.py:146-148
    ctc_loss = nll.mean()
    reg = sum(0.5 * (p ** 2).sum() for p in params.values()) if l2 else 0.0
    return dict(inference=logits, ctc_loss=ctc_loss, ctc_nll=nll, loss=ctc_loss + l2 * reg, rnn=rnn)


def loss_and_grads(kind, inputs, params_np, n_layers, dtype=torch.float64, **kw):
    """Run forward + autograd.  Returns (outputs as numpy, grads {name: numpy})."""
    params = to_torch(params_np, dtype, requires_grad=True)
    tin = {k: (torch.tensor(np.asarray(v), dtype=dtype) if np.asarray(v).dtype.kind == 'f' else torch.tensor(np.asarray(v)))
           for k, v in inputs.items()}
    if kind == 'asr':
        out = forward_asr(tin['net_in'], tin['seq_len'], tin['labels'].long(), tin['lab_len'], params, n_layers, **kw)
    elif kind == 'si':
        out = forward_si(tin['net_in'], tin['target'], tin['mask'], tin['seq_len'], params, n_layers, **kw)
    elif kind == 'ssnn':
        out = forward_si_ssnn(tin['net_in'], tin['delta_inp'], tin['target'], tin['mask'], tin['seq_len'], params, n_layers, **kw)
    else:
        out = forward_mtl(tin['net_in'], tin['target'], tin['mask'], tin['seq_len'], tin['labels'].long(),
                          tin['lab_len'], params, n_layers, **kw)
    out['loss'].backward()
    grads = {k: (p.grad.numpy().copy() if p.grad is not None else np.zeros(tuple(p.shape))) for k, p in params.items()}
    outs = {k: v.detach().numpy() for k, v in out.items() if torch.is_tensor(v)}
    return outs, grads
