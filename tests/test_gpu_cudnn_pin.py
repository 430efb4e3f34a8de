"""The stacked BLSTM of this repo against cuDNN's own LSTM on the same GPU.

The reference trains through tf.contrib.cudnn_rnn.CudnnLSTM (models.py:95-104): its recurrent arithmetic IS cuDNN's
cudnnRNNForwardTraining / BackwardData / BackwardWeights.  TensorFlow cannot run here, but torch.nn.LSTM on a CUDA device
binds the same cuDNN entry points (torch.backends.cudnn), so this test holds the hand-written kernels -- forward outputs
and every parameter gradient -- to the library the reference executes, in fp32, on identical inputs and weights.  What it
does NOT pin is TensorFlow's mapping between its canonical checkpoint layout and cuDNN's parameter order (gate order i, c, f, o
-> cuDNN i, f, c, o; one TF bias = the sum of cuDNN's two; forget_bias 0 in the cuDNN-compatible cell): that mapping is the
one used throughout (`oracle/blstm.py`, `layout.py`) and is applied here to load cuDNN.  Test infrastructure only."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from test_gpu_model import _build

pytestmark = pytest.mark.gpu

TOL = 2e-3


def _cudnn_stack(canon, in_dim, hidden, n_layers, device):
    from oracle import blstm as oblstm
    lstm = torch.nn.LSTM(in_dim, hidden, num_layers=n_layers, batch_first=True, bidirectional=True).to(device).float()
    H = hidden
    perm = np.concatenate([np.arange(0, H), np.arange(2 * H, 3 * H), np.arange(H, 2 * H), np.arange(3 * H, 4 * H)])
    with torch.no_grad():
        for l in range(n_layers):
            i_l = in_dim if l == 0 else 2 * H
            for d, suf in (('fw', ''), ('bw', '_reverse')):
                k = np.asarray(canon[oblstm.cell_prefix(l, d) + '/kernel'], np.float64)[:, perm]     # TF i,c,f,o -> cuDNN i,f,c,o
                b = np.asarray(canon[oblstm.cell_prefix(l, d) + '/bias'], np.float64)[perm]
                getattr(lstm, 'weight_ih_l%d%s' % (l, suf)).copy_(torch.tensor(k[:i_l].T))
                getattr(lstm, 'weight_hh_l%d%s' % (l, suf)).copy_(torch.tensor(k[i_l:].T))
                getattr(lstm, 'bias_ih_l%d%s' % (l, suf)).copy_(torch.tensor(b))
                getattr(lstm, 'bias_hh_l%d%s' % (l, suf)).zero_()
    lstm.flatten_parameters()
    return lstm, perm


@pytest.mark.parametrize('model_name,B,audio_len', [('av-blstm', 6, 9600),        # mma.sync recurrence kernels, T = 50
                                                    ('a-blstm', 230, 3840),      # tcgen05 kernels, ragged second batch tile
                                                    ('av-blstm', 256, 48000),    # tcgen05 kernels at the GRID length, T = 250
                                                    ('av-blstm', 2048, 48000),   # the launch bench.py times: 2048 DISTINCT utterances
                                                    ('av-blstm', 256, 320000)])  # 20 s utterances (configs[4]): 1667-step chains
def test_blstm_stack_matches_cudnn_forward_and_gradients(model_name, B, audio_len):
    from oracle import blstm as oblstm
    assert torch.backends.cudnn.is_available() and torch.backends.cudnn.enabled
    model, batch, canon, inp = _build(model_name, B, audio_len, seed=B + 2)
    dev = model.device
    H, L = 250, 3
    x = model.net_inputs.float()                                   # [B,T,I]: the fp16 network input both sides consume
    T, I = x.shape[1], x.shape[2]
    target = model.target_spec_norm.float()
    seq = torch.as_tensor(batch['seq_len'], device=dev)
    lstm, perm = _cudnn_stack(canon, I, H, L, dev)
    w = torch.tensor(np.asarray(canon['logits/weights'], np.float32), device=dev, requires_grad=True)
    bb = torch.tensor(np.asarray(canon['logits/biases'], np.float32), device=dev, requires_grad=True)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        rnn, _ = lstm(x)                                           # cuDNN
        inference = (rnn.reshape(B * T, 2 * H) @ w + bb).reshape(B, T, -1)
        sm = (torch.arange(T, device=dev)[None, :] < seq[:, None]).float()[:, :, None]
        loss = (target - sm * inference).abs().mean()              # models.py:135-151
        loss.backward()
    assert rel_l2(model.inference.cpu().numpy(), inference.detach().cpu().numpy()) < TOL
    assert abs(float(model.loss) - float(loss.detach())) < 1e-4 * abs(float(loss.detach()))
    grads = model.canonical_gradients()
    inv = np.argsort(perm)
    ref = {'logits/weights': w.grad.cpu().numpy(), 'logits/biases': bb.grad.cpu().numpy()}
    for l in range(L):
        for d, suf in (('fw', ''), ('bw', '_reverse')):
            gih = getattr(lstm, 'weight_ih_l%d%s' % (l, suf)).grad.cpu().numpy()
            ghh = getattr(lstm, 'weight_hh_l%d%s' % (l, suf)).grad.cpu().numpy()
            gb = getattr(lstm, 'bias_ih_l%d%s' % (l, suf)).grad.cpu().numpy()
            ref[oblstm.cell_prefix(l, d) + '/kernel'] = np.concatenate([gih.T, ghh.T], 0)[:, inv]
            ref[oblstm.cell_prefix(l, d) + '/bias'] = gb[inv]
    assert set(ref) == set(grads)
    ga = np.concatenate([np.asarray(grads[k], np.float64).ravel() for k in sorted(ref)])
    gb_ = np.concatenate([np.asarray(ref[k], np.float64).ravel() for k in sorted(ref)])
    worst = max((rel_l2(grads[k], ref[k]), k) for k in ref if np.linalg.norm(ref[k]) > 0)
    print('cudnn pin %s B=%d T=%d: inference rel-L2 %.2e, loss %.6f vs %.6f, gradient rel-L2 %.2e, worst variable %.2e (%s)'
          % (model_name, B, T, rel_l2(model.inference.cpu().numpy(), inference.detach().cpu().numpy()), float(model.loss),
             float(loss.detach()), rel_l2(ga, gb_), worst[0], worst[1].split('/')[-3] + '/' + worst[1].split('/')[-1]))
    assert rel_l2(ga, gb_) < TOL, 'gradient vs cuDNN: %.3e (worst %s)' % (rel_l2(ga, gb_), worst)
    assert worst[0] < 2 * TOL, worst
