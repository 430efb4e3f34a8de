"""The BLSTM / CTC / Adam oracle is unpinned by the reference (no tests, TF absent): cross-check it
against independent implementations."""
import numpy as np
import torch

from oracle import adam, blstm, ctc


def test_lstm_matches_torch_nn_lstm():
    rng = np.random.default_rng(0)
    B, T, I, H = 3, 7, 5, 4
    params = blstm.init_params(I, H, 1, out_dim=6, seed=3, bias_scale=0.1)
    x = torch.tensor(rng.standard_normal((B, T, I)))
    tp = blstm.to_torch(params)
    y = blstm.blstm_stack(x, tp, 1)
    ref = torch.nn.LSTM(I, H, batch_first=True, bidirectional=True).double()
    perm = np.concatenate([np.arange(0, H), np.arange(2 * H, 3 * H), np.arange(H, 2 * H), np.arange(3 * H, 4 * H)])
    with torch.no_grad():
        for d, suf in (('fw', ''), ('bw', '_reverse')):
            k = params[blstm.cell_prefix(0, d) + '/kernel'][:, perm]      # TF i,j,f,o -> torch i,f,g,o
            b = params[blstm.cell_prefix(0, d) + '/bias'][perm]
            getattr(ref, 'weight_ih_l0' + suf).copy_(torch.tensor(k[:I].T))
            getattr(ref, 'weight_hh_l0' + suf).copy_(torch.tensor(k[I:].T))
            getattr(ref, 'bias_ih_l0' + suf).copy_(torch.tensor(b))
            getattr(ref, 'bias_hh_l0' + suf).zero_()
        yr, _ = ref(x)
    assert torch.allclose(y, yr, atol=1e-12)


def test_si_gradients_match_finite_differences():
    rng = np.random.default_rng(1)
    B, T, I, H, F = 2, 5, 6, 3, 4
    params = blstm.init_params(I, H, 2, out_dim=F, seed=5, bias_scale=0.1)
    mask = np.ones((B, T, F))
    mask[:, 2:4] = 0
    inputs = dict(net_in=rng.standard_normal((B, T, I)), target=rng.standard_normal((B, T, F)), mask=mask,
                  seq_len=np.array([5, 4]))
    outs, grads = blstm.loss_and_grads('si', inputs, params, 2)
    name = blstm.cell_prefix(0, 'bw') + '/kernel'
    eps = 1e-6
    for idx in [(0, 0), (3, 5), (7, 11)]:
        p2 = {k: v.copy() for k, v in params.items()}
        p2[name][idx] += eps
        lp = blstm.loss_and_grads('si', inputs, p2, 2)[0]['loss']
        p2[name][idx] -= 2 * eps
        lm = blstm.loss_and_grads('si', inputs, p2, 2)[0]['loss']
        assert abs((lp - lm) / (2 * eps) - grads[name][idx]) < 1e-7
    # padded frames of the second sample contribute |target| only and carry no gradient to the head bias
    assert outs['prediction'][1, 4].tolist() == [0.0] * F


def test_asr_mode_is_ctc_of_head_logits_and_differentiates():
    """models_asr.py:87-160 restated: loss = mean CTC NLL of the head logits; gradient checked by central differences
    and the NLL against torch's own CTC on the same logits."""
    rng = np.random.default_rng(2)
    B, T, I, H, C = 2, 9, 5, 3, 6
    params = blstm.init_params(I, H, 2, out_dim=C, seed=9, bias_scale=0.1)
    inputs = dict(net_in=rng.standard_normal((B, T, I)), seq_len=np.array([9, 7]), labels=np.array([[0, 1, 1], [4, 2, 0]]),
                  lab_len=np.array([3, 2]))
    outs, grads = blstm.loss_and_grads('asr', inputs, params, 2)
    lg = torch.tensor(outs['inference']).transpose(0, 1)
    ref = torch.nn.functional.ctc_loss(torch.log_softmax(lg, 2), torch.tensor(inputs['labels']), torch.tensor(inputs['seq_len']),
                                       torch.tensor(inputs['lab_len']), blank=C - 1, reduction='none')
    assert np.allclose(outs['ctc_nll'], ref.numpy(), atol=1e-9) and abs(outs['loss'] - ref.mean().item()) < 1e-9
    eps = 1e-6
    for name, idx in ((blstm.cell_prefix(1, 'fw') + '/kernel', (2, 7)), ('logits/weights', (4, 1)),
                      (blstm.cell_prefix(0, 'bw') + '/bias', (5,))):
        p2 = {k: v.copy() for k, v in params.items()}
        p2[name][idx] += eps
        lp = blstm.loss_and_grads('asr', inputs, p2, 2)[0]['loss']
        p2[name][idx] -= 2 * eps
        lm = blstm.loss_and_grads('asr', inputs, p2, 2)[0]['loss']
        assert abs((lp - lm) / (2 * eps) - grads[name][idx]) < 1e-7, name


def test_ctc_three_ways():
    rng = np.random.default_rng(0)
    T, B, C = 30, 3, 34
    lg = rng.standard_normal((T, B, C)) * 2
    labels = rng.integers(0, 33, (B, 50))
    labels[0, 1] = labels[0, 0]                    # repeated label needs a blank in between
    ll, sl = np.array([5, 8, 1]), np.array([30, 25, 12])
    nll, g = ctc.ctc_alpha_beta(lg, labels, ll, sl)
    x = torch.tensor(lg, requires_grad=True)
    n2 = ctc.ctc_nll_torch(x, labels, ll, sl)
    n2.sum().backward()
    assert np.allclose(nll, n2.detach().numpy(), atol=1e-10) and np.abs(g - x.grad.numpy()).max() < 1e-10
    x3 = torch.tensor(lg, requires_grad=True)
    n3 = torch.nn.functional.ctc_loss(torch.log_softmax(x3, 2), torch.tensor(labels), torch.tensor(sl),
                                      torch.tensor(ll), blank=33, reduction='none')
    n3.sum().backward()
    assert np.allclose(nll, n3.detach().numpy(), atol=1e-9) and np.abs(g - x3.grad.numpy()).max() < 1e-9
    assert np.all(g[25:, 1] == 0)                  # frames beyond sequence_length: zero gradient
    for lab in ([0, 1], [0, 0], [1]):
        lg2 = rng.standard_normal((4, 1, 3))
        bf = ctc.ctc_bruteforce(lg2[:, 0], lab)
        ab = ctc.ctc_alpha_beta(lg2, np.array([lab + [0] * (2 - len(lab))]), [len(lab)], [4])[0][0]
        assert abs(bf - ab) < 1e-10


def test_adam_tf_epsilon_hat_form():
    rng = np.random.default_rng(0)
    th, g = rng.standard_normal(10), rng.standard_normal(10)
    m = np.zeros(10)
    v = np.zeros(10)
    th1, m1, v1 = adam.adam_tf_step(th, g, m, v, 1)
    # first step of TF Adam: theta - lr * g / (|g| + eps * sqrt(1-b2)/(1-b1) ...) ~ lr * sign(g)
    assert np.allclose(th1, th - 1e-3 * np.sign(g), atol=1e-6)
    t = torch.tensor(th, requires_grad=True)
    opt = torch.optim.Adam([t], lr=1e-3, eps=1e-8)
    t.grad = torch.tensor(g)
    opt.step()
    assert np.allclose(th1, t.detach().numpy(), atol=1e-7)      # same up to where epsilon enters


def test_mtl_gradients_match_finite_differences():
    """models.py:1873-1963 restated (hole-normalised L1 + weighted mean CTC): central differences on both heads and a cell."""
    rng = np.random.default_rng(6)
    B, T, I, H, F, C = 2, 8, 5, 3, 4, 5
    params = blstm.init_params(I, H, 2, out_dim=F, n_classes=C, seed=11, bias_scale=0.1)
    mask = np.ones((B, T, F))
    mask[:, 2:5] = 0
    inputs = dict(net_in=rng.standard_normal((B, T, I)), target=rng.standard_normal((B, T, F)), mask=mask,
                  seq_len=np.array([8, 7]), labels=np.array([[0, 1, 1], [3, 2, 0]]), lab_len=np.array([3, 2]))
    outs, grads = blstm.loss_and_grads('mtl', inputs, params, 2, ctc_weight=0.3)
    assert abs(outs['loss'] - (outs['loss_hole'] + 0.3 * outs['ctc_loss'])) < 1e-12
    eps = 1e-6
    for name, idx in (('inpainting/weights', (1, 2)), ('asr/weights', (4, 3)), ('asr/biases', (1,)),
                      (blstm.cell_prefix(0, 'fw') + '/kernel', (6, 10))):
        p2 = {k: v.copy() for k, v in params.items()}
        p2[name][idx] += eps
        lp = blstm.loss_and_grads('mtl', inputs, p2, 2, ctc_weight=0.3)[0]['loss']
        p2[name][idx] -= 2 * eps
        lm = blstm.loss_and_grads('mtl', inputs, p2, 2, ctc_weight=0.3)[0]['loss']
        assert abs((lp - lm) / (2 * eps) - grads[name][idx]) < 1e-7, name
