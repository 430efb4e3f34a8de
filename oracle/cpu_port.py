"""Oracle-side CPU port used ONLY as the timed CPU baseline of bench.py (test infrastructure,
never on the product path).

The reference's own implementation of this path is TensorFlow-1.x graph code (audio_processing.py,
models.py) which cannot be installed here (no wheels for Python 3.12, no network), and its training
branch (CudnnLSTM, models.py:95-104) is GPU-only anyway.  This port restates the same AV-SI
training step in float32 on the host cores with torch-CPU tensor ops (multi-threaded through
torch's intra-op pool): framing + Hann + rFFT-512 + log + normalise + mask (models.py:30-45),
3-layer BLSTM with the CudnnCompatibleLSTMCell equations (models.py:106-115), linear head, L1 loss
(models.py:151), autograd backward, TF-form Adam.  ``kind`` = "port" in the bench JSON.
"""
import math
import os
import time

import torch


def make_params(in_dim, hidden=250, n_layers=3, out_dim=257, seed=1):
    g = torch.Generator().manual_seed(seed)
    params = []
    for l in range(n_layers):
        i_l = in_dim if l == 0 else 2 * hidden
        for _ in range(2):
            lim = math.sqrt(6.0 / (i_l + hidden + 4 * hidden))
            params.append(((torch.rand(i_l + hidden, 4 * hidden, generator=g) * 2 - 1) * lim).requires_grad_())
            params.append(torch.zeros(4 * hidden).requires_grad_())
    params.append((torch.randn(2 * hidden, out_dim, generator=g) / math.sqrt(2 * hidden)).requires_grad_())
    params.append(torch.zeros(out_dim).requires_grad_())
    return params


def frontend(wav, mask, mean, std, video, frame_len=384, hop=192, nfft=512):
    B, N = wav.shape
    T = -(-N // hop)
    need = frame_len + hop * (T - 1)
    x = torch.nn.functional.pad(wav, (0, max(0, need - N)))
    frames = x.unfold(1, frame_len, hop)[:, :T]
    win = torch.hann_window(frame_len, periodic=True)
    spec = torch.fft.rfft(frames * win, n=nfft).abs()
    tsn = (torch.log(spec + 1e-6) - mean) / std
    net_in = tsn * mask
    if video is not None:
        net_in = torch.cat([net_in, video], dim=2)
    return tsn, net_in


def _direction(x, kernel, bias, reverse):
    B, T, I = x.shape
    H = kernel.shape[1] // 4
    xw = (x.reshape(B * T, I) @ kernel[:I]).reshape(B, T, 4 * H) + bias     # all-timestep projection
    wh = kernel[I:]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        z = xw[:, t] + h @ wh
        i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(j)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def train_step(params, adam_state, wav, mask, mean, std, video, step, n_layers=3, lr=1e-3):
    tsn, x = frontend(wav, mask, mean, std, video)
    B, T, F = tsn.shape
    for l in range(n_layers):
        kf, bf, kb, bb = params[4 * l:4 * l + 4]
        x = torch.cat([_direction(x, kf, bf, False), _direction(x, kb, bb, True)], dim=2)
    pred = (x.reshape(B * T, -1) @ params[-2] + params[-1]).reshape(B, T, F)
    loss = (tsn - pred).abs().mean()
    grads = torch.autograd.grad(loss, params)
    lr_t = lr * math.sqrt(1.0 - 0.999 ** step) / (1.0 - 0.9 ** step)
    with torch.no_grad():
        for p, g, (m, v) in zip(params, grads, adam_state):
            m.mul_(0.9).add_(g, alpha=0.1)
            v.mul_(0.999).addcmul_(g, g, value=0.001)
            p.sub_(lr_t * m / (v.sqrt() + 1e-8))
    return float(loss.detach())


def time_train_steps(batch, steps=1, warmup=0, threads=None):
    """batch: dict of numpy arrays (wav, mask, mean, std, video [B,T,V]).  Returns (utt/s, cores, seconds/step)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    wav, mask = torch.from_numpy(batch['wav']), torch.from_numpy(batch['mask'])
    mean, std = torch.from_numpy(batch['mean']), torch.from_numpy(batch['std'])
    video = torch.from_numpy(batch['video'])
    params = make_params(mask.shape[2] + video.shape[2])
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    for w in range(warmup):
        train_step(params, state, wav, mask, mean, std, video, w + 1)
    t0 = time.perf_counter()
    for s in range(steps):
        train_step(params, state, wav, mask, mean, std, video, warmup + s + 1)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return wav.shape[0] / dt, threads, dt


def time_inference(batch, steps=3, warmup=1, threads=None):
    """BASELINE configs[0]: A-SI inference (`prediction` of the is_training=False graph, models.py:106-138) on the host
    cores.  batch: numpy wav, mask, mean, std, seq_len.  Returns (utt/s, cores, seconds/step)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    wav, mask = torch.from_numpy(batch['wav']), torch.from_numpy(batch['mask'])
    mean, std = torch.from_numpy(batch['mean']), torch.from_numpy(batch['std'])
    seq = torch.from_numpy(batch['seq_len'].astype('int64'))
    params = make_params(mask.shape[2])

    def run():
        with torch.no_grad():
            tsn, x = frontend(wav, mask, mean, std, None)
            B, T, F = tsn.shape
            for l in range(3):
                kf, bf, kb, bb = params[4 * l:4 * l + 4]
                x = torch.cat([_direction(x, kf, bf, False), _direction(x, kb, bb, True)], dim=2)
            pred = (x.reshape(B * T, -1) @ params[-2] + params[-1]).reshape(B, T, F)
            return pred * (torch.arange(T)[None, :] < seq[:, None]).float()[:, :, None]
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return wav.shape[0] / dt, threads, dt
