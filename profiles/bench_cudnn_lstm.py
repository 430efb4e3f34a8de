#!/usr/bin/env python
"""The BLSTM stack of the AV-SI train step (3 x BLSTM-250 + head + L1 loss, forward + backward, no front end, no optimiser)
on this repo's kernels and on cuDNN's LSTM through torch.nn.LSTM -- the library the reference's CudnnLSTM (models.py:95-104)
would run on the same GPU -- at the bench shape.  A GPU-side baseline next to the CPU one; cuDNN is only ever the thing
compared against.    usage: python profiles/bench_cudnn_lstm.py [B] [audio_len]    (prints one JSON line per arm)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch


def timed(fn, reps=4, warm=2):
    ts = []
    for i in range(warm + reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def main():
    from avsi_b200 import av_sync, models, synth
    from avsi_b200.layout import init_canonical
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    audio_len = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
    batch = synth.make_batch(B, audio_len=audio_len, seed=0)
    cfg = synth.default_config('av-blstm', batch_size=B, audio_len=audio_len)
    cls, inp = models.MODEL_REGISTRY['av-blstm']
    video = av_sync.video_pipeline(batch['landmarks'], batch['T'], batch['vmean'], batch['vstd'])
    model = cls(batch['seq_len'], batch['wav'], batch['mask'], batch['mean'], batch['std'], 0.0, cfg, video_features=video, input=inp)
    model.assign_vars(init_canonical(model.engine.layout, seed=1, bias_scale=0.05))
    model.compute_gradients()                                  # front end runs once and stays cached for this feed
    med, best = timed(lambda: model.compute_gradients())
    T = batch['T']
    rows = [{'arm': 'this repo (fp16 operands, fp32 accumulate / state): projections + recurrence + head + L1 + BPTT + dX + dW',
             'B': B, 'T': T, 'ms_median': med, 'ms_min': best, 'utterances_per_s': B / (med * 1e-3)}]
    print(json.dumps(rows[-1]), flush=True)
    x32 = model.net_inputs.float().contiguous()
    target = model.target_spec_norm.float()
    dev = x32.device
    I = x32.shape[2]
    del model
    torch.cuda.empty_cache()
    for name, dtype, tf32 in (('cuDNN fp32 (TF32 off): what the reference graph asks for', torch.float32, False),
                              ('cuDNN fp32 with TF32 tensor cores allowed', torch.float32, True),
                              ('cuDNN fp16', torch.float16, False)):
        lstm = torch.nn.LSTM(I, 250, num_layers=3, batch_first=True, bidirectional=True).to(dev).to(dtype)
        head = torch.nn.Linear(500, 257).to(dev).to(dtype)
        x = x32.to(dtype)
        tg = target.to(dtype)

        def step():
            for p in list(lstm.parameters()) + list(head.parameters()):
                p.grad = None
            y, _ = lstm(x)
            loss = (tg - head(y)).abs().float().mean()
            loss.backward()
        try:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=tf32):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                med, best = timed(step)
            rows.append({'arm': name, 'B': B, 'T': T, 'ms_median': med, 'ms_min': best, 'utterances_per_s': B / (med * 1e-3),
                         'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9})
        except Exception as exc:                                  # e.g. out of memory at this batch
            rows.append({'arm': name, 'B': B, 'T': T, 'error': repr(exc)[:200]})
        print(json.dumps(rows[-1]), flush=True)
        del lstm, head, x, tg
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False


if __name__ == '__main__':
    main()
