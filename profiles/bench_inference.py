"""BASELINE configs[0]: A-SI BLSTM inference (`prediction`, is_training=False graph of models.py:106-138) on synthetic
3 s utterances with random-init weights: utterances/s on the B200 (inputs resident; CUDA events; inputs + activations
exceed L2 from B = 64 on) next to the torch-CPU port of the reference graph on the host cores.

    python profiles/bench_inference.py [--batches 8,32,256,2048] [--out gpurun_out/inference.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from avsi_b200 import models, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='8,32,256,2048')
    ap.add_argument('--out', default='gpurun_out/inference.json')
    ap.add_argument('--cpu-batch', type=int, default=32)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    rows = []
    for B in [int(x) for x in a.batches.split(',')]:
        h = synth.make_batch(B, audio_len=48000, seed=0)
        cfg = synth.default_config('a-blstm', batch_size=B, audio_len=48000)
        res = {k: torch.from_numpy(np.ascontiguousarray(h[k])).to(dev) for k in ('wav', 'mask', 'mean', 'std', 'seq_len')}
        model = models.StackedBLSTMModel(res['seq_len'], res['wav'], res['mask'], res['mean'], res['std'], 0.0, cfg,
                                         input='a', is_training=False, device=dev)

        def step():
            model.feed(sequence_lengths=res['seq_len'], target_sources=res['wav'], masks=res['mask'])
            return model.prediction
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        reps = 20 if B <= 256 else 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rows.append(dict(impl='b200', B=B, ms_per_batch=round(ms, 4), utt_per_s=round(B / ms * 1e3, 1)))
        print(json.dumps(rows[-1]), flush=True)
        del model
        torch.cuda.empty_cache()
    from oracle import cpu_port
    for B in sorted({8, a.cpu_batch}):
        h = synth.make_batch(B, audio_len=48000, seed=0)
        ups, cores, dt = cpu_port.time_inference(h, steps=2, warmup=1)
        rows.append(dict(impl='cpu port of the reference graph (torch fp32)', B=B, cores=cores, ms_per_batch=round(dt * 1e3, 1),
                         utt_per_s=round(ups, 2)))
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    with open(a.out, 'w') as f:
        json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
