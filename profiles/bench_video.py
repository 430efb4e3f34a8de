"""Landmark-feature kernel alone (CUDA events, L2 flushed between launches) next to a device copy of the same size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsi_b200 import av_sync, synth
d = torch.device('cuda:0')
B, T = 2048, 250
h = synth.make_batch(B, audio_len=48000, seed=0)
lm = torch.from_numpy(h['landmarks']).to(d).float()
vm, vs = torch.from_numpy(h['vmean']).to(d), torch.from_numpy(h['vstd']).to(d)
flush = torch.zeros(64 << 20, device=d)
ref = torch.empty(B, T, 136, device=d)
def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
out = av_sync.video_pipeline(lm, T, vm, vs, device=d)
print('video_features %.3f ms   fill of the same output %.3f ms   copy %.3f ms  (%.0f MB out)' % (
    timed(lambda: av_sync.video_pipeline(lm, T, vm, vs, device=d)), timed(lambda: ref.fill_(1.0)),
    timed(lambda: ref.copy_(out)), out.numel() * 4 / 1e6))
