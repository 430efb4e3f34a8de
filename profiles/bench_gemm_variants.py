#!/usr/bin/env python
"""A/B of the CTA-pair GEMM's tile shape (AVSI_GEMM_WIDE: 0 = 256 x 256 tiles with a double-buffered accumulator everywhere,
1 = 256 x 512 tiles for the split-K and K >= 2048 shapes, 2 = 256 x 512 wherever N allows) on the shapes of the training
step (B = 2048, T = 250): projection, dX, dW_ih, dW_hh.  CUDA events, L2 flushed between launches.
(Round-2 history: the same script measured an L2 prefetch cursor in the TMA producer, AVSI_GEMM_PF = 4..32 k-blocks ahead:
proj 1.03 -> 1.15-1.26 ms, dW_ih 1.00 -> 1.27-1.62 ms -- removed.)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsi_b200 import _lib
from avsi_b200.blstm import A_IL, C_IL, gemm, pick_split_k

d = torch.device('cuda:0')
M, NG, NY = 2048 * 250, 2048, 512
X = torch.randn(M, NY, device=d).half()
W = (torch.randn(NG, NY, device=d) * 0.05).half()
WT = W.t().contiguous()
G = torch.randn(M, NG, device=d).half()          # used as interleaved storage (values are random anyway)
Gout = torch.empty(M, NG, dtype=torch.float16, device=d)
dY = torch.empty(M, NY, dtype=torch.float16, device=d)
dW = torch.zeros(NG, NY, device=d)
dWhh = torch.zeros(1024, 256, device=d)
Ybuf = torch.randn(M + 4096, NY, device=d).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
p = _lib.ptr
shapes = {
    'proj  [M,512]x[2048,512]^T -> IL f16': (lambda: gemm(p(X), NY, p(W), NY, p(Gout), NG, None, M, NG, NY, 0, 0, layout=C_IL), 2.0 * M * NG * NY),
    'dX    IL[M,2048]x[512,2048]^T -> IL f16': (lambda: gemm(p(G), NG, p(WT), NG, p(dY), NY, None, M, NY, NG, 0, 0, layout=A_IL | C_IL), 2.0 * M * NG * NY),
    'dW_ih IL[M,2048]^T x [M,512] -> f32 splitK': (lambda: gemm(p(G), NG, p(X), NY, p(dW), NY, None, NG, NY, M, 1, 2, pick_split_k(NG, NY, M), layout=A_IL), 2.0 * M * NG * NY),
    'dW_hh IL[M,1024]^T x [M,256] -> f32 splitK': (lambda: gemm(G.data_ptr(), NG, Ybuf.data_ptr(), NY, p(dWhh), 256, None, 1024, 256, M, 1, 2, pick_split_k(1024, 256, M, 148), layout=A_IL), 2.0 * M * 1024 * 256),
}
X0 = torch.randn(M, 448, device=d).half()
W0 = (torch.randn(NG, 448, device=d) * 0.05).half()
shapes['proj0 [M,448]x[2048,448]^T -> IL f16'] = (lambda: gemm(p(X0), 448, p(W0), 448, p(Gout), NG, None, M, NG, 448, 0, 0, layout=C_IL), 2.0 * M * NG * 448)
# variants alternate launch by launch (the box's clocks drift under load: back-to-back blocks per variant are biased)
variants = ((0, 0), (1, 0), (1, 1))
times = {v: {n: [] for n in shapes} for v in variants}
for it in range(13):
    for name, (fn, flops) in shapes.items():
        for v in variants:
            _lib.set_env(AVSI_GEMM_WIDE=v[0], AVSI_GEMM_ASTAT=v[1])
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it >= 1:
                times[v][name].append(e0.elapsed_time(e1))
for v in variants:
    row = {'AVSI_GEMM_WIDE': v[0], 'AVSI_GEMM_ASTAT': v[1]}
    for name, (fn, flops) in shapes.items():
        ts = sorted(times[v][name])
        med = ts[len(ts) // 2]
        row[name.split()[0]] = {'ms_median': round(med, 4), 'ms_min': round(ts[0], 4), 'tflops_median': round(flops / med / 1e9, 1)}
    print(json.dumps(row), flush=True)
_lib.set_env(AVSI_GEMM_WIDE=None, AVSI_GEMM_ASTAT=None)
