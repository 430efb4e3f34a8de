#!/usr/bin/env python
"""Micro-benchmark of the recurrence kernels alone: us per time step vs batch size (CUDA events).
usage: python profiles/bench_lstm.py [T] [B1,B2,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsi_b200 import _lib

lib = _lib.load()
d = torch.device('cuda:0')
T = int(sys.argv[1]) if len(sys.argv) > 1 else 250
Y_IL = int(os.environ.get('AVSI_Y_IL', '0'))          # 1: interleaved layer outputs (B % 32 == 0)
BS = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '16,64,112,128,224,448,512,896').split(',')]
for B in BS:
    g0 = torch.randn(-(-T * B // 32) * 32, 2048, device=d).half()
    gates = g0.clone()
    whh = (torch.randn(2048, 256, device=d) * 0.05).half()
    whhT = whh.t().contiguous()
    bias = torch.zeros(2048, device=d)
    y = torch.empty(T * B, 512, dtype=torch.float16, device=d)
    cst = torch.empty(-(-T * B // 32) * 32, 512, device=d)
    dy = torch.randn(-(-T * B // 32) * 32, 512, device=d).half()          # interleaved dL/dy (values are random anyway)
    dbias = torch.zeros(2048, device=d)
    scratch = torch.empty(int(lib.avsi_lstm_bwd_scratch_bytes(B)) // 4 + 4, device=d)

    def fwd():
        _lib.check(lib.avsi_lstm_fwd(_lib.ptr(gates), _lib.ptr(whh), _lib.ptr(bias), _lib.ptr(y), _lib.ptr(cst), T, B,
                                     Y_IL if B % 32 == 0 else 0, _lib.stream_ptr()))

    def bwd():
        _lib.check(lib.avsi_lstm_bwd(_lib.ptr(gates), _lib.ptr(whhT), _lib.ptr(cst), _lib.ptr(dy), _lib.ptr(dbias),
                                     _lib.ptr(scratch), T, B, _lib.stream_ptr()))

    res = {'fwd': 0.0, 'bwd': 0.0}
    reps = 3
    for it in range(reps + 2):
        gates.copy_(g0)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        fwd()
        e[1].record()
        bwd()
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            res['fwd'] += e[0].elapsed_time(e[1]) / reps
            res['bwd'] += e[1].elapsed_time(e[2]) / reps
    print('B=%4d T=%d  fwd %.3f ms (%.2f us/step)  bwd %.3f ms (%.2f us/step)  -> %.0f utt/s for one layer fwd+bwd'
          % (B, T, res['fwd'], 1e3 * res['fwd'] / T, res['bwd'], 1e3 * res['bwd'] / T,
             B / ((res['fwd'] + res['bwd']) * 1e-3)))

if os.environ.get('AVSI_L4_TIMING'):
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    torch.cuda.synchronize()
    gates.copy_(g0)
    fwd()
    torch.cuda.synchronize()
    lib.avsi_debug_lstm4_timing.argtypes = [ctypes.c_void_p]
    lib.avsi_debug_lstm4_timing(buf)
    for o, who, names in ((0, 'control', ['wait_staged0', 'wait_staged1(+push0)', 'wait_hfull0(+push1)', 'mma0', 'wait_hfull1', 'mma1']),
                          (8, 'compute0', ['prefetch', 'wait_done', 'pass0', 'pass1', 'ystore'])):
        print(who, '  '.join('%s %.0f' % (n, buf[o + i] / T) for i, n in enumerate(names)),
              ' total %.0f cyc/step' % (sum(buf[o:o + 8]) / T))


if os.environ.get('AVSI_B4_TIMING'):
    import ctypes
    buf = (ctypes.c_ulonglong * 24)()
    torch.cuda.synchronize()
    gates.copy_(g0)
    fwd()
    bwd()
    torch.cuda.synchronize()
    lib.avsi_debug_lstm4_bwd_timing.argtypes = [ctypes.c_void_p]
    lib.avsi_debug_lstm4_bwd_timing(buf)
    for o, who, names in ((0, 'control', ['wait_slotfull', 'wait_stagedA0', 'mma0', 'wait_stagedA1', 'mma1', 'send_consumed']),
                          (12, 'compute0', ['wait_slotfull+own', 'pass0', 'write A0', 'pass1', 'wait+write A1', '-', '-',
                                            'wait_done0', 'extract+st.async', 'issue loads'])):
        print(who, '  '.join('%s %.0f' % (n, buf[o + i] / T) for i, n in enumerate(names)),
              ' total %.0f cyc/step' % (sum(buf[o:o + 12]) / T))
