#!/usr/bin/env python
"""A/B of the tcgen05 recurrence kernels' store paths: plain STG stores vs shared-memory
staging + cp.async.bulk for the forward kernel's gates (AVSI_L4_BULK), and the position of the BPTT kernel's
second-half loads relative to its proxy fence (AVSI_B4_LATE), one layer, CUDA events, same box, same process.  Also checks that
both variants write bit-identical tensors.    usage: python profiles/bench_lstm_variants.py [T] [B1,B2,...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsi_b200 import _lib

lib = _lib.load()
d = torch.device('cuda:0')
T = int(sys.argv[1]) if len(sys.argv) > 1 else 250
BS = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '2048').split(',')]
out = []
for B in BS:
    Mp = -(-T * B // 32) * 32
    g0 = torch.randn(Mp, 2048, device=d).half()
    whh = (torch.randn(2048, 256, device=d) * 0.05).half()
    whhT = whh.t().contiguous()
    bias = torch.zeros(2048, device=d)
    dy = torch.randn(Mp, 512, device=d).half()
    scratch = torch.empty(int(lib.avsi_lstm_bwd_scratch_bytes(B)) // 4 + 4, device=d)
    ref = {}
    if os.environ.get('AVSI_VARIANT_SET', '') == 'r02a':
        variants = [('stg', dict(AVSI_L4_BULK=0, AVSI_B4_LATE=0)), ('stg+late', dict(AVSI_L4_BULK=0, AVSI_B4_LATE=1)),
                    ('bulk+late', dict(AVSI_L4_BULK=1, AVSI_B4_LATE=1))]
    elif os.environ.get('AVSI_VARIANT_SET', '') == 'r02h':   # (the late-store variants were slower and are gone from the code)
        base = dict(AVSI_L4_BULK=0, AVSI_B4_LATE=0, AVSI_L4_CFENCE=0, AVSI_B4_CFENCE=0)
        variants = [('writer-fence', dict(base)), ('consumer-fence', dict(base, AVSI_L4_CFENCE=1, AVSI_B4_CFENCE=1))]
    elif os.environ.get('AVSI_VARIANT_SET', '') == 'r02i':   # (issuing the first K-half under the second pass was 13 % slower and is gone)
        off = dict(AVSI_L4_BULK=0, AVSI_B4_LATE=0, AVSI_L4_CFENCE=0, AVSI_B4_CFENCE=0, AVSI_L4_BPF=0, AVSI_B4_BPF=0)
        cf = dict(off, AVSI_L4_CFENCE=1, AVSI_B4_CFENCE=1)
        variants = [('writer-fence', off), ('consumer-fence', cf), ('consumer-fence+bulk-prefetch', dict(cf, AVSI_L4_BPF=1, AVSI_B4_BPF=1))]
    elif os.environ.get('AVSI_VARIANT_SET', '') == 'r02j':   # distance of the control thread's bulk L2 prefetch
        cf = dict(AVSI_L4_BULK=0, AVSI_B4_LATE=0, AVSI_L4_CFENCE=1, AVSI_B4_CFENCE=1, AVSI_L4_BPF=0, AVSI_B4_BPF=0)
        variants = [('per-thread-prefetch', cf)] + [('bulk-prefetch-%d' % k, dict(cf, AVSI_L4_BPF=1, AVSI_B4_BPF=1, AVSI_L4_PREFETCH=k, AVSI_B4_PFD=k))
                                                     for k in (1, 2, 3, 4)]
    elif os.environ.get('AVSI_VARIANT_SET', '') == 'r02k':   # BPTT dG through TMA tensor stores of the control thread (out of the A-half) vs STG.128 per thread
        # (the same for the forward kernel's activated gates -- 32 KB staging per pass + 4 TMA stores, AVSI_L4_STMA -- was 16 %
        # slower, r02l_*: the staging is refilled one pass later, before the store queued behind the DSMEM pushes has read it)
        variants = [('stg', dict(AVSI_B4_STMA=0)), ('tma-store', dict(AVSI_B4_STMA=1))]
    elif os.environ.get('AVSI_VARIANT_SET', '') == 'r02n':
        # BPTT reduce-scatter by st.async straight out of TMEM (no staging / control-thread hop); second chain as two N = 128
        # halves (AVSI_B4_NSPLIT, gone from the code: 1.382 vs 1.323 ms) -- logs r02n_*
        variants = [('default', dict())]
    # (r02y: forward activations as tanh.approx.f16x2 -- SASS shows two MUFU.TANH.F16 per instruction, no packed MUFU: 1.161 vs
    # 1.142 ms; with every MUFU compiled out of the cell update (wrong values) the launch still takes 1.121 ms: the XU pipe is
    # not on the step's critical path.  Log r02y_*, code not kept)
    # (r02y: forward pre-activations through TMA tensor loads + shared memory instead of LDG.128: +4 / +6 %, log kept, code not)
    # (r02u: the forward TMA store again with the first pass's store AHEAD of its pushes and the second BEHIND them: +7 %, log kept)
    else:   # current defaults against the round-1 forms that are still selectable
        on = dict(AVSI_L4_CFENCE=1, AVSI_B4_CFENCE=1, AVSI_L4_BPF=1, AVSI_B4_BPF=1, AVSI_B4_STMA=1)
        variants = [('default', on), ('writer-fence,stg,per-thread-prefetch', dict(on, AVSI_L4_CFENCE=0, AVSI_B4_CFENCE=0))]
    for name, env in variants + variants:
        _lib.set_env(AVSI_LSTM_FWD='l4', AVSI_LSTM_BWD='l4', **env)
        gates = g0.clone()
        y = torch.zeros(T * B, 512, dtype=torch.float16, device=d)
        cst = torch.zeros(Mp, 512, device=d)
        dbias = torch.zeros(2048, device=d)
        res = {'fwd': [], 'bwd': []}
        for it in range(6):
            gates.copy_(g0)
            dbias.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            _lib.check(lib.avsi_lstm_fwd(_lib.ptr(gates), _lib.ptr(whh), _lib.ptr(bias), _lib.ptr(y), _lib.ptr(cst), T, B, 0,
                                         _lib.stream_ptr()))
            e[1].record()
            if it == 5:
                torch.cuda.synchronize()
                act = gates.clone()
            _lib.check(lib.avsi_lstm_bwd(_lib.ptr(gates), _lib.ptr(whhT), _lib.ptr(cst), _lib.ptr(dy), _lib.ptr(dbias),
                                         _lib.ptr(scratch), T, B, _lib.stream_ptr()))
            e[2].record()
            torch.cuda.synchronize()
            if 2 <= it < 5:
                res['fwd'].append(e[0].elapsed_time(e[1]))
                res['bwd'].append(e[1].elapsed_time(e[2]))
        cur = dict(act=act, y=y.clone(), c=cst.clone(), dg=gates.clone(), db=dbias.clone())
        same = None
        if ref:
            same = {k: bool(torch.equal(cur[k], ref[k])) for k in ('act', 'y', 'c', 'dg')}
            same['db_close'] = bool(torch.allclose(cur['db'], ref['db'], rtol=1e-3, atol=1e-2))
        else:
            ref = cur
        row = dict(B=B, T=T, variant=name, fwd_ms=min(res['fwd']), bwd_ms=min(res['bwd']),
                   fwd_us_step=1e3 * min(res['fwd']) / T, bwd_us_step=1e3 * min(res['bwd']) / T, same_as_first=same)
        out.append(row)
        print(json.dumps(row), flush=True)
_lib.set_env(AVSI_LSTM_FWD=None, AVSI_LSTM_BWD=None, AVSI_L4_BULK=None, AVSI_B4_LATE=None, AVSI_L4_CFENCE=None,
             AVSI_B4_CFENCE=None, AVSI_L4_BPF=None, AVSI_B4_BPF=None, AVSI_L4_PREFETCH=None, AVSI_B4_PFD=None, AVSI_B4_STMA=None,
             AVSI_LSTM_ACT=None)
