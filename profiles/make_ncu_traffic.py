#!/usr/bin/env python
"""Build profiles/ncu_traffic.json from the `ncu --set full` captures of profiles/run_ncu.sh <tag> (read here, no GPU
needed): dram__bytes_read.sum + dram__bytes_write.sum per launch for the kernels bench.py puts a roofline on, stamped
with the digest of each kernel's sources (bench.KERNEL_SOURCES) so that bench.py drops a figure whose kernel has changed
since.    usage: python profiles/make_ncu_traffic.py <tag> [batch] [audio_len]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def launches(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        rd, wr = d['dram__bytes_read.sum'], d['dram__bytes_write.sum']
        byts = float(rd[0].replace(',', '')) * UNIT[rd[1]] + float(wr[0].replace(',', '')) * UNIT[wr[1]]
        dur = d['gpu__time_duration.sum']
        res.append({'kernel': d['Kernel Name'][0], 'dram_bytes': byts, 'duration': dur[0] + ' ' + dur[1]})
    return res


def main():
    tag = sys.argv[1]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    audio_len = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
    g = lambda k: os.path.join(ROOT, 'gpurun_out', 'prof_%s_%s.ncu-rep' % (k, tag))
    out = {'comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures %s '
                      '(profiles/run_ncu.sh %s %d: bench.py --batch %d); bench.py copies an entry into roofline.traffic when it '
                      'runs that workload and the sources of that kernel (bench.KERNEL_SOURCES) still hash to '
                      'source_digests[entry]' % (tag, tag, batch, batch),
           'workload': {'batch': batch, 'audio_len': audio_len}, 'captures': '%s_*_summary.txt' % tag, 'source_digests': {}}
    for name, kern in (('lstm_bwd', 'lstm4_bwd_kernel'), ('lstm_fwd', 'lstm4_fwd_kernel'), ('frontend', 'frontend_train_kernel'),
                       ('gemm_proj_fwd', 'gemm_f16_2sm_astat_kernel')):
        if os.path.exists(g(kern)):
            out[name] = int(launches(g(kern))[0]['dram_bytes'])
            out['source_digests'][name] = bench.kernel_source_digest(name)
    # the 13 CTA-pair GEMM launches of one backward pass, in issue order (blstm.py backward): head dW, head dX, then per
    # layer dW_ih, dW_hh forward direction, dW_hh backward direction, dX (no dX below layer 0)
    if os.path.exists(g('gemm_f16_2sm_kernel')):
        ls = launches(g('gemm_f16_2sm_kernel'))
        if len(ls) == 13:
            spans = {'gemm_dw_head': [0], 'gemm_dw_ih': [2, 6, 10], 'gemm_dw_hh': [3, 4, 7, 8, 11, 12], 'gemm_dx': [1, 5, 9]}
            which = {}
            for name, idx in spans.items():
                out[name] = int(sum(ls[i]['dram_bytes'] for i in idx) / len(idx))
                out['source_digests'][name] = bench.kernel_source_digest('gemm_dx')
                for i in idx:
                    which[i] = name
            out['gemm_backward_launches'] = [{'i': i, 'span': which[i], 'dram_bytes': int(l['dram_bytes']), 'duration': l['duration']}
                                             for i, l in enumerate(ls)]
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'), 'w'), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ('comment', 'gemm_backward_launches')}, indent=1))


if __name__ == '__main__':
    main()
