#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key throughput metrics, top warp-stall
reasons, and the hottest SASS/source lines.   usage: python profiles/ncu_summary.py file.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

KEYS = [
    'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
    'launch__cluster_size', 'launch__cluster_max_active', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
]


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print('=' * 100)
        for k in KEYS:
            if k in d:
                print('%-75s %s %s' % (k, d[k][0], d[k][1]))
        stalls = []
        for h in d:
            if 'warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    stalls.append((float(d[h][0]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        print('-- top stall reasons (warps stalled per issue-active cycle)')
        for v, h in sorted(stalls, reverse=True)[:8]:
            print('   %8.3f  %s' % (v, h))
    src = run(['-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'])
    rows = list(csv.reader(io.StringIO(src)))
    while rows and not (rows[0] and rows[0][0] in ('Address', '#', 'Line') or (len(rows[0]) > 3 and 'Source' in rows[0])):
        rows.pop(0)
    if len(rows) > 2:
        hdr = rows[0]
        try:
            i_samp = hdr.index('# Samples') if '# Samples' in hdr else [i for i, h in enumerate(hdr) if 'Samples' in h][0]
        except IndexError:
            return
        i_src = hdr.index('Source') if 'Source' in hdr else 1
        body = []
        for r in rows[1:]:
            try:
                body.append((float(r[i_samp]), r[i_src], r))
            except (ValueError, IndexError):
                pass
        tot = sum(b[0] for b in body) or 1.0
        print('-- hottest lines (%% of %d samples)' % tot)
        for s, text, r in sorted(body, key=lambda b: -b[0])[:nlines]:
            print('   %5.1f%%  %s' % (100 * s / tot, text[:150]))


if __name__ == '__main__':
    main()
