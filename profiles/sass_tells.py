#!/usr/bin/env python
"""Occurrences of the Blackwell-specific SASS mnemonics per kernel of libavsi_b200.so (no GPU needed).
usage: python profiles/sass_tells.py > profiles/sass_tells.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'audio-visual-speech-inpainting_b200', 'libavsi_b200.so')
TELLS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'UBLKPF', 'UTCBAR', 'SYNCS', 'HMMA', 'LDGSTS', 'FFMA2', 'MUFU.TANH',
         'FENCE.VIEW.ASYNC', 'MEMBAR', 'ST.ASYNC', 'STAS', 'LDL', 'STL', 'STG', 'LDG']

sass = subprocess.run('cuobjdump -sass %s | c++filt' % LIB, shell=True, capture_output=True, text=True).stdout
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r'Function : (.*)', line)
    if m:
        name = m.group(1).strip()
        name = re.sub(r'^void ', '', name)
        name = re.sub(r'^avsi::', '', name)
        name = re.sub(r'\(.*', '', name)
        cur = name
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    for t in TELLS:
        if op == t or op.startswith(t + '.') or (t == 'HMMA' and op.startswith('HMMA')):
            counts[cur][t] += 1
print('SASS tells of libavsi_b200.so -- `cuobjdump -sass libavsi_b200.so | c++filt`, occurrences per kernel (sm_100a, nvcc 12.9); profiles/sass_tells.py')
print('UTCHMMA = tcgen05.mma kind::f16 | LDTM / STTM = tcgen05.ld / st (TMEM) | UTMALDG / UTMASTG = TMA tensor loads / stores | UBLKCP = cp.async.bulk')
print('(DSMEM pushes of the forward recurrence, bulk stores) | UBLKPF / UTMAPF = bulk / tensor L2 prefetch | UTCBAR = tcgen05.commit | SYNCS = mbarrier ops')
print('HMMA = mma.sync (small-batch recurrence) | LDGSTS = cp.async | FFMA2 = packed fp32x2 | STAS = st.async (DSMEM) | LDL / STL = spills')
print()
for name in sorted(order):
    c = counts[name]
    if not any(c[t] for t in TELLS[:12]) and 'lstm' not in name and 'frontend' not in name:
        continue
    print('%-52s %s' % (name, ' '.join('%s=%d' % (t, c[t]) for t in TELLS if c[t])))
