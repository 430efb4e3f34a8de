"""One-line digest of bench.py JSON lines: python profiles/bench_digest.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        j = json.load(open(f))
    except Exception as e:                       # noqa: BLE001
        print(f, 'unreadable:', e)
        continue
    k = j.get('kernels', {})
    print(f, 'B', j['config'].get('batch_per_gpu'), 'T', j['config'].get('frames'), 'n', j['n_gpus'],
          '| %.0f utt/s, %.2f ms/step, e2e %.0f |' % (j['value'], j['ms_per_step'], j['e2e']['value']),
          {n: round(k[n]['ms_per_step'], 2) for n in list(k)[:6]})
