import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
out = []
for r in data:
    n = int(r[idx['# Samples']] or 0)
    if n == 0: continue
    st = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:3]
    out.append((n, r[idx['Address']], r[idx['Source']][:70], r[idx['Instructions Executed']], st))
out.sort(reverse=True)
for n, a, s, ie, st in out[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print('%5.1f%% %s %-70s ex=%s %s' % (100.0 * n / tot, a[-5:], s, ie, ' '.join('%s:%d' % (x[6:], c) for c, x in st if c)))
