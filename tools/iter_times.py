"""Print the per-kernel times of one RAFT iteration from an ncu launch list."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
seq = []
for r in csv.DictReader(lines):
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
        continue
    u = r['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else v
    seq.append((re.sub(r'\(.*', '', r['Kernel Name']).replace('pvraft::', '').replace('void ', '')[:24], v))
idx = [i for i, s in enumerate(seq) if 'corr_lookup' in s[0]]
i0, i1 = idx[3], idx[4]
print('iteration: ' + ', '.join(f'{n}={t:.0f}' for n, t in seq[i0:i1] if t > 5), ' sum=%.0f' % sum(t for _, t in seq[i0:i1]))
print('total %.1f ms' % (sum(t for _, t in seq) / 1e3))
