"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`).
python tools/launch_summary.py X.csv [rows]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
    agg[r[ki][:60]][0] += 1
    agg[r[ki][:60]][1] += v
tot = sum(v[1] for v in agg.values())
print('total us', tot, 'launches', sum(v[0] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f'{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% n={v[0]:4d} avg={v[1] / v[0]:8.1f}  {k}')
