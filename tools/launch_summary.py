"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
        continue
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:60]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f'total {tot:.1f} us over {sum(c for c, _ in agg.values())} launches')
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f'{t:10.1f} us {100 * t / tot:5.1f}% n={c:4d} avg={t / c:8.1f}  {k}')
