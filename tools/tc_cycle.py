"""Average duration of every k_tc_linear launch position inside the RAFT iteration (9 launches per iteration) from an ncu
launch list (--metrics gpu__time_duration.sum --csv).  python tools/tc_cycle.py <launches.csv> [launches per iteration]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
per = int(sys.argv[2]) if len(sys.argv) > 2 else 9
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
tc = [float(r[mv]) / 1000 for r in rows[h + 1:] if len(r) > mv and 'k_tc_linear' in r[kn]]
loop = tc[-per * 32:]
names = ['vox fc1', 'feature+cc', 'motion', 'GRU zr', 'GRU q', 'P', 'fc2', 'fc3', 'flow head']
for i in range(per):
    v = loop[i::per]
    print(f'{names[i] if per == 9 else i:12s} {sum(v) / len(v):6.1f} us')
print('sum per iteration', round(sum(loop) / 32, 1), ' pre-loop tc launches', len(tc) - len(loop), round(sum(tc[:-per * 32]), 1))
