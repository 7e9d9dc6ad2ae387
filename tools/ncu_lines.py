"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py <rep> <points> [min_exec_per_point]   (run where ncu is installed; no GPU needed)"""
import csv, io, subprocess, sys


def _f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0

rep, pts = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
cur = None
agg = {}
fname = ''
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        ie, isamp, iwf, iwfi = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != '':   # a source line (its numbers are the sum of its SASS)
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [r[1].strip(), 0.0, 0.0, 0.0, 0.0])
        a[1] += _f(r[ie]); a[2] += _f(r[isamp]); a[3] += _f(r[iwf]); a[4] += _f(r[iwfi])
tot = sum(a[1] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f'total warp-inst/pt {tot / pts:.1f}  samples {tots:.0f}')
for (f, ln), a in sorted(agg.items()):
    if a[1] / pts >= thr or a[2] / max(tots, 1) > 0.01:
        print(f'{f}:{ln:4d} exec/pt={a[1] / pts:7.1f} samp={100 * a[2] / max(tots, 1):5.1f}% wf/pt={a[3] / pts:6.1f} ideal={a[4] / pts:6.1f} | {a[0][:110]}')
