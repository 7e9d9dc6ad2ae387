"""Summarise an `ncu --page source --csv` dump: runs of SASS with similar execution counts."""
import csv
import sys

path, pts = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr, data = rows[1], rows[2:]
ia, isrc, isamp, iw = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('L1 Wavefronts Shared')
out = [(int(r[ia]), int(r[isamp]), r[isrc].strip(), int(r[iw]) if r[iw].isdigit() else 0) for r in data]
tot = sum(o[0] for o in out)
print(f'total warp-inst {tot}  per point {tot / pts:.1f}  sass lines {len(out)}  smem wavefronts/pt {sum(o[3] for o in out) / pts:.1f}')
start = 0
for i in range(1, len(out) + 1):
    if i == len(out) or abs(out[i][0] - out[start][0]) > 0.2 * max(out[start][0], 1):
        n = i - start
        c = sum(o[0] for o in out[start:i]); s = sum(o[1] for o in out[start:i]); wv = sum(o[3] for o in out[start:i])
        if c / pts > float(sys.argv[3]) if len(sys.argv) > 3 else 10:
            print(f'[{start:5d},{i:5d}) n={n:4d} exec/pt={c / pts:8.1f} x{c / pts / n:6.2f} samples={s:6d} smemwf/pt={wv / pts:6.1f}  {out[start][2][:60]}')
        start = i
