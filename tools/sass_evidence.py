"""Blackwell-native evidence from the SHIPPED library: per kernel, the SASS mnemonics that prove tcgen05 / TMEM / TMA / packed
fp32x2 (B200_PROFILING.md "What proves a Blackwell-native kernel").  python tools/sass_evidence.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'pvraft_b200', 'libpvraft_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pat = {'UTC*MMA (tcgen05.mma)': r'\bUTC[A-Z]*MMA', 'LDTM (tcgen05.ld)': r'\bLDTM', 'UTMALDG (TMA tensor load)': r'\bUTMALDG',
       'UBLKCP (TMA bulk copy)': r'\bUBLKCP', 'SYNCS (mbarrier)': r'\bSYNCS', 'FFMA2/FADD2/FMUL2 (fp32x2)': r'\bF(FMA|ADD|MUL)2\b',
       'LDGSTS (cp.async)': r'\bLDGSTS', 'HMMA (legacy mma.sync)': r'\bHMMA', 'REDUX/CREDUX (warp reduce)': r'\bC?REDUX',
       'MATCH (warp match)': r'\bMATCH'}
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        cur = cur.replace('void ', '').replace('pvraft::', '')
        per.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    for name, rx in pat.items():
        if re.search(rx, line):
            per[cur][name] += 1
lines = [f'cuobjdump -sass pvraft_b200/libpvraft_b200.so ({os.path.getsize(so)} bytes) -- instruction counts per kernel (template instances summed)', '']
tot = collections.Counter()
for k, c in per.items():
    if c:
        lines.append(f'{k}: ' + ', '.join(f'{n} x{v}' for n, v in c.items()))
        tot.update(c)
lines += ['', 'whole library: ' + ', '.join(f'{n} x{v}' for n, v in tot.items())]
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_evidence.txt')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[-8:]))
