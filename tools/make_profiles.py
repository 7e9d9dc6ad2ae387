"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.
python tools/make_profiles.py <round-tag> <launch-list.csv> <lookup.ncu-rep> [<other.ncu-rep> ...]"""
import csv
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out = os.path.join(ROOT, 'profiles')
os.makedirs(out, exist_ok=True)
shutil.copy(launches, os.path.join(out, f'{tag}_launches.csv'))
txt = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'launch_summary.py'), launches, '30'],
                     capture_output=True, text=True).stdout
open(os.path.join(out, f'{tag}_launches_summary.txt'), 'w').write(
    'ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/profile_forward.py\n'
    '(one RSF.forward of the bench workload: B=8, N=8192, K=512, iters=32; per-launch times are cold-cache and\n'
    'serialised: compare SHARES)\n\n' + txt)
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_active.max',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for rep in reps:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    h = rows[0]
    name = os.path.splitext(os.path.basename(rep))[0]
    lines = [f'ncu --set full --clock-control none --import-source on  ->  {os.path.basename(rep)}', '']
    for r in rows[2:]:
        lines.append('kernel: ' + r[h.index('Kernel Name')])
        for k in h:
            if k in KEYS or ('issue_stalled' in k and k.endswith('per_issue_active.ratio')):
                v = r[h.index(k)]
                if v not in ('0', '0.0', ''):
                    lines.append(f'  {k:90s} {v} {rows[1][h.index(k)]}')
        lines.append('')
        if 'k_corr_lookup' in r[h.index('Kernel Name')]:
            rd = float(r[h.index('dram__bytes_read.sum')]); wr = float(r[h.index('dram__bytes_write.sum')])
            unit = rows[1][h.index('dram__bytes_read.sum')]
            mul = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(unit, 1)
            sha = hashlib.sha256(open(os.path.join(ROOT, 'pvraft_b200', 'csrc', 'corr_lookup.cu'), 'rb').read()).hexdigest()
            json.dump({'kernel': 'k_corr_lookup', 'source': os.path.basename(rep), 'workload': 'B=8, N=8192, K=512 (one launch)',
                       'source_sha256': sha,
                       'dram_bytes_read': rd * mul, 'dram_bytes_write': wr * mul, 'dram_bytes_per_launch': (rd + wr) * mul},
                      open(os.path.join(out, 'lookup_dram_bytes.json'), 'w'), indent=1)
    open(os.path.join(out, f'{tag}_{name}_ncu_summary.txt'), 'w').write('\n'.join(lines))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    tmp = os.path.join(out, f'.{name}_source.csv')
    open(tmp, 'w').write(src)
    if 'lookup' in name:
        s = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), rep, '65536', '9'],
                           capture_output=True, text=True).stdout
        open(os.path.join(out, f'{tag}_{name}_source_hotspots.txt'), 'w').write(
            'per source line: warp-instructions per point (exec/pt, B*N = 65536 points), share of stall samples, shared-memory wavefronts\n' + s)
    os.remove(tmp)
print(os.listdir(out))
