"""Micro-benchmark of the fused correlation-lookup kernel alone (HBM roofline of SURVEY.md 8d).
python tools/bench_lookup.py [--batch 8] [--points 8192] [--k 512] [--box 10 3] [--reps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvraft_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--points', type=int, default=8192)
ap.add_argument('--k', type=int, default=512)
ap.add_argument('--box', type=float, nargs='+', default=[10.0, 3.0])
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--profile', action='store_true')
ap.add_argument('--sweep', action='store_true', help='BASELINE config 5: N in {2048..32768}, B sized so the state is >= 4x L2')
a = ap.parse_args()
dev = torch.device('cuda:0')
peaks, kind = bench.measured_peaks()
cases = [(a.batch, a.points, a.k, box) for box in a.box]
if a.sweep:
    cases = [(64, 2048, 512, 10.0), (32, 4096, 512, 10.0), (16, 8192, 512, 10.0), (8, 16384, 512, 10.0), (4, 32768, 512, 10.0)]
for b, n, k, box in cases:
    g = torch.Generator(device=dev).manual_seed(1)
    xyz2 = box * torch.rand(b, n, 3, device=dev, generator=g)
    start = torch.randint(0, n, (b, n, 1), device=dev, generator=g)
    step = torch.randint(0, n // 2, (b, n, 1), device=dev, generator=g) * 2 + 1
    idx = ((start + torch.arange(k, device=dev).view(1, 1, k) * step) % n).to(torch.int32).contiguous()
    coords = xyz2[:, torch.randperm(n, device=dev, generator=g)] + 0.2 * (torch.rand(b, n, 3, device=dev, generator=g) * 2 - 1)
    corr = torch.sort(torch.randn(b, n, k, device=dev, generator=g) * 5 + 20, dim=2, descending=True).values.contiguous()
    tab = ops.xyz_pad(xyz2.contiguous())
    corr, idx = ops.corr_reorder(corr, idx)
    coords = coords.contiguous()
    out = ops.corr_lookup(corr, idx, tab, coords, 3, 0.25)
    torch.cuda.synchronize()
    if a.profile:
        torch.cuda.profiler.start()
    evs = []
    for _ in range(a.reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out['moments'].zero_()   # also holds the kernel's per-sample work counter
        s.record()
        ops.corr_lookup(corr, idx, tab, coords, 3, 0.25, vox=out['vox'], knn_sel=out['knn_sel'], moments=out['moments'])
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    if a.profile:
        torch.cuda.profiler.stop()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    med = ms[len(ms) // 2]
    alg = bench.alg_bytes_lookup(n, k) * b
    valid = float((out['vox'][..., :81].reshape(b, n, 3, 27) != 0).float().sum(-1).mean())
    print(json.dumps({'box': box, 'B': b, 'N': n, 'K': k, 'median_ms': med, 'min_ms': ms[0], 'us_per_sample': med * 1e3 / b,
                      'alg_GBps': alg / med / 1e6, 'frac_of_hbm_peak': alg / med / 1e6 / peaks['hbm_gbs'], 'peak': kind,
                      'nonempty_cells_per_level': valid}))
