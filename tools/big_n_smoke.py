"""BASELINE configs[4] support check: full RSF forwards at N = 16384 and 32768 (K = 512, 2 iterations) run and stay finite, with
the peak memory they take.  python tools/big_n_smoke.py"""
import os, sys, types, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pvraft_b200 import RSF
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = RSF(bench.make_args()).to(dev).eval()
for n, b in ((16384, 2), (32768, 1)):
    pc1, pc2 = [t.to(dev) for t in bench.synthetic_clouds(b, n, 1)]
    torch.cuda.reset_peak_memory_stats()
    with torch.no_grad():
        m([pc1, pc2], 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = m([pc1, pc2], 2)
        e1.record()
        torch.cuda.synchronize()
    print(json.dumps({'N': n, 'B': b, 'iters': 2, 'ms': e0.elapsed_time(e1), 'finite': bool(torch.isfinite(out[-1]).all()),
                      'mean_abs_flow': float(out[-1].abs().mean()), 'peak_GB': torch.cuda.max_memory_allocated() / 1e9}))
