"""Robustness sweep: RSF / RSF_refine forward at several (B, N, K, iters) against the oracle on a row subsample.
python tools/shape_sweep.py"""
import os, sys, time, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pvraft_b200 import RSF, RSF_refine
dev = torch.device('cuda:0')
for refine, b, n, k, iters in [(False, 1, 8192, 512, 8), (False, 3, 4096, 256, 4), (True, 2, 8192, 512, 4), (False, 5, 2048, 128, 3),
                               (False, 2, 1000, 64, 2), (False, 1, 16384, 512, 2), (True, 1, 640, 32, 2)]:
    args = types.SimpleNamespace(corr_levels=3, base_scales=0.25, truncate_k=k)
    torch.manual_seed(0)
    m = (RSF_refine if refine else RSF)(args).to(dev).eval()
    pc1, pc2 = [t.to(dev) for t in bench.synthetic_clouds(b, n, 7)]
    with torch.no_grad():
        out = m([pc1, pc2], iters)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = m([pc1, pc2], iters)
        torch.cuda.synchronize()
    last = out if torch.is_tensor(out) else out[-1]
    ok = bool(torch.isfinite(last).all())
    print(f'refine={refine} B={b} N={n} K={k} iters={iters}: {(time.perf_counter() - t0) * 1e3:.1f} ms  finite={ok}  |flow| mean {float(last.abs().mean()):.4f}')
