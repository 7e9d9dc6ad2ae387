"""Run ONE RSF.forward of the bench workload inside a cudaProfilerStart/Stop range (for ncu
--profile-from-start off).  python tools/profile_forward.py [--batch 8] [--iters 32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvraft_b200 import RSF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--iters', type=int, default=32)
ap.add_argument('--warm', type=int, default=2)
a = ap.parse_args()
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = RSF(bench.make_args()).to(dev).eval()
model.use_cuda_graph = False   # per-launch profiling needs the eager launch sequence
pc1, pc2 = [t.to(dev) for t in bench.synthetic_clouds(a.batch, bench.N_POINTS, 1234)]
with torch.no_grad():
    for _ in range(a.warm):
        model([pc1, pc2], a.iters)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model([pc1, pc2], a.iters)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('done')
