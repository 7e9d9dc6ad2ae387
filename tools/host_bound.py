"""Is the forward pass limited by the host launch path or by the GPU?  Prints the time the python call takes to
return (all launches enqueued) next to the time until the GPU is done.  python tools/host_bound.py [--batch 8]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvraft_b200 import RSF, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--iters', type=int, default=32)
ap.add_argument('--graph', action='store_true', help='replay the forward as one CUDA graph')
a = ap.parse_args()
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = RSF(bench.make_args()).to(dev).eval()
model.use_cuda_graph = a.graph
pc1, pc2 = [t.to(dev) for t in bench.synthetic_clouds(a.batch, bench.N_POINTS, 1234)]
with torch.no_grad():
    for _ in range(2):
        model([pc1, pc2], a.iters)
    for rep in range(3):
        torch.cuda.synchronize()
        n0 = ops.launch_count
        t0 = time.perf_counter()
        model([pc1, pc2], a.iters)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print('launches %d  enqueue %.2f ms  (%.1f us/launch)  gpu done %.2f ms' % (ops.launch_count - n0, (t1 - t0) * 1e3, (t1 - t0) * 1e6 / max(1, ops.launch_count - n0), (t2 - t0) * 1e3))
