"""Run ONE training step (BASELINE configs[3] shape: B=2, N=8192, 8 iterations) inside a cudaProfilerStart/Stop range (for ncu
--profile-from-start off).  python tools/profile_train.py [--batch 2] [--iters 8]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvraft_b200 import RSF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=2)
ap.add_argument('--iters', type=int, default=8)
a = ap.parse_args()
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = RSF(bench.make_args()).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
pc1, pc2 = [t.to(dev) for t in bench.synthetic_clouds(a.batch, bench.N_POINTS, 1234)]


def step():
    opt.zero_grad(set_to_none=True)
    flows = model([pc1, pc2], num_iters=a.iters)
    n = len(flows)
    loss = sum(0.8 ** (n - i - 1) * (flows[i] - (pc2 - pc1)).abs().sum(-1).mean() for i in range(n))
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
