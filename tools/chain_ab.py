"""A/B of the per-sample chained tensor-core launches on the bench workload: PVRAFT_TC_CHAIN=0/1 python tools/chain_ab.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pvraft_b200 import RSF  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
model = RSF(bench.make_args()).to(dev).eval()
model.use_cuda_graph = False   # the comparison is about eager launches
pc1, pc2 = [t.to(dev) for t in bench.synthetic_clouds(8, bench.N_POINTS, 1234)]
with torch.no_grad():
    for _ in range(4):
        out = model([pc1, pc2], 32)[-1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = model([pc1, pc2], 32)[-1]
    e1.record()
    torch.cuda.synchronize()
print('chain', os.environ.get('PVRAFT_TC_CHAIN', '0'), 'ms/forward', e0.elapsed_time(e1) / 10, 'checksum', float(out.double().abs().sum()))
