"""kNN graph build time vs grid cell occupancy: python tools/bench_knn.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pvraft_b200 import ops
dev = torch.device('cuda:0')
pc, _ = [t.to(dev) for t in bench.synthetic_clouds(8, 8192, 1234)]
for occ in ('1.5', '3', '4', '6', '8', '12', '16'):
    os.environ['PVRAFT_KNN_OCC'] = occ
    for _ in range(2):
        ops.knn(pc, pc, 32, mode=0, want_rel=True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(5):
        ops.knn(pc, pc, 32, mode=0, want_rel=True)
    e.record(); torch.cuda.synchronize()
    print('occ', occ, 'knn %.3f ms' % (s.elapsed_time(e) / 5))
os.environ.pop('PVRAFT_KNN_OCC')
a = ops.knn(pc, pc, 32, use_sweep=True).sort(-1).values; c = ops.knn(pc, pc, 32, use_sweep=False).sort(-1).values
print('equal to brute force:', bool(torch.equal(a, c)))
