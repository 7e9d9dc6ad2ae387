"""Time the per-row top-K (B=8, N=M=8192, K=512): python tools/bench_topk.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvraft_b200 import ops
dev = torch.device('cuda:0')
corr = torch.randn(8, 8192, 8192, device=dev)
for _ in range(2):
    ops.corr_topk(corr, 512)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
s.record()
for _ in range(5):
    ops.corr_topk(corr, 512)
e.record(); torch.cuda.synchronize()
print('corr_topk %.3f ms  (%.0f GB/s of the 2.1 GB matrix)' % (s.elapsed_time(e) / 5, corr.numel() * 4 / (s.elapsed_time(e) / 5) / 1e6))
