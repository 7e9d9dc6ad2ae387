import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvraft_b200 import ops
dev = torch.device('cuda:0')
corr = torch.randn(1, 64, 8192, device=dev)
val, idx = ops.corr_topk(corr, 512)
torch.cuda.synchronize()
print('ok', float(val.sum()))
