"""Small end-to-end workload for compute-sanitizer: inference forward (RSF_refine, fp32 and bf16 state, eager launches),
a lookup with the debug outputs, and one training step.  python tools/sanitize_workload.py"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvraft_b200 import RSF, RSF_refine  # noqa: E402

dev = torch.device('cuda:0')
args = types.SimpleNamespace(corr_levels=3, base_scales=0.25, truncate_k=128)
g = torch.Generator().manual_seed(0)
pc1 = (3.0 * torch.rand(2, 512, 3, generator=g)).to(dev)
pc2 = pc1 + 0.1 * torch.randn(2, 512, 3, generator=g).to(dev)
torch.manual_seed(0)
m = RSF_refine(args).to(dev).eval()
m.use_cuda_graph = False
with torch.no_grad():
    out = m([pc1, pc2], 2)
    m.set_precision('bf16')
    out16 = m([pc1, pc2], 2)
    m.set_precision('fp32')
    m.corr_block.lookup(pc1, want_slots=True, want_cube=True)
torch.manual_seed(0)
t = RSF(args).to(dev).train()
flows = t([pc1, pc2], num_iters=2)
sum((f - (pc2 - pc1)).abs().mean() for f in flows).backward()
torch.cuda.synchronize()
print('sanitize workload done', float(out.abs().mean()), float(out16.abs().mean()))
