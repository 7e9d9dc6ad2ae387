"""Does a spatially coherent point order help the SetConv edge kernel (gathers of neighbour rows)?
python tools/bench_edge.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pvraft_b200 import ops, Graph
dev = torch.device('cuda:0')
b, n, c = 8, 8192, 64
pc, _ = [t.to(dev) for t in bench.synthetic_clouds(b, n, 1234)]


def morton_order(p):
    lo, hi = p.amin(1, keepdim=True), p.amax(1, keepdim=True)
    q = ((p - lo) / (hi - lo).clamp_min(1e-9) * 1023.0).long().clamp_(0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[..., 0]) | (spread(q[..., 1]) << 1) | (spread(q[..., 2]) << 2)
    return code.argsort(1)


def run(points, label):
    g = Graph.construct_graph(points, 32)
    x = torch.randn(b, n, c, device=dev)
    w = torch.randn(c, c + 3, device=dev)
    st = torch.zeros(b, 8, 2, dtype=torch.float64, device=dev)
    big = torch.randn(8192, 8192, device=dev)
    for _ in range(3):
        ops.setconv_edge(x, g.nbr, g._rel, w, c, st)
    for _ in range(2): big @ big
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        ops.setconv_edge(x, g.nbr, g._rel, w, c, st)
    e.record(); torch.cuda.synchronize()
    print(label, 'setconv_edge %.1f us' % (s.elapsed_time(e) * 1e3 / 20))


run(pc, 'input order ')
perm = morton_order(pc)
run(torch.gather(pc, 1, perm.unsqueeze(-1).expand(-1, -1, 3)).contiguous(), 'morton order')
