"""Stress the small-N ConvGRU (two tcgen05 launches) after the rest of the GPU suite has run in this process, and
report where a run differs from the first one.  python tools/flake_probe2.py [reps]"""
import os, sys
import pytest, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
pytest.main([os.path.join(ROOT, 'tests', 'test_gpu_parity.py'), '-q', '-m', 'gpu', '-x', '-k', 'lookup or topk or knn or corr_matmul'])
import test_gpu_parity as T
dev = torch.device('cuda:0')
arr, W, m = T.golden_model('small_rsf_refine.npz', dev, True)
T.install_golden_state(m, arr, dev)
g = T.golden_graph(arr, dev)
inp = torch.relu(arr['fct1'][:, 64:]).to(dev)
net = torch.tanh(arr['fct1'][:, :64].double()).float().to(dev)
flow = (arr['it0/coords'] - arr['pc1']).to(dev)
gcorr = arr['it0/corr'].to(dev)
bad = 0
with torch.no_grad():
    ref_net, ref_delta = m.update_block(net, inp, gcorr, flow, g)
    ref_net, ref_delta = ref_net.clone(), ref_delta.clone()
    for rep in range(reps):
        if rep % 7 == 0:
            junk = [torch.randn(4096 + 13 * rep, device=dev) for _ in range(rep % 4)]
        n2, d2 = m.update_block(net, inp, gcorr, flow, g)
        if not torch.equal(n2, ref_net):
            diff = (n2 - ref_net).abs()
            idx = (diff > 0).nonzero()
            bad += 1
            if bad <= 6:
                print('rep', rep, 'mismatch: count', idx.shape[0], 'max', float(diff.max()), 'layout [B,C,N] first', idx[:6].tolist(),
                      'unique b', idx[:, 0].unique().tolist(), 'unique c', idx[:, 1].unique().tolist()[:16], 'n range', int(idx[:, 2].min()), int(idx[:, 2].max()))
print('reps', reps, 'mismatching runs', bad)
