"""Diagnostic (GPU box): where does the pre-loop path at N=8192 leave the oracle?  per-layer, per-sample errors."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from conftest import default_weights
from oracle import pvraft_oracle as O
from pvraft_b200 import RSF, Graph

torch.set_num_threads(16)
dev = torch.device('cuda:0')
N, K = 8192, 512
args = types.SimpleNamespace(corr_levels=3, base_scales=0.25, truncate_k=K)
W = default_weights(args=args)
m = RSF(args); m.load_state_dict(W); m = m.to(dev).eval()
b = 2
pc1, pc2 = O.synthetic_clouds(b, N, seed=1234)


def rel(a, b_):
    a, b_ = a.double(), b_.double()
    return [float((a[i] - b_[i]).abs().max() / b_[i].abs().max()) for i in range(a.shape[0])]


with torch.no_grad():
    og = O.construct_graph(pc1, 32)
    g = Graph.construct_graph(pc1.to(dev), 32)
    nb = g.nbr.long().cpu().sort(-1).values
    ref = (og.edges.reshape(b, N, 32) - (torch.arange(b) * N).view(b, 1, 1)).sort(-1).values
    print('graph rows differing per sample:', (nb != ref).any(-1).float().mean(1).tolist())
    d = O.pairwise_sqdist_expanded(pc1)
    bad = (nb != ref).any(-1)
    kth_a = torch.gather(d, 2, nb).max(-1).values
    kth_b = torch.gather(d, 2, ref).max(-1).values
    print('  of which not an exact tie at the 32nd distance:', ((kth_a != kth_b) & bad).float().mean(1).tolist())
    # encoder layers with the ORACLE graph on both sides
    from pvraft_b200.graph import Graph as PG
    pg = PG(ref.to(torch.int32).to(dev), og.edge_feats.reshape(b, N, 32, 3).to(dev).contiguous(), 32, [b * N, b * N])
    x = pc1
    enc = m.feature_extractor
    for name in ('feat_conv1', 'feat_conv2', 'feat_conv3'):
        want = O.set_conv(W, 'feature_extractor.' + name, x, og)
        got = getattr(enc, name)(x.to(dev), pg)
        print(name, 'same graph, oracle input: rel err per sample', rel(got.cpu(), want))
        x = want
    fm_o, _ = O.flot_encoder(W, 'feature_extractor', pc1, og)
    fm_p, _ = enc(pc1.to(dev), graph=pg)
    print('encoder (same graph) rel', rel(fm_p.cpu(), fm_o))
    fm_p2, _ = enc(pc1.to(dev), graph=g)
    print('encoder (own graph) rel', rel(fm_p2.cpu(), fm_o))
    both = torch.cat([pc1, pc2], 0).to(dev)
    fm_b, g2 = enc(both, point_major=True)
    fm2_o, _ = O.flot_encoder(W, 'feature_extractor', pc2)
    print('encoder batched 2B rel: pc1', rel(fm_b[:b].transpose(1, 2).cpu(), fm_o), 'pc2', rel(fm_b[b:].transpose(1, 2).cpu(), fm2_o))
    # correlation + topk from ORACLE feature maps
    st = O.corr_init(fm_o, fm2_o, pc2, K)
    m.corr_block.init_module(fm_o.to(dev), fm2_o.to(dev), pc2.to(dev))
    print('truncated corr from oracle fmaps rel', rel(m.corr_block.truncated_corr.cpu(), st.truncated_corr))
    print('  rows with different candidate sets', (m.corr_block.corr_idx.long().cpu().sort(-1).values != st.indices.sort(-1).values).any(-1).float().mean(1).tolist())
