"""Pin the CPU oracle (oracle/pvraft_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import pvraft_oracle as O

TOL = 2e-5   # max-abs / max-abs; the reference's own fp32-vs-fp64 spread is ~3e-7


def graph_from(arr):
    e = arr['graph_edges']
    b, n, k = e.shape
    return O.Graph(e.reshape(-1).long(), arr['graph_edge_feats'], k, (b * n, b * n))


def state_from(arr):
    txyz = arr['truncate_xyz2']
    return O.CorrState(arr['truncated_corr'], None, txyz)


@pytest.mark.parametrize('fixture', ['small_rsf_refine.npz', 'oddscale_rsf.npz'])
def test_teacher_forced_modules(fixture):
    arr, W = load_golden(fixture)
    b, n, k, levels, iters = [int(v) for v in arr['meta']]
    bs = float(arr['base_scale'])
    st, g = state_from(arr), graph_from(arr)
    inp = torch.relu(arr['fct1'][:, 64:])
    net = torch.tanh(arr['fct1'][:, :64].double()).float()   # (fp32 CPU tanh: see test_gpu_parity)
    for it in range(iters):
        coords = arr[f'it{it}/coords']
        vox = O.voxel_feature(W, st, coords, levels, bs)
        knn = O.knn_feature(W, st, coords)
        assert rel_err(vox, arr[f'it{it}/voxel_feature']) < TOL
        assert rel_err(knn, arr[f'it{it}/knn_feature']) < TOL
        corr = arr[f'it{it}/corr']
        flow = coords - arr['pc1']
        mot = O.motion_encoder(W, flow, corr, 'update_block.motion_encoder')
        assert rel_err(mot, arr[f'it{it}/motion']) < TOL
        net2, delta = O.update_block(W, net, inp, corr, flow, g)
        assert rel_err(net2, arr[f'it{it}/net']) < TOL
        assert rel_err(delta, arr[f'it{it}/delta']) < 5e-5
        net = arr[f'it{it}/net']


@pytest.mark.parametrize('fixture', ['small_rsf_refine.npz', 'oddscale_rsf.npz'])
def test_indices_bit_exact(fixture):
    arr, W = load_golden(fixture)
    b, n, k, levels, iters = [int(v) for v in arr['meta']]
    bs = float(arr['base_scale'])
    st = state_from(arr)
    it = 1
    coords = arr[f'it{it}/coords']
    for lvl in range(levels):
        cube, valid = O.voxel_cube_index(st, coords, bs * 2 ** lvl)
        assert torch.equal(cube.to(torch.int8), arr[f'it{it}/cube_idx_l{lvl}'])
        assert torch.equal(valid, arr[f'it{it}/valid_l{lvl}'])
    d = O.knn_sqdist(st, coords)
    assert torch.equal(d, arr[f'it{it}/knn_dist'])          # bit-exact fp32 distances
    slots = O.knn_select(st, coords)
    assert torch.equal(torch.sort(slots, -1).values, torch.sort(arr[f'it{it}/knn_slots'].long(), -1).values)


def test_encoder_corr_init_and_free_running_small():
    arr, W = load_golden('small_rsf_refine.npz')
    b, n, k, levels, iters = [int(v) for v in arr['meta']]
    pc1, pc2 = arr['pc1'], arr['pc2']
    li = O.prepare(W, pc1, pc2, k)
    assert torch.equal(li.graph.edges.reshape(b, n, -1).sort(-1).values,
                       arr['graph_edges'].long().sort(-1).values)
    assert rel_err(li.state.truncated_corr, arr['truncated_corr']) < TOL
    # candidate SETS must agree; order may swap at value near-ties (rounding of the encoder GEMMs)
    sa = li.state.truncate_xyz2[..., 0].sort(-1).values
    sb = arr['truncate_xyz2'][..., 0].sort(-1).values
    assert (sa != sb).any(-1).float().mean() <= 0.01
    flows = O.raft_loop(W, li, pc1, iters, levels, float(arr['base_scale']))
    for it in range(iters):
        assert rel_err(flows[it], arr[f'it{it}/flow']) < 1e-4
    refined = O.flot_refine(W, 'refine_block', flows[-1], li.feat_graph)
    assert rel_err(refined, arr['refined']) < 1e-4


def test_medium_default_init_free_running():
    arr, _ = load_golden('medium_rsf.npz')
    from conftest import default_weights
    W = default_weights()
    b, n, k, levels, iters = [int(v) for v in arr['meta']]
    li = O.prepare(W, arr['pc1'], arr['pc2'], k)
    cs = arr['truncated_corr_checksum']
    assert abs(float(li.state.truncated_corr.double().sum()) - float(cs[0])) < 1e-6 * float(cs[1])
    trace = []
    flows = O.raft_loop(W, li, arr['pc1'], iters, levels, float(arr['base_scale']), trace)
    for it in range(iters):
        assert rel_err(trace[it]['corr'], arr[f'it{it}/corr']) < 1e-4
        assert rel_err(flows[it], arr[f'it{it}/flow']) < 1e-3
    for lvl in range(levels):
        cube, _ = O.voxel_cube_index(li.state, arr['it1/coords'], 0.25 * 2 ** lvl)
        mism = (cube.to(torch.int8) != arr[f'cube_idx_l{lvl}']).float().mean()
        assert mism < 1e-4     # free-running coords differ in the last ulp -> allow rare flips


def test_knn_point_golden():
    arr, _ = load_golden('knn_point.npz')
    idx = O.knn_point(16, arr['xyz'], arr['query'])
    assert torch.equal(idx.sort(-1).values.int(), arr['idx'])
