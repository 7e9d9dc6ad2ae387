"""Tensor-level wrappers over the C ABI (include/pvraft_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every arithmetic step of the
hot path happens inside libpvraft_b200.so.  All wrappers require contiguous CUDA tensors and raise
on anything else -- there is deliberately no CPU / eager fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, IN_GN, IN_GN_MINMAX, IN_PLAIN, KNN, MOMENTS, check, lib)

launch_count = 0   # C-ABI calls that launch a kernel (bench.py reports it as gpu_launches)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=torch.float32):
    if t is None:
        return None
    if not torch.is_tensor(t):
        raise TypeError(f'expected a tensor, got {type(t)}')
    if not t.is_cuda:
        raise _lib.PvraftError('pvraft_b200 kernels need CUDA tensors (no CPU fallback exists)')
    if t.dtype != dtype:
        raise TypeError(f'expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError('expected a contiguous tensor')
    return t.data_ptr()


def _count(rc, what):
    global launch_count
    check(rc, what)
    launch_count += 1


def new_stats(b, device, n=1):
    """Zeroed GroupNorm accumulators: n x [B,8,2] doubles (one cudaMemset for all of them)."""
    return torch.zeros(n, b, 8, 2, dtype=torch.float64, device=device)


def corr_reorder(val, idx):
    """Bank-aware permutation of every row of the truncated state (val [B,N,K] f32, idx [B,N,K] int32)."""
    b, n, k = val.shape
    val_out, idx_out = torch.empty_like(val), torch.empty_like(idx)
    _count(lib().pvraft_corr_reorder(_p(val), _p(idx, torch.int32), b * n, k, _p(val_out), _p(idx_out, torch.int32), _stream()),
           'corr_reorder')
    return val_out, idx_out


def corr_matmul(fmap1_pm, fmap2_pm):
    """Point-major feature maps [B,N,C] -> all-pairs correlation [B,N,N] / sqrt(C) on tcgen05 (3xTF32)."""
    b, n, c = fmap1_pm.shape
    corr = torch.empty(b, n, n, dtype=torch.float32, device=fmap1_pm.device)
    ws = torch.empty(int(lib().pvraft_corr_matmul_workspace_bytes(b, n, c)), dtype=torch.uint8, device=fmap1_pm.device)
    _count(lib().pvraft_corr_matmul_fwd(_p(fmap1_pm), _p(fmap2_pm), b, n, c, _p(corr), _p(ws, torch.uint8), _stream()),
           'corr_matmul')
    return corr


def corr_topk(corr, k):
    """corr [B,N,M] -> (val [B,N,K] f32, idx [B,N,K] int32): the K largest per row, ascending column order."""
    b, n, m = corr.shape
    val = torch.empty(b, n, k, dtype=torch.float32, device=corr.device)
    idx = torch.empty(b, n, k, dtype=torch.int32, device=corr.device)
    _count(lib().pvraft_corr_topk_fwd(_p(corr), b, n, m, k, _p(val), _p(idx, torch.int32), _stream()), 'corr_topk')
    return val, idx


def corr_lookup(corr_val, corr_idx, xyz2, coords, levels, base_scale, vox=None, knn_sel=None, moments=None,
                want_slots=False, want_cube=False):
    """-> dict(vox [B,N,pad4(levels*27)], knn_sel [B,N,32,4], moments [B,16] f64, [knn_slot], [cube])."""
    b, n, k = corr_val.shape
    dev = corr_val.device
    vox_ld = (levels * 27 + 3) // 4 * 4          # rows padded to a multiple of 4 floats (zero-filled by the kernel)
    if vox is None:
        vox = torch.empty(b, n, vox_ld, dtype=torch.float32, device=dev)
    if knn_sel is None:
        knn_sel = torch.empty(b, n, KNN, 4, dtype=torch.float32, device=dev)
    if moments is None:
        moments = torch.zeros(b, MOMENTS, dtype=torch.float64, device=dev)
    slots = torch.empty(b, n, KNN, dtype=torch.int32, device=dev) if want_slots else None
    cube = torch.empty(b, n, k, levels, dtype=torch.int8, device=dev) if want_cube else None
    _count(lib().pvraft_corr_lookup_fwd(_p(corr_val), _p(corr_idx, torch.int32), _p(xyz2), _p(coords), b, n, k, levels,
                                        float(base_scale), _p(vox), vox.shape[-1], _p(knn_sel), _p(slots, torch.int32),
                                        _p(moments, torch.float64), _p(cube, torch.int8), _stream()), 'corr_lookup')
    return dict(vox=vox, knn_sel=knn_sel, moments=moments, knn_slot=slots, cube=cube)


def linear(x, weight, bias=None, *, cin=None, w_ld=0, w_cin=0, in_mode=IN_PLAIN, in_min=None, in_stats=None, in_gamma=None,
           in_beta=None, in_count=0.0, in_act=ACT_NONE, in_slope=0.0, out_act=ACT_NONE, out=None, out_stats=None,
           residual=None, cout=None):
    """Fused [GN -> act ->] 1x1 conv [+bias] [-> ReLU] [+ residual] over x [B,N,cin] -> [B,N,cout]."""
    b, n, c = x.shape
    cin = c if cin is None else cin
    cout = weight.shape[0] if cout is None else cout
    if out is None:
        out = torch.empty(b, n, cout, dtype=torch.float32, device=x.device)
    a = _lib.LinearArgs(_p(x), _p(in_min), _p(in_stats, torch.float64), _p(in_gamma), _p(in_beta), float(in_count),
                        in_mode, in_act, float(in_slope), _p(weight), int(w_ld), int(w_cin), _p(bias), _p(residual), out_act,
                        _p(out), _p(out_stats, torch.float64), b, n, cin, cout)
    _count(lib().pvraft_linear_fwd(C.byref(a), _stream()), 'linear')
    return out


def gn_act(x, stats, gamma, beta, count, act=ACT_LRELU, slope=0.1, transpose_out=False):
    b, n, c = x.shape
    out = torch.empty((b, c, n) if transpose_out else (b, n, c), dtype=torch.float32, device=x.device)
    _count(lib().pvraft_gn_act_fwd(_p(x), _p(stats, torch.float64), _p(gamma), _p(beta), float(count), act, float(slope),
                                   b, n, c, int(transpose_out), _p(out), _stream()), 'gn_act')
    return out


def transpose(x):
    """[B,R,C] -> [B,C,R] contiguous."""
    b, r, c = x.shape
    out = torch.empty(b, c, r, dtype=torch.float32, device=x.device)
    _count(lib().pvraft_transpose_fwd(_p(x), b, r, c, _p(out), _stream()), 'transpose')
    return out


def corr_feature(args):
    _count(lib().pvraft_corr_feature_fwd(C.byref(args), _stream()), 'corr_feature')


def gru(args):
    _count(lib().pvraft_gru_fwd(C.byref(args), _stream()), 'gru')


def flow_out(args):
    _count(lib().pvraft_flow_out_fwd(C.byref(args), _stream()), 'flow_out')


def setconv_edge(fc1p, nbr, edge_feats, w_fc1, cin, stats, ymax=None, ymin=None):
    b, n, c = fc1p.shape
    if ymax is None:
        ymax = torch.empty_like(fc1p)
    if ymin is None:
        ymin = torch.empty_like(fc1p)
    _count(lib().pvraft_setconv_edge_fwd(_p(fc1p), _p(nbr, torch.int32), _p(edge_feats), _p(w_fc1), cin, b, n, c, _p(ymax),
                                         _p(ymin), _p(stats, torch.float64), _stream()), 'setconv_edge')
    return ymax, ymin


def knn(xyz, query, k, mode=0, want_rel=False, use_sweep=True):
    """-> int32 [B,S,k] local ids of the k nearest `xyz` points of every query (unordered)
    [, rel [B,S,k,3] = xyz[idx] - query]."""
    b, n, _ = xyz.shape
    s = query.shape[1]
    out = torch.empty(b, s, k, dtype=torch.int32, device=xyz.device)
    rel = torch.empty(b, s, k, 3, dtype=torch.float32, device=xyz.device) if want_rel else None
    ws_bytes = int(lib().pvraft_knn_workspace_bytes(b, n)) if use_sweep else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device) if ws_bytes > 0 else None
    _count(lib().pvraft_knn_fwd(_p(xyz), _p(query), b, n, s, k, mode, _p(out, torch.int32), _p(rel),
                                _p(ws, torch.uint8), _stream()), 'knn')
    return (out, rel) if want_rel else out


def device_info():
    sm, smem = C.c_int(0), C.c_int(0)
    check(lib().pvraft_device_info(C.byref(sm), C.byref(smem)), 'device_info')
    return sm.value, smem.value
