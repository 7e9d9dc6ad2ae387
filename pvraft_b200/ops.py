"""Tensor-level wrappers over the C ABI (include/pvraft_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every arithmetic step of the
hot path happens inside libpvraft_b200.so.  All wrappers require contiguous CUDA tensors and raise
on anything else -- there is deliberately no CPU / eager fallback.
"""
import ctypes as C
import os
import threading
import weakref

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, IN_GN, IN_GN_MINMAX, IN_PLAIN, KNN, MOMENTS, check, lib)

launch_count = 0   # C-ABI calls that launch a kernel (bench.py reports it as gpu_launches)


def _stream():
    """The current stream of the CURRENT device: the library launches on the current device, so every operand has to live
    there (`_p` checks it) -- RSF.forward enters `torch.cuda.device(input.device)` itself; callers of the inner modules on
    a non-default GPU do the same (or `torch.cuda.set_device`), as under DDP / DataParallel."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=torch.float32):
    if t is None:
        return None
    if not torch.is_tensor(t):
        raise TypeError(f'expected a tensor, got {type(t)}')
    if not t.is_cuda:
        raise _lib.PvraftError('pvraft_b200 kernels need CUDA tensors (no CPU fallback exists)')
    if t.device.index != torch.cuda.current_device():
        raise _lib.PvraftError(f'tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}: the kernels '
                               'launch on the current device -- wrap the call in `with torch.cuda.device(t.device):`')
    if t.dtype != dtype:
        raise TypeError(f'expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError('expected a contiguous tensor')
    return t.data_ptr()


def _count(rc, what):
    global launch_count
    check(rc, what)
    launch_count += 1


_TLS = threading.local()   # .arena = [buffer [n,B,8,2] f64, next free block]: zeroed accumulators handed out inside a
                           # `stats_arena` scope (per thread: nn.DataParallel runs one replica per thread)


class stats_arena:
    """Scope in which `new_stats` (and the lookup's moment accumulator) are slices of ONE zero-filled buffer instead of a
    memset launch each: the RAFT loop needs 5 blocks per iteration (`with ops.stats_arena(b, dev, 5 * iters)`)."""

    def __init__(self, b, device, blocks):
        self.buf = [torch.zeros(blocks, b, 8, 2, dtype=torch.float64, device=device), 0]

    def __enter__(self):
        self.prev = getattr(_TLS, 'arena', None)
        _TLS.arena = self.buf
        return self

    def __exit__(self, *exc):
        _TLS.arena = self.prev
        _TLS.last_tc = None   # (drops the operand references a launch chain holds)
        return False


def new_flags(b, device):
    """[B] zeroed int32 `done` counters for one tensor-core launch, carved out of the arena (None outside an arena scope:
    no chaining there)."""
    arena = getattr(_TLS, 'arena', None)
    if arena is None:
        return None
    buf, pos = arena
    if pos + 1 > buf.shape[0] or buf.shape[1] != b or buf.device != torch.device(device):
        return None
    arena[1] = pos + 1
    return buf[pos].view(torch.int32).reshape(-1)[:b]


def new_stats(b, device, n=1):
    """Zeroed GroupNorm accumulators: n x [B,8,2] doubles (one cudaMemset for all of them)."""
    arena = getattr(_TLS, 'arena', None)
    if arena is not None:
        buf, pos = arena
        if pos + n <= buf.shape[0] and buf.shape[1] == b and buf.device == torch.device(device):
            arena[1] = pos + n
            return buf[pos:pos + n]
    return torch.zeros(n, b, 8, 2, dtype=torch.float64, device=device)


def corr_reorder(val, idx):
    """Bank-aware permutation of every row of the truncated state (val [B,N,K] f32, idx [B,N,K] int32)."""
    b, n, k = val.shape
    val_out, idx_out = torch.empty_like(val), torch.empty_like(idx)
    _count(lib().pvraft_corr_reorder(_p(val), _p(idx, torch.int32), b * n, k, _p(val_out), _p(idx_out, torch.int32), _stream()),
           'corr_reorder')
    return val_out, idx_out


def corr_state_pack_bf16(val, idx):
    """Reordered fp32 / int32 state -> (bf16 values, uint16 ids held in an int16 tensor): 4 B per candidate and iteration."""
    if int(idx.shape[1]) > 65536:
        raise ValueError('uint16 candidate ids need N <= 65536')
    v16 = torch.empty(val.shape, dtype=torch.bfloat16, device=val.device)
    i16 = torch.empty(idx.shape, dtype=torch.int16, device=idx.device)
    _count(lib().pvraft_corr_state_pack_bf16(_p(val), _p(idx, torch.int32), val.numel(), _p(v16, torch.bfloat16), _p(i16, torch.int16),
                                             _stream()), 'corr_state_pack_bf16')
    return v16, i16


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


def xyz_pad(xyz):
    """[B,N,3] -> [B,N,4] = (x,y,z,0): the lookup kernel's gather table (one 128-bit load per candidate), built once per forward."""
    b, n, _ = xyz.shape
    out = torch.empty(b, n, 4, dtype=torch.float32, device=xyz.device)
    _count(lib().pvraft_xyz_pad_fwd(_p(xyz), b * n, _p(out), _stream()), 'xyz_pad')
    return out


def corr_lookup(corr_val, corr_idx, xyz2_pad, coords, levels, base_scale, vox=None, knn_sel=None, moments=None,
                want_slots=False, want_cube=False, vox_ld=None):
    """-> dict(vox [B,N,pad4(levels*27)], knn_sel [B,N,32,4], moments [B,16] f64, [knn_slot], [cube]).
    `cube` [B,N,K,levels] int8 is the fused kernel's own cell decision for every candidate (test hook)."""
    b, n, k = corr_val.shape
    dev = corr_val.device
    if vox_ld is None:
        vox_ld = (levels * 27 + 3) // 4 * 4      # rows padded to a multiple of 4 floats (zero-filled by the kernel)
    if vox is None:
        vox = torch.empty(b, n, vox_ld, dtype=torch.float32, device=dev)
    if knn_sel is None:
        knn_sel = torch.empty(b, n, KNN, 4, dtype=torch.float32, device=dev)
    if moments is None:
        moments = new_stats(b, dev, 1).view(b, MOMENTS) if MOMENTS == 16 else torch.zeros(b, MOMENTS, dtype=torch.float64, device=dev)
    slots = torch.empty(b, n, KNN, dtype=torch.int32, device=dev) if want_slots else None
    cube = torch.empty(b, n, k, levels, dtype=torch.int8, device=dev) if want_cube else None
    if corr_val.dtype == torch.bfloat16:      # reduced-precision state: bf16 values + uint16 ids (stored as int16)
        _count(lib().pvraft_corr_lookup_bf16_fwd(_p(corr_val, torch.bfloat16), _p(corr_idx, torch.int16), _p(xyz2_pad), _p(coords), b, n,
                                                 k, levels, float(base_scale), _p(vox), vox.shape[-1], _p(knn_sel),
                                                 _p(slots, torch.int32), _p(moments, torch.float64), _p(cube, torch.int8), _stream()),
               'corr_lookup_bf16')
    else:
        _count(lib().pvraft_corr_lookup_fwd(_p(corr_val), _p(corr_idx, torch.int32), _p(xyz2_pad), _p(coords), b, n, k, levels,
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


_TC_WEIGHTS = {}
_EARLY_PARAMS = os.environ.get('PVRAFT_TC_EARLY_PARAMS', '1') != '0'
_CHAIN = os.environ.get('PVRAFT_TC_CHAIN', '0') == '1'   # opt-in: see tc_linear(chain=...)
TC_PLAIN, TC_GRU_ZR, TC_GRU_Q, TC_FLOW = 0, 1, 2, 3


def tc_weights(weights, col0=0, cols=None, k_pad=None, kcat=False, transposed=None):
    """tf32 hi/lo split of a (stack of) [cout, cin(,1,1)] weight(s) -> (hi, lo) [n_pad, k_pad], cached per parameter
    version (inference weights are static, so this runs once; the training path re-splits after every optimizer step).
    transposed=(r0, r1): the split of W^T[r0:r1, :] instead -- the operand of dx = dy . W for input columns r0..r1."""
    if torch.is_tensor(weights):
        weights = (weights,)
    # keyed by the identity of the source tensor OBJECTS (validated through weak references and version counters):
    # a data_ptr key would go stale when the allocator hands a freed weight's address to a new tensor
    key = tuple(id(w) for w in weights) + (col0, cols, k_pad, kcat, transposed)
    hit = _TC_WEIGHTS.get(key)
    if hit is not None:
        refs, versions, ptrs, result = hit
        if all(r() is w and w._version == v and w.data_ptr() == p for r, w, v, p in zip(refs, weights, versions, ptrs)):
            return result
    _TLS.unsettled = 3   # the split below writes what the next tensor-core launches read: see tc_linear()
    mats = [w.detach().reshape(w.shape[0], -1) for w in weights]
    if transposed is not None:
        mats = [m.t()[transposed[0]:transposed[1]].contiguous() for m in mats]
    if kcat:   # [W_a | W_b | ...] along K: one GEMM over concatenated sources adds the layers' outputs
        mats = [torch.cat(mats, 1).contiguous()]
    ld = mats[0].shape[1]
    ncols = ld - col0 if cols is None else cols
    kp = (ncols + 31) // 32 * 32 if k_pad is None else k_pad
    rows = sum(m.shape[0] for m in mats)
    n_pad = (rows + 15) // 16 * 16
    # (the split kernel writes every entry of the rows it is given, padding columns included: zero-fill only for padding rows)
    alloc = torch.empty if n_pad == rows else torch.zeros
    hi = alloc(n_pad, kp, dtype=torch.float32, device=mats[0].device)
    lo = alloc(n_pad, kp, dtype=torch.float32, device=mats[0].device)
    r0 = 0
    for m in mats:
        check(lib().pvraft_tc_weight_split(_p(m.contiguous()), m.shape[0], ncols, ld, col0, m.shape[0], kp,
                                           hi[r0:].data_ptr(), lo[r0:].data_ptr(), _stream()), 'tc_weight_split')
        r0 += m.shape[0]
    if len(_TC_WEIGHTS) > 512:
        _TC_WEIGHTS.clear()
    result = (hi, lo, n_pad, rows)
    _TC_WEIGHTS[key] = (tuple(weakref.ref(w) for w in weights), tuple(w._version for w in weights),
                        tuple(w.data_ptr() for w in weights), result)
    return result


_DERIVED = {}


def derived(tensors, tag, fn):
    """Cache of small host- or device-side values derived from parameters (a bias sum, a PReLU slope read back once);
    keyed by parameter identity, re-derived when a parameter's version or storage changes."""
    key = tuple(id(t) for t in tensors) + (tag,)
    hit = _DERIVED.get(key)
    if hit is not None:
        refs, versions, ptrs, value = hit
        if all(r() is t and t._version == v and t.data_ptr() == p for r, t, v, p in zip(refs, tensors, versions, ptrs)):
            return value
    _TLS.unsettled = 3   # fn may launch kernels that write a folded weight / bias
    value = fn(*tensors)
    if len(_DERIVED) > 512:
        _DERIVED.clear()
    _DERIVED[key] = (tuple(weakref.ref(t) for t in tensors), tuple(t._version for t in tensors),
                     tuple(t.data_ptr() for t in tensors), value)
    return value


def point_order(points, as_int32=False):
    """[B,N,3] -> [B,N] int64 (int32 with as_int32): Morton order over the cells of the kNN grid, from the library's
    in-shared-memory sort (one launch; the torch formulation below costs ~40 launches)."""
    b, n, _ = points.shape
    ws_bytes = int(lib().pvraft_knn_workspace_bytes(b, n))
    if ws_bytes <= 0 or n < 64:
        perm = morton_order(points)
        return perm.to(torch.int32).contiguous() if as_int32 else perm
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points.device)
    perm = torch.empty(b, n, dtype=torch.int32, device=points.device)
    _count(lib().pvraft_point_order_fwd(_p(points), b, n, _p(perm, torch.int32), _p(ws, torch.uint8), _stream()), 'point_order')
    return perm if as_int32 else perm.long()


def morton_order(points):
    """[B,N,3] -> [B,N] int64 permutation that sorts every cloud along a 30-bit Morton (Z-order) curve: neighbouring points
    get neighbouring rows, so the 32 neighbour rows a SetConv gathers for consecutive points overlap in L1/L2."""
    lo, hi = points.amin(1, keepdim=True), points.amax(1, keepdim=True)
    q = ((points - lo) / (hi - lo).clamp_min(1e-12) * 1023.0).long().clamp_(0, 1023)

    def spread(v):   # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        return (v | (v << 2)) & 0x09249249

    code = spread(q[..., 0]) | (spread(q[..., 1]) << 1) | (spread(q[..., 2]) << 2)
    return torch.sort(code, dim=1, stable=True).indices


def tc_supported(n_points, *channels):
    """The tcgen05 layer needs 128-point tiles that do not straddle samples and 32-channel k-blocks."""
    return n_points % 128 == 0 and all(c % 32 == 0 for c in channels)


def tc_linear(sources, w, bias=None, *, in_min=None, in_stats=None, in_gamma=None, in_beta=None, in_count=0.0,
              in_act=ACT_NONE, in_slope=0.0, out_act=ACT_NONE, residual=None, out=None, out_stats=None, epilogue=TC_PLAIN,
              bias2=None, out2=None, h=None, z=None, cout=None, tail=None, w3=None, b3=None, coords1=None, coords2=None,
              coords2_out=None, flow_out=None, flow_user=None, row_map=None, chain=False):
    """Fused layer on the tcgen05 tensor cores.  sources: list of [B,N,C_i] tensors concatenated along K (the
    GroupNorm prologue applies to sources[0]); w = (hi, lo, n_pad, rows) from tc_weights(); tail [B,N,3] fills the
    output columns cout..cout+2.
    chain=True: the caller states that everything this layer reads was produced by the tensor-core launch issued
    immediately before it (or earlier).  Inside a `stats_arena` scope the kernel then waits per SAMPLE on that launch's
    completion counters instead of on the whole grid, so CTAs start while the previous layer's last tiles still run."""
    hi, lo, n_pad, rows = w
    b, n, _ = sources[0].shape
    cout = rows if cout is None else cout
    if out is None:
        out = torch.empty(b, n, cout + (3 if tail is not None else 0), dtype=torch.float32, device=sources[0].device)
    a = _lib.TcLinearArgs()
    for i, src in enumerate(sources):
        a.in_[i] = _p(src)
        a.in_channels[i] = src.shape[-1]
    a.in_min, a.in_stats, a.in_gamma, a.in_beta = _p(in_min), _p(in_stats, torch.float64), _p(in_gamma), _p(in_beta)
    a.in_count, a.in_act, a.in_slope = float(in_count), in_act, float(in_slope)
    a.w_hi, a.w_lo, a.n_pad, a.cout = _p(hi), _p(lo), n_pad, cout
    a.bias, a.bias2, a.out_act, a.residual = _p(bias), _p(bias2), out_act, _p(residual)
    a.out, a.out2, a.h, a.z = _p(out), _p(out2), _p(h), _p(z)
    a.out_stats, a.epilogue, a.B, a.N = _p(out_stats, torch.float64), epilogue, b, n
    a.tail = _p(tail)
    a.w3, a.b3, a.coords1, a.coords2 = _p(w3), _p(b3), _p(coords1), _p(coords2)
    a.coords2_out, a.flow_out = _p(coords2_out), _p(flow_out)
    a.flow_user, a.row_map = _p(flow_user), _p(row_map, torch.int32)
    # The kernel may fetch its parameters while the previous kernel drains (PDL) once they are settled: not during the
    # three tensor-core launches that follow a weight split or a re-derived folded parameter on this thread.
    pending = getattr(_TLS, 'unsettled', 0)
    a.params_settled = 1 if pending == 0 and _EARLY_PARAMS else 0
    if pending:
        _TLS.unsettled = pending - 1
    done = new_flags(b, out.device) if _CHAIN else None
    a.done = _p(done, torch.int32)
    prev = getattr(_TLS, 'last_tc', None)
    # chain only on the launch that directly precedes this one (no other library launch in between), same tiling
    # (worth it only when CTAs walk several tiles: with one tile per CTA the per-sample spin costs more than the grid-wide
    #  wait it replaces -- measured -1.7 % at B=2, N=8192 = 128 tiles, +0.8 % at 512 tiles with eager launches and nothing
    #  under graph replay, which is why it is opt-in, PVRAFT_TC_CHAIN=1)
    chained = (chain and done is not None and prev is not None and prev[0] == launch_count and prev[1] is not None
               and prev[2] == (b, n) and bool(a.params_settled) and b * n // 128 > _sm_count())
    if chained:
        a.wait_on = _p(prev[1], torch.int32)
    _count(lib().pvraft_tc_linear_fwd(C.byref(a), _stream()), 'tc_linear')
    # A chained successor runs while its predecessors are still reading their operands, so the allocator must not hand that
    # storage out for a successor's outputs: the operands of a whole chain stay referenced until a launch with a full
    # grid-wide wait follows (chains are short: motion -> zr -> q -> P, fc3 -> flow head).
    keep = (sources, in_min, residual, out, out2, h, z, tail, coords2) if done is not None else None
    _TLS.last_tc = (launch_count, done, (b, n), ((keep,) + (prev[3] if chained else ())) if keep is not None else ())
    return out


def gn_act(x, stats, gamma, beta, count, act=ACT_LRELU, slope=0.1, transpose_out=False, slope_dev=None):
    """slope_dev: optional one-element device tensor (a PReLU weight) the kernel reads instead of the scalar `slope`."""
    b, n, c = x.shape
    out = torch.empty((b, c, n) if transpose_out else (b, n, c), dtype=torch.float32, device=x.device)
    _count(lib().pvraft_gn_act_fwd(_p(x), _p(stats, torch.float64), _p(gamma), _p(beta), float(count), act, float(slope),
                                   b, n, c, int(transpose_out), _p(out), _p(slope_dev), _stream()), 'gn_act')
    return out


def transpose(x):
    """[B,R,C] -> [B,C,R] contiguous."""
    b, r, c = x.shape
    out = torch.empty(b, c, r, dtype=torch.float32, device=x.device)
    _count(lib().pvraft_transpose_fwd(_p(x), b, r, c, _p(out), _stream()), 'transpose')
    return out


def corr_feature(args):
    _count(lib().pvraft_corr_feature_fwd(C.byref(args), _stream()), 'corr_feature')


def knn_branch(args):
    _count(lib().pvraft_knn_branch_fwd(C.byref(args), _stream()), 'knn_branch')


def gru(args):
    _count(lib().pvraft_gru_fwd(C.byref(args), _stream()), 'gru')


def flow_out(args):
    _count(lib().pvraft_flow_out_fwd(C.byref(args), _stream()), 'flow_out')


def setconv_edge(fc1p, nbr, edge_feats, w_fc1, cin, stats, ymax=None, ymin=None, order=None):
    b, n, c = fc1p.shape
    if ymax is None:
        ymax = torch.empty_like(fc1p)
    if ymin is None:
        ymin = torch.empty_like(fc1p)
    _count(lib().pvraft_setconv_edge_fwd(_p(fc1p), _p(nbr, torch.int32), _p(edge_feats), _p(w_fc1), cin, b, n, c, _p(ymax),
                                         _p(ymin), _p(stats, torch.float64), _p(order, torch.int32), _stream()), 'setconv_edge')
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


# ---- gradient contract (include/pvraft_b200.h, "Gradient contract"): thin wrappers used by pvraft_b200/train.py ------------
def linear_wgrad(x, dy, dw, db=None):
    """dw [cout,cin] += dy^T x, db [cout] += column sums of dy (both zeroed by the caller); x [B,R,cin], dy [B,R,cout]."""
    rows = x.shape[0] * x.shape[1]
    _count(lib().pvraft_linear_wgrad(_p(x), _p(dy), rows, x.shape[-1], dy.shape[-1], _p(dw), dw.shape[-1], _p(db), _stream()),
           'linear_wgrad')


def linear_bwd_small(x, dy, w, dw, db=None, want_dx=False):
    """cin <= 4, cout in {16,32,48,64,96,128}: dw += dy^T x, db += column sums, and dx = dy w (returned, or None) in one pass over dy."""
    rows = x.shape[0] * x.shape[1]
    dx = torch.empty_like(x) if want_dx else None
    _count(lib().pvraft_linear_bwd_small(_p(x), _p(dy), _p(w), rows, x.shape[-1], dy.shape[-1], w.shape[-1], _p(dw), dw.shape[-1], _p(db),
                                         _p(dx), _stream()), 'linear_bwd_small')
    return dx


def gn_act_maxk(x, stats, gamma, beta, count, act, slope, slope_dev=None):
    """x [B, pts*32, C] -> (max over each point's 32 rows of act(GN(x)) [B,pts,C], arg uint8 [B,pts,C]), one pass."""
    b, rows, c = x.shape
    y = torch.empty(b, rows // 32, c, dtype=torch.float32, device=x.device)
    arg = torch.empty(b, rows // 32, c, dtype=torch.uint8, device=x.device)
    _count(lib().pvraft_gn_act_maxk_fwd(_p(x), _p(stats, torch.float64), _p(gamma), _p(beta), float(count), act, float(slope), b,
                                        rows // 32, c, _p(y), _p(arg, torch.uint8), _p(slope_dev), _stream()), 'gn_act_maxk')
    return y, arg


def gn_act_bwd(x, dy, stats, gamma, beta, count, act, slope, want_dslope=False, slope_dev=None, arg=None):
    """-> (dx, dgamma [C] f32, dbeta [C] f32, dslope [1] f32 or None)."""
    b, rows, c = x.shape
    dev = x.device
    scratch = torch.zeros(b * 16 + 2 * c + 1, dtype=torch.float64, device=dev)   # gsum | dgamma | dbeta | dslope
    gsum, dgamma, dbeta, dslope = scratch[:b * 16], scratch[b * 16:b * 16 + c], scratch[b * 16 + c:b * 16 + 2 * c], scratch[-1:]
    dx = torch.empty_like(x)
    _count(lib().pvraft_gn_act_bwd(_p(x), _p(dy), _p(stats, torch.float64), _p(gamma), _p(beta), float(count), act, float(slope), b, rows,
                                   c, gsum.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), dslope.data_ptr() if want_dslope else None,
                                   _p(dx), _p(slope_dev), _p(arg, torch.uint8), _stream()), 'gn_act_bwd')
    return dx, dgamma.float(), dbeta.float(), (dslope.float() if want_dslope else None)


def edge_fwd(p, nbr, e, stats=None):
    """e [B,N*32,C] <- p[nbr] - p[centre] + e in place; stats [B,8,2] f64 accumulated."""
    b, n, c = p.shape
    _count(lib().pvraft_edge_fwd(_p(p), _p(nbr, torch.int32), _p(e), b, n, c, _p(stats, torch.float64), _stream()), 'edge_fwd')
    return e


def edge_bwd(dt, nbr, dp):
    b, n, c = dp.shape
    _count(lib().pvraft_edge_bwd(_p(dt), _p(nbr, torch.int32), b, n, c, _p(dp), _stream()), 'edge_bwd')
    return dp


def maxk_fwd(x, pts, c):
    y = torch.empty(pts, c, dtype=torch.float32, device=x.device)
    arg = torch.empty(pts, c, dtype=torch.uint8, device=x.device)
    _count(lib().pvraft_maxk_fwd(_p(x), pts, c, _p(y), _p(arg, torch.uint8), _stream()), 'maxk_fwd')
    return y, arg


def maxk_bwd(dy, arg, pts, c):
    dx = torch.empty(pts, KNN, c, dtype=torch.float32, device=dy.device)
    _count(lib().pvraft_maxk_bwd(_p(dy), _p(arg, torch.uint8), pts, c, _p(dx), _stream()), 'maxk_bwd')
    return dx


def corr_lookup_bwd(corr_idx, xyz2_pad, coords, slots, g_vox, g_sel, levels, base_scale):
    b, n, k = corr_idx.shape
    d_corr = torch.empty(b, n, k, dtype=torch.float32, device=corr_idx.device)
    _count(lib().pvraft_corr_lookup_bwd(_p(corr_idx, torch.int32), _p(xyz2_pad), _p(coords), _p(slots, torch.int32), _p(g_vox),
                                        g_vox.shape[-1], _p(g_sel), b, n, k, levels, float(base_scale), _p(d_corr), _stream()),
           'corr_lookup_bwd')
    return d_corr


def corr_init_bwd(g, idx, fmap1, fmap2):
    b, n, c = fmap1.shape
    d1 = torch.empty_like(fmap1)
    d2 = torch.zeros_like(fmap2)
    _count(lib().pvraft_corr_init_bwd(_p(g), _p(idx, torch.int32), _p(fmap1), _p(fmap2), b, n, c, g.shape[-1], _p(d1), _p(d2), _stream()),
           'corr_init_bwd')
    return d1, d2


_SM_COUNT = {}


def _sm_count():
    dev = torch.cuda.current_device()
    if dev not in _SM_COUNT:
        _SM_COUNT[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SM_COUNT[dev]


def device_info():
    sm, smem = C.c_int(0), C.c_int(0)
    check(lib().pvraft_device_info(C.byref(sm), C.byref(smem)), 'device_info')
    return sm.value, smem.value
