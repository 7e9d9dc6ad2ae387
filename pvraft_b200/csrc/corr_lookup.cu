// Fused point-voxel correlation lookup (index + reduce part) -- the HBM-bound headline kernel.
//
// Replaces CorrBlock.get_voxel_feature up to out_conv (reference model/corr.py:47-71) and
// CorrBlock.get_knn_feature up to knn_conv (model/corr.py:75-91) with ONE pass over the K
// candidates of every point:  a warp owns a point, streams its K (corr, index) pairs with 128-bit
// loads (8 B per candidate -- the candidate xyz is NOT materialised as in the reference's
// [B,N,K,3] tensor but gathered from a per-sample float4 table staged in shared memory),
// and produces
//   * the 27-cell x `levels` voxel means (lanes 0..26 each own one cell; candidates that pass the
//     coarsest-level cube test are compacted, in ascending candidate order, into a per-warp list and
//     summed sequentially -> the sums are bit-identical to a sequential scatter_add),
//   * the 32 nearest candidates (exact threshold found by a bitwise bisection on the fp32 distance
//     bits, warp-wide population counts, no sorting),
//   * double-precision first/second moments of the kNN 4-vectors, from which the consumer derives
//     the GroupNorm statistics of knn_conv's output without materialising its [B,64,N,32] tensor.
//
// Arithmetic that decides indices is bit-faithful to the reference's fp32 op sequence: separate
// rn subtract / multiply / add (no FMA contraction), true IEEE division, round-half-even.
#include "common.cuh"

namespace pvraft {

constexpr int kLookupThreads = 512;

struct LookupParams {
    const float* corr_val;
    const int32_t* corr_idx;
    const float4* tab;   // [B,N] (x,y,z,0)
    const float* coords; // [B,N,3]
    float* vox;          // [B,N,levels*27]
    float4* knn_sel;     // [B,N,32]
    int32_t* knn_slot;   // [B,N,32] or null
    double* moments;     // [B,16] or null
    int8_t* dbg_cube;    // [B,N,K,levels] or null
    int B, N, K, levels;
    float r[4];          // cell edge per level
    float inv_r[4];      // exact reciprocal when r is a power of two
    int warps;           // warps per block actually carved in shared memory
};

template <bool POW2>
__device__ __forceinline__ float div_r(float d, float r, float inv_r) {
    return POW2 ? __fmul_rn(d, inv_r) : __fdiv_rn(d, r);
}

// cell id in [0,27) of offset (dx,dy,dz) at cell edge r, or 0xFF when outside the 3x3x3 cube
// (model/corr.py:54-57: round((xyz - coords) / r), |.| <= 1 on all axes, (qx+1)*9+(qy+1)*3+(qz+1))
template <bool POW2>
__device__ __forceinline__ unsigned cell_code(float dx, float dy, float dz, float r, float inv_r) {
    const float qx = rintf(div_r<POW2>(dx, r, inv_r));
    const float qy = rintf(div_r<POW2>(dy, r, inv_r));
    const float qz = rintf(div_r<POW2>(dz, r, inv_r));
    const bool ok = (fabsf(qx) <= 1.f) && (fabsf(qy) <= 1.f) && (fabsf(qz) <= 1.f);
    const int cell = (int)(qx + 1.f) * 9 + (int)(qy + 1.f) * 3 + (int)(qz + 1.f);
    return ok ? (unsigned)cell : 0xFFu;
}

template <int KPL, bool POW2, bool SMEM_TAB>
__global__ void __launch_bounds__(kLookupThreads, 1) k_corr_lookup(const LookupParams p) {
    constexpr int VEC = KPL >= 4 ? 4 : KPL;   // consecutive candidates per lane per load
    constexpr int NJ = KPL / VEC;             // loads per lane
    constexpr int K = KPL * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_tab = reinterpret_cast<float4*>(smem_raw);
    const size_t tab_bytes = SMEM_TAB ? (size_t)p.N * sizeof(float4) : 0;
    // per-warp scratch: compacted (code,val) list [K] + kNN slot list [32]
    const int w = warp_id(), lane = lane_id();
    uint2* s_list = reinterpret_cast<uint2*>(smem_raw + tab_bytes) + (size_t)w * K;
    int* s_slots = reinterpret_cast<int*>(smem_raw + tab_bytes + (size_t)p.warps * K * sizeof(uint2)) + w * 32;
    const bool active_warp = w < p.warps;

    const long long total = (long long)p.B * p.N;
    long long pt_begin, pt_end;
    split_range(total, gridDim.x, blockIdx.x, pt_begin, pt_end);
    const unsigned lt_mask = (1u << lane) - 1u;
    const int L = p.levels;
    const float rc = p.r[L - 1], inv_rc = p.inv_r[L - 1];   // coarsest level

    long long seg = pt_begin;
    while (seg < pt_end) {
        const int b = (int)(seg / p.N);
        long long seg_end = (long long)(b + 1) * p.N;
        if (seg_end > pt_end) seg_end = pt_end;
        const float4* tab_g = p.tab + (size_t)b * p.N;
        if (SMEM_TAB) {
            __syncthreads();   // previous segment's readers are done
            for (int i = threadIdx.x; i < p.N; i += blockDim.x) s_tab[i] = tab_g[i];
            __syncthreads();
        }
        double mom[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) mom[i] = 0.0;

        if (active_warp) {
            for (long long pt = seg + w; pt < seg_end; pt += p.warps) {
                const float cx = __ldg(p.coords + pt * 3 + 0);
                const float cy = __ldg(p.coords + pt * 3 + 1);
                const float cz = __ldg(p.coords + pt * 3 + 2);
                const float* rv = p.corr_val + pt * K;
                const int32_t* ri = p.corr_idx + pt * K;

                // ---- stream the row: slot(j,s) = j*32*VEC + lane*VEC + s -------------------------
                float cv[KPL];
                int ci[KPL];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (VEC == 4) {
                        const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(rv + j * 128) + lane);
                        const int4 c = ld_stream_i4(reinterpret_cast<const int4*>(ri + j * 128) + lane);
                        cv[j * 4 + 0] = a.x; cv[j * 4 + 1] = a.y; cv[j * 4 + 2] = a.z; cv[j * 4 + 3] = a.w;
                        ci[j * 4 + 0] = c.x; ci[j * 4 + 1] = c.y; ci[j * 4 + 2] = c.z; ci[j * 4 + 3] = c.w;
                    } else {
#pragma unroll
                        for (int s = 0; s < VEC; ++s) {
                            cv[j * VEC + s] = __ldg(rv + j * 32 * VEC + lane * VEC + s);
                            ci[j * VEC + s] = __ldg(ri + j * 32 * VEC + lane * VEC + s);
                        }
                    }
                }

                // ---- distances + coarsest-level cube test + ordered compaction ---------------------
                unsigned dist[KPL];   // fp32 bits of the (non-negative) squared distance
                int list_n = 0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    unsigned vmask[VEC];
                    unsigned any = 0;
#pragma unroll
                    for (int s = 0; s < VEC; ++s) {
                        const int e = j * VEC + s;
                        const float4 q = SMEM_TAB ? s_tab[ci[e]] : __ldg(tab_g + ci[e]);
                        const float dx = __fsub_rn(q.x, cx), dy = __fsub_rn(q.y, cy), dz = __fsub_rn(q.z, cz);
                        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                        dist[e] = __float_as_uint(d2);
                        const float amax = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
                        // |round(d/r)| <= 1 on every axis  <=>  max|d|/r < 1.5 (round-half-even sends 1.5 to 2)
                        const bool v = div_r<POW2>(amax, rc, inv_rc) < 1.5f;
                        vmask[s] = __ballot_sync(kFull, v);
                        any |= vmask[s];
                        if (p.dbg_cube) {
                            const int slot = j * 32 * VEC + lane * VEC + s;
                            for (int l = 0; l < L; ++l) {
                                const unsigned c = cell_code<POW2>(dx, dy, dz, p.r[l], p.inv_r[l]);
                                p.dbg_cube[((size_t)pt * K + slot) * L + l] = c == 0xFFu ? (int8_t)-1 : (int8_t)c;
                            }
                        }
                    }
                    if (any) {   // warp-uniform
                        int pos = list_n;
#pragma unroll
                        for (int s = 0; s < VEC; ++s) pos += __popc(vmask[s] & lt_mask);
#pragma unroll
                        for (int s = 0; s < VEC; ++s) {
                            if ((vmask[s] >> lane) & 1u) {
                                const int e = j * VEC + s;
                                const float4 q = SMEM_TAB ? s_tab[ci[e]] : __ldg(tab_g + ci[e]);
                                const float dx = __fsub_rn(q.x, cx), dy = __fsub_rn(q.y, cy), dz = __fsub_rn(q.z, cz);
                                unsigned code = 0xFFFFFFFFu;
                                for (int l = 0; l < L; ++l) {
                                    const unsigned c = cell_code<POW2>(dx, dy, dz, p.r[l], p.inv_r[l]);
                                    code = (code & ~(0xFFu << (8 * l))) | (c << (8 * l));
                                }
                                s_list[pos] = make_uint2(code, __float_as_uint(cv[e]));
                                ++pos;
                            }
                            list_n += __popc(vmask[s]);
                        }
                    }
                }
                __syncwarp();

                // ---- voxel means: lane c (<27) owns cell c of every level; sequential sums --------
                {
                    float sum[4] = {0.f, 0.f, 0.f, 0.f};
                    int cnt[4] = {0, 0, 0, 0};
                    for (int i = 0; i < list_n; ++i) {
                        const uint2 e = s_list[i];
                        const float val = __uint_as_float(e.y);
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            if (((e.x >> (8 * l)) & 0xFFu) == (unsigned)lane) {
                                sum[l] = __fadd_rn(sum[l], val);
                                cnt[l] += 1;
                            }
                        }
                    }
                    if (lane < 27) {
                        float* vo = p.vox + pt * (L * 27);
                        for (int l = 0; l < L; ++l) {
                            const float c = (float)(cnt[l] < 1 ? 1 : cnt[l]);   // clamp(count, 1, N), corr.py:65
                            vo[l * 27 + lane] = __fdiv_rn(sum[l], c);
                        }
                    }
                }

                // ---- kNN: smallest threshold T with count(d <= T) >= 32, by bisection on the bits ----
                unsigned lmin = dist[0];
#pragma unroll
                for (int e = 1; e < KPL; ++e) lmin = min(lmin, dist[e]);
                unsigned hi = __reduce_max_sync(kFull, lmin);   // 32 distinct candidates are <= hi
                unsigned lo = __reduce_min_sync(kFull, lmin);
                unsigned T = hi;
                while (lo < hi) {
                    const unsigned mid = lo + ((hi - lo) >> 1);
                    int c = 0;
#pragma unroll
                    for (int e = 0; e < KPL; ++e) c += dist[e] <= mid ? 1 : 0;
                    c = __reduce_add_sync(kFull, c);
                    if (c == PVRAFT_KNN) { hi = mid; lo = mid; break; }
                    if (c > PVRAFT_KNN) hi = mid; else lo = mid + 1;
                }
                T = hi;
                // slots with d < T first, then ties d == T in ascending slot order until 32 are taken
                int n_lt = 0;
#pragma unroll
                for (int e = 0; e < KPL; ++e) n_lt += dist[e] < T ? 1 : 0;
                n_lt = __reduce_add_sync(kFull, n_lt);
                int base_lt = 0, base_eq = n_lt;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    unsigned mlt[VEC], meq[VEC];
#pragma unroll
                    for (int s = 0; s < VEC; ++s) {
                        mlt[s] = __ballot_sync(kFull, dist[j * VEC + s] < T);
                        meq[s] = __ballot_sync(kFull, dist[j * VEC + s] == T);
                    }
                    int plt = base_lt, peq = base_eq;
#pragma unroll
                    for (int s = 0; s < VEC; ++s) {
                        plt += __popc(mlt[s] & lt_mask);
                        peq += __popc(meq[s] & lt_mask);
                    }
#pragma unroll
                    for (int s = 0; s < VEC; ++s) {
                        const int slot = j * 32 * VEC + lane * VEC + s;
                        if ((mlt[s] >> lane) & 1u) s_slots[plt++] = slot;
                        if ((meq[s] >> lane) & 1u) {
                            if (peq < PVRAFT_KNN) s_slots[peq] = slot;
                            ++peq;
                        }
                        base_lt += __popc(mlt[s]);
                        base_eq += __popc(meq[s]);
                    }
                }
                __syncwarp();
                {
                    const int slot = s_slots[lane];
                    const float c = __ldg(rv + slot);
                    const int id = __ldg(ri + slot);
                    const float4 q = SMEM_TAB ? s_tab[id] : __ldg(tab_g + id);
                    const float dx = __fsub_rn(q.x, cx), dy = __fsub_rn(q.y, cy), dz = __fsub_rn(q.z, cz);
                    p.knn_sel[pt * 32 + lane] = make_float4(c, dx, dy, dz);
                    if (p.knn_slot) p.knn_slot[pt * 32 + lane] = slot;
                    const double f0 = c, f1 = dx, f2 = dy, f3 = dz;
                    mom[0] += f0; mom[1] += f1; mom[2] += f2; mom[3] += f3;
                    mom[4] += f0 * f0; mom[5] += f0 * f1; mom[6] += f0 * f2; mom[7] += f0 * f3;
                    mom[8] += f1 * f1; mom[9] += f1 * f2; mom[10] += f1 * f3;
                    mom[11] += f2 * f2; mom[12] += f2 * f3; mom[13] += f3 * f3;
                }
                __syncwarp();
            }
            if (p.moments) {
#pragma unroll
                for (int i = 0; i < 14; ++i) {
                    const double s = warp_sum(mom[i]);
                    if (lane == 0 && s != 0.0) atomicAdd(p.moments + (size_t)b * PVRAFT_MOMENTS + i, s);
                }
                if (lane == 0 && w == 0)
                    atomicAdd(p.moments + (size_t)b * PVRAFT_MOMENTS + 14, (double)(seg_end - seg) * 32.0);
            }
        }
        seg = seg_end;
    }
}

__global__ void k_pad_xyz(const float* __restrict__ xyz, long long n, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], 0.f);
}

static bool is_pow2f(float r) {
    int e;
    return r > 0.f && frexpf(r, &e) == 0.5f;
}

template <int KPL, bool POW2>
static int launch_lookup(LookupParams& p, cudaStream_t st) {
    const int K = KPL * 32;
    const size_t per_warp = (size_t)K * sizeof(uint2) + 32 * sizeof(int);
    const size_t tab = (size_t)p.N * sizeof(float4);
    bool smem_tab = tab + 4 * per_warp <= (size_t)kSmemBudget;
    size_t avail = (size_t)kSmemBudget - (smem_tab ? tab : 0);
    int warps = (int)(avail / per_warp);
    if (warps > kLookupThreads / 32) warps = kLookupThreads / 32;
    if (warps < 1) return fail(PVRAFT_ERR_SMEM, "corr_lookup: K=%d does not fit shared memory", K);
    p.warps = warps;
    const size_t smem = (smem_tab ? tab : 0) + warps * per_warp;
    const long long total = (long long)p.B * p.N;
    int grid = sm_count();
    const long long min_pts = warps;   // no point in more blocks than (points / warps)
    if ((long long)grid * min_pts > total) grid = (int)((total + min_pts - 1) / min_pts);
    if (grid < 1) grid = 1;
    int rc;
    if (smem_tab) {
        auto k = k_corr_lookup<KPL, POW2, true>;
        if ((rc = opt_in_smem(k, smem))) return rc;
        k<<<grid, kLookupThreads, smem, st>>>(p);
    } else {
        auto k = k_corr_lookup<KPL, POW2, false>;
        if ((rc = opt_in_smem(k, smem))) return rc;
        k<<<grid, kLookupThreads, smem, st>>>(p);
    }
    return check_launch("corr_lookup");
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int pvraft_pad_xyz(const float* xyz, int64_t n, float* xyz4, void* stream) {
    if (!xyz || !xyz4 || n <= 0) return fail(PVRAFT_ERR_BAD_ARG, "pad_xyz: bad argument");
    k_pad_xyz<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz, n, reinterpret_cast<float4*>(xyz4));
    return check_launch("pad_xyz");
}

extern "C" int pvraft_corr_lookup_fwd(const float* corr_val, const int32_t* corr_idx, const float* xyz2p,
                                      const float* coords, int B, int N, int K, int levels, float base_scale,
                                      float* vox, float* knn_sel, int32_t* knn_slot, double* moments,
                                      int8_t* dbg_cube, void* stream) {
    if (!corr_val || !corr_idx || !xyz2p || !coords || !vox || !knn_sel) return fail(PVRAFT_ERR_BAD_ARG, "corr_lookup: null pointer");
    if (B <= 0 || N <= 0) return fail(PVRAFT_ERR_BAD_ARG, "corr_lookup: B=%d N=%d", B, N);
    if (levels < 1 || levels > 4) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_lookup: levels=%d (1..4 supported)", levels);
    if (!(base_scale > 0.f)) return fail(PVRAFT_ERR_BAD_ARG, "corr_lookup: base_scale must be > 0");
    LookupParams p{};
    p.corr_val = corr_val; p.corr_idx = corr_idx; p.tab = reinterpret_cast<const float4*>(xyz2p); p.coords = coords;
    p.vox = vox; p.knn_sel = reinterpret_cast<float4*>(knn_sel); p.knn_slot = knn_slot; p.moments = moments;
    p.dbg_cube = dbg_cube; p.B = B; p.N = N; p.K = K; p.levels = levels;
    bool pow2 = true;
    for (int l = 0; l < 4; ++l) {
        // model/corr.py:53: r = base_scale * 2**i evaluated in double, then used as an fp32 divisor
        const float r = (float)((double)base_scale * (double)(1 << l));
        p.r[l] = r;
        p.inv_r[l] = 1.0f / r;
        if (l < levels && !is_pow2f(r)) pow2 = false;
    }
    cudaStream_t st = (cudaStream_t)stream;
#define PVRAFT_LOOKUP_CASE(KPL_)                                                            \
    case KPL_ * 32:                                                                         \
        return pow2 ? launch_lookup<KPL_, true>(p, st) : launch_lookup<KPL_, false>(p, st);
    switch (K) {
        PVRAFT_LOOKUP_CASE(1)
        PVRAFT_LOOKUP_CASE(2)
        PVRAFT_LOOKUP_CASE(4)
        PVRAFT_LOOKUP_CASE(8)
        PVRAFT_LOOKUP_CASE(16)
        PVRAFT_LOOKUP_CASE(32)
        default:
            return fail(PVRAFT_ERR_UNSUPPORTED, "corr_lookup: truncate_k=%d (supported: 32,64,128,256,512,1024)", K);
    }
#undef PVRAFT_LOOKUP_CASE
}
