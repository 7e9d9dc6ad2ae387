// Per-row top-K of the dense correlation matrix: the truncation step of CorrBlock.init_module
// (reference model/corr.py:37-40, torch.topk(corr, k, dim=2, sorted=True)).
//
// One CTA per row.  The row is staged once in shared memory as order-preserving uint32 keys (one pad word
// per 32 so that a thread can later walk its own 32-key segment without bank conflicts); the K-th largest
// key is found by a 4-pass (8 bits per pass) radix select on shared-memory histograms; then every thread
// counts the survivors of its segment, one block scan gives it its output offset, and the K survivors are
// written in ASCENDING COLUMN order (ties at the threshold: lowest columns win).  The reference sorts the K
// values descending; nothing downstream depends on that order (pvraft_corr_reorder rearranges every row
// anyway), so the sort is not done here -- CorrBlock.truncated_corr sorts on demand for API parity.
#include "common.cuh"

namespace pvraft {

constexpr int kTopkThreads = 256;

__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // larger float -> larger key
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ int padded(int i) { return i + (i >> 5); }

__global__ void __launch_bounds__(kTopkThreads) k_corr_topk(const float* __restrict__ corr, int M, int K,
                                                            float* __restrict__ val, int32_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* s_key = reinterpret_cast<unsigned*>(smem_raw);   // [padded(M)]
    __shared__ int s_hist[256];
    __shared__ unsigned s_prefix, s_need;
    __shared__ unsigned s_warp[kTopkThreads / 32];
    const size_t row = blockIdx.x;
    const float* src = corr + row * (size_t)M;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < M; i += kTopkThreads) s_key[padded(i)] = f2key(__ldg(src + i));
    if (tid == 0) { s_prefix = 0u; s_need = (unsigned)K; }
    // ---- radix select: after pass p the top 8*(p+1) bits of the K-th largest key are known --------------
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        s_hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < M; i += kTopkThreads) {
            const unsigned k = s_key[padded(i)];
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 0xFF], 1);
        }
        __syncthreads();
        if (w == 0) {
            // lane l owns bins 8l .. 8l+7; suffix sums locate the bin holding the `need`-th largest key
            int h[8], mine = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q] = s_hist[lane * 8 + q]; mine += h[q]; }
            int above = mine;   // inclusive suffix sum over lanes >= l
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_down_sync(kFull, above, o);
                if (lane + o < 32) above += a;
            }
            const int need = (int)s_need;
            const int higher = above - mine;   // keys in bins owned by higher lanes
            const bool here = higher < need && above >= need;
            if (here) {
                int acc = higher, bin = 7;
                for (; bin > 0; --bin) {
                    if (acc + h[bin] >= need) break;
                    acc += h[bin];
                }
                s_need = (unsigned)(need - acc);          // rank of the target inside the chosen bin
                s_prefix = prefix | ((unsigned)(lane * 8 + bin) << shift);
            }
        }
        __syncthreads();
    }
    const unsigned T = s_prefix;       // exact K-th largest key
    const int need_eq = (int)s_need;   // how many keys == T belong to the top-K (lowest columns first)
    // ---- ordered compaction: thread t owns columns [t*seg, (t+1)*seg) ----------------------------------------
    const int seg = (M + kTopkThreads - 1) / kTopkThreads;
    const int c0 = tid * seg, c1 = min(M, c0 + seg);
    unsigned n_gt = 0, n_eq = 0;
    for (int i = c0; i < c1; ++i) {
        const unsigned k = s_key[padded(i)];
        n_gt += k > T ? 1u : 0u;
        n_eq += k == T ? 1u : 0u;
    }
    // block exclusive scan of (n_gt | n_eq << 16): M <= 49152 keeps both fields below 65536
    unsigned packed = n_gt | (n_eq << 16), incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned a = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += a;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int q = 0; q < w; ++q) base += s_warp[q];
    const unsigned excl = base + incl - packed;
    const int gt_before = (int)(excl & 0xFFFFu), eq_before = (int)(excl >> 16);
    // output position = (# kept elements in lower columns) = gt_before + min(eq_before, need_eq)
    int pos = gt_before + min(eq_before, need_eq);
    int eq_seen = eq_before;
    for (int i = c0; i < c1; ++i) {
        const unsigned k = s_key[padded(i)];
        bool keep = k > T;
        if (k == T) { keep = eq_seen < need_eq; ++eq_seen; }
        if (keep) {
            val[row * K + pos] = key2f(k);
            idx[row * K + pos] = i;
            ++pos;
        }
    }
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int pvraft_corr_topk_fwd(const float* corr, int B, int N, int M, int K, float* val, int32_t* idx, void* stream) {
    if (!corr || !val || !idx) return fail(PVRAFT_ERR_BAD_ARG, "corr_topk: null pointer");
    if (B <= 0 || N <= 0 || M <= 0) return fail(PVRAFT_ERR_BAD_ARG, "corr_topk: bad shape");
    if (K < 1 || K > M || K > 1024) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_topk: K=%d with M=%d (need 1 <= K <= min(M,1024))", K, M);
    if (M > 49152) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_topk: M=%d columns (max 49152)", M);
    const size_t smem = (size_t)(M + (M >> 5) + 4) * 4;
    int rc;
    if ((rc = opt_in_smem(k_corr_topk, smem))) return rc;
    const long long rows = (long long)B * N;
    if (rows > 0x7fffffffLL) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_topk: too many rows");
    k_corr_topk<<<(unsigned)rows, kTopkThreads, smem, (cudaStream_t)stream>>>(corr, M, K, val, idx);
    return check_launch("corr_topk");
}
