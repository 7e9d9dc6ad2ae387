// Per-row top-K (sorted descending) of the dense correlation matrix: the truncation step of
// CorrBlock.init_module (reference model/corr.py:37-40, torch.topk(corr, k, dim=2, sorted=True)).
//
// One CTA per row.  The row is staged once in shared memory as order-preserving uint32 keys; the
// K-th largest key is found by a 4-pass (8 bits per pass) radix select on shared-memory histograms,
// the K survivors are compacted (ties at the threshold: lowest column first) and ordered with an
// in-shared-memory bitonic sort on (key desc, column asc).  Output: fp32 values + int32 columns --
// the 8-byte-per-candidate state streamed by the lookup kernel every RAFT iteration.
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

__global__ void __launch_bounds__(kTopkThreads) k_corr_topk(const float* __restrict__ corr, int M, int K, int KP /*pow2 >= K*/,
                                                            float* __restrict__ val, int32_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* s_key = reinterpret_cast<unsigned*>(smem_raw);          // [M]
    unsigned long long* s_sel = reinterpret_cast<unsigned long long*>(s_key + ((M + 1) & ~1));   // [KP]
    __shared__ int s_hist[256];
    __shared__ unsigned s_prefix, s_need;
    __shared__ int s_cnt_gt, s_cnt_eq;
    const size_t row = blockIdx.x;
    const float* src = corr + row * (size_t)M;
    const int tid = threadIdx.x;
    for (int i = tid; i < M; i += kTopkThreads) s_key[i] = f2key(__ldg(src + i));
    if (tid == 0) { s_prefix = 0u; s_need = (unsigned)K; }
    __syncthreads();
    // ---- radix select: after pass p the top (8*(p+1)) bits of the K-th largest key are known ------------
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        s_hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < M; i += kTopkThreads) {
            const unsigned k = s_key[i];
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 0xFF], 1);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned need = s_need;
            int bin = 255;
            for (; bin > 0; --bin) {
                const unsigned h = (unsigned)s_hist[bin];
                if (h >= need) break;
                need -= h;
            }
            s_need = need;               // rank of the target inside the chosen bin
            s_prefix = prefix | ((unsigned)bin << shift);
        }
        __syncthreads();
    }
    const unsigned T = s_prefix;          // exact K-th largest key
    const unsigned need_eq = s_need;      // how many keys == T belong to the top-K
    if (tid == 0) { s_cnt_gt = 0; s_cnt_eq = 0; }
    for (int i = tid; i < KP; i += kTopkThreads) s_sel[i] = 0ull;   // pads sort to the end (key 0 < any real key)
    __syncthreads();
    // ---- compaction: keys > T anywhere; keys == T in ascending column order (deterministic ties) --------
    // ordered handling of ties: chunked scan so that lower columns claim their slot first
    const int n_gt_total = K - (int)need_eq;
    for (int base = 0; base < M; base += kTopkThreads) {
        const int i = base + tid;
        const unsigned k = i < M ? s_key[i] : 0u;
        const bool gt = i < M && k > T;
        const bool eq = i < M && k == T;
        if (gt) {
            const int pos = atomicAdd(&s_cnt_gt, 1);
            s_sel[pos] = ((unsigned long long)k << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
        }
        // ties: warp-ordered ballot inside the chunk, chunks in order
        const unsigned m = __ballot_sync(kFull, eq);
        __shared__ int s_warp_eq[kTopkThreads / 32];
        if ((tid & 31) == 0) s_warp_eq[tid >> 5] = __popc(m);
        __syncthreads();
        if (eq) {
            int before = s_cnt_eq;
            for (int w = 0; w < (tid >> 5); ++w) before += s_warp_eq[w];
            before += __popc(m & ((1u << (tid & 31)) - 1u));
            if (before < (int)need_eq)
                s_sel[n_gt_total + before] = ((unsigned long long)k << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kTopkThreads / 32; ++w) tot += s_warp_eq[w];
            s_cnt_eq += tot;
        }
        __syncthreads();
    }
    // ---- bitonic sort, descending on (key, ~column) => value desc, column asc -----------------------------
    for (int size = 2; size <= KP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < KP / 2; t += kTopkThreads) {
                const int lo = (t / stride) * (stride << 1) + (t % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = s_sel[lo], b = s_sel[hi];
                if ((a < b) == desc) { s_sel[lo] = b; s_sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += kTopkThreads) {
        const unsigned long long e = s_sel[i];
        val[row * K + i] = key2f((unsigned)(e >> 32));
        idx[row * K + i] = (int32_t)(0xFFFFFFFFu - (unsigned)(e & 0xFFFFFFFFull));
    }
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int pvraft_corr_topk_fwd(const float* corr, int B, int N, int M, int K, float* val, int32_t* idx, void* stream) {
    if (!corr || !val || !idx) return fail(PVRAFT_ERR_BAD_ARG, "corr_topk: null pointer");
    if (B <= 0 || N <= 0 || M <= 0) return fail(PVRAFT_ERR_BAD_ARG, "corr_topk: bad shape");
    if (K < 1 || K > M || K > 1024) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_topk: K=%d with M=%d (need 1 <= K <= min(M,1024))", K, M);
    int KP = 1;
    while (KP < K) KP <<= 1;
    const size_t smem = (size_t)((M + 1) & ~1) * 4 + (size_t)KP * 8;
    int rc;
    if ((rc = opt_in_smem(k_corr_topk, smem))) return rc;
    const long long rows = (long long)B * N;
    if (rows > 0x7fffffffLL) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_topk: too many rows");
    k_corr_topk<<<(unsigned)rows, kTopkThreads, smem, (cudaStream_t)stream>>>(corr, M, K, KP, val, idx);
    return check_launch("corr_topk");
}
