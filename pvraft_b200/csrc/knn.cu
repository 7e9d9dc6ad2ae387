// Brute-force k-nearest-neighbour search: backs knn_point (reference model/pointconv.py:28-39) and the
// adjacency of Graph.construct_graph (reference model/flot/graph.py:53-60, which sorts a full N x N
// distance matrix to keep 32 columns).
//
// One warp per query.  The candidate cloud is staged through shared memory in tiles; lane l scores
// candidate (tile_base + 32*i + l).  The warp keeps the current k best as one (distance, id) pair per
// lane plus the running k-th distance tau; a candidate enters only if it beats tau (rare after the
// first few tiles), replacing the current worst.  Distances reproduce the reference's expanded form
// bit-for-bit on the CPU oracle: |q|^2 and |x|^2 as (x*x+y*y)+z*z, q.x as fma(z,z',fma(y,y',x*x')).
#include "common.cuh"

namespace pvraft {

constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 2048;   // candidates per shared-memory tile (x,y,z,|x|^2 -> 32 KB)

__device__ __forceinline__ float sqnorm(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// (distance, id) ordering: smaller distance first, ties -> smaller id first
__device__ __forceinline__ bool worse(float d1, int i1, float d2, int i2) { return d1 > d2 || (d1 == d2 && i1 > i2); }

__global__ void __launch_bounds__(kKnnThreads) k_knn(const float* __restrict__ xyz, const float* __restrict__ query, int N, int S,
                                                     int k, int mode, int32_t* __restrict__ out, float* __restrict__ rel) {
    __shared__ float4 s_pts[kKnnTile];
    const int b = blockIdx.y;
    const int lane = lane_id(), w = warp_id();
    const int q = blockIdx.x * (kKnnThreads / 32) + w;
    const bool live = q < S;
    const float* X = xyz + (size_t)b * N * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float* Q = query + ((size_t)b * S + q) * 3;
        qx = __ldg(Q); qy = __ldg(Q + 1); qz = __ldg(Q + 2);
    }
    const float qn = sqnorm(qx, qy, qz);
    // lane l holds the l-th member of the current best set (unordered); lanes >= k hold -inf sentinels so
    // that they are never the "worst".
    float bd = lane < k ? INFINITY : -INFINITY;
    int bi = lane < k ? 0x7fffffff - lane : -1;   // distinct sentinels: exactly one lane is "the worst"
    float tau = INFINITY;   // current worst (largest) of the k kept
    int tau_i = 0x7fffffff;

    for (int base = 0; base < N; base += kKnnTile) {
        const int cnt = min(kKnnTile, N - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += kKnnThreads) {
            const float x = __ldg(X + (size_t)(base + i) * 3), y = __ldg(X + (size_t)(base + i) * 3 + 1), z = __ldg(X + (size_t)(base + i) * 3 + 2);
            s_pts[i] = make_float4(x, y, z, sqnorm(x, y, z));
        }
        __syncthreads();
        if (!live) continue;
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            const int i = i0 + lane;
            float d = INFINITY;
            const int id = base + i;
            if (i < cnt) {
                const float4 p = s_pts[i];
                const float dot = __fmaf_rn(qz, p.z, __fmaf_rn(qy, p.y, __fmul_rn(qx, p.x)));
                if (mode == 0) d = __fsub_rn(__fadd_rn(qn, p.w), __fmul_rn(2.f, dot));              // graph.py:53-57
                else d = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), qn), p.w);                        // pointconv.py:21-24
            }
            unsigned cand = __ballot_sync(kFull, i < cnt && !worse(d, id, tau, tau_i));
            while (cand) {
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const float cd = __shfl_sync(kFull, d, src);
                const int cid = __shfl_sync(kFull, id, src);
                if (worse(cd, cid, tau, tau_i)) continue;   // tau may have tightened since the ballot
                // replace the current worst member (unique: (distance,id) pairs are distinct)
                if (bd == tau && bi == tau_i) { bd = cd; bi = cid; }
                // recompute the worst over the k kept lanes
                float md = bd; int mi = bi;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float od = __shfl_xor_sync(kFull, md, o);
                    const int oi = __shfl_xor_sync(kFull, mi, o);
                    if (worse(od, oi, md, mi)) { md = od; mi = oi; }
                }
                tau = md; tau_i = mi;
            }
        }
    }
    if (live && lane < k) {
        const size_t o = ((size_t)b * S + q) * k + lane;
        out[o] = bi;
        if (rel) {   // edge feature of Graph.construct_graph: neighbour - centre (graph.py:69-74)
            rel[o * 3 + 0] = __fsub_rn(__ldg(X + (size_t)bi * 3 + 0), qx);
            rel[o * 3 + 1] = __fsub_rn(__ldg(X + (size_t)bi * 3 + 1), qy);
            rel[o * 3 + 2] = __fsub_rn(__ldg(X + (size_t)bi * 3 + 2), qz);
        }
    }
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int pvraft_knn_fwd(const float* xyz, const float* query, int B, int N, int S, int k, int mode, int32_t* idx, float* rel, void* stream) {
    if (!xyz || !query || !idx) return fail(PVRAFT_ERR_BAD_ARG, "knn: null pointer");
    if (B <= 0 || N <= 0 || S <= 0) return fail(PVRAFT_ERR_BAD_ARG, "knn: bad shape");
    if (k < 1 || k > 32 || k > N) return fail(PVRAFT_ERR_UNSUPPORTED, "knn: k=%d (need 1 <= k <= min(32, N=%d))", k, N);
    if (mode != 0 && mode != 1) return fail(PVRAFT_ERR_BAD_ARG, "knn: mode=%d", mode);
    if (B > 65535) return fail(PVRAFT_ERR_UNSUPPORTED, "knn: B=%d", B);
    dim3 grid((S + kKnnThreads / 32 - 1) / (kKnnThreads / 32), B);
    k_knn<<<grid, kKnnThreads, 0, (cudaStream_t)stream>>>(xyz, query, N, S, k, mode, idx, rel);
    return check_launch("knn");
}
