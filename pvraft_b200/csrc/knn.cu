// k-nearest-neighbour search: backs knn_point (reference model/pointconv.py:28-39) and the adjacency of
// Graph.construct_graph (reference model/flot/graph.py:53-60, which sorts a full N x N distance matrix to keep
// 32 columns).
//
// Two kernels:
//   k_sort_x   one CTA per sample: bitonic sort of the candidates by x in shared memory -> float4 (x,y,z,|p|^2)
//              in sorted order + original ids + the largest |p|^2 of the sample
//   k_knn_sweep one warp per query: binary search for the query's x, then sweep outwards in both directions, 16
//              candidates per side per step.  The warp keeps the current k best as one (distance, id) pair per
//              lane plus the running k-th distance tau; a candidate enters only if it beats tau.  A side stops once
//              (x - qx)^2 exceeds tau by more than the rounding slack of the distance formula, so only a slab of
//              width ~2*sqrt(tau) is ever scored instead of the whole cloud.
// Distances reproduce the reference's expanded form bit-for-bit on the CPU oracle: |q|^2 and |x|^2 as
// (x*x+y*y)+z*z, q.x as fma(z,z',fma(y,y',x*x')); ranking is on (distance, original id), so the result does not
// depend on the sweep order.  Clouds too large for the shared-memory sort fall back to the brute-force kernel.
#include "common.cuh"

namespace pvraft {

constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 2048;      // candidates per shared-memory tile of the brute-force kernel
constexpr int kSortMaxN = 16384;    // (key, id) pairs of the in-smem sort: 8 B * 16384 = 128 KB

__device__ __forceinline__ float sqnorm(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// (distance, id) ordering: smaller distance first, ties -> smaller id first
__device__ __forceinline__ bool worse(float d1, int i1, float d2, int i2) { return d1 > d2 || (d1 == d2 && i1 > i2); }

__device__ __forceinline__ float ref_distance(int mode, float qx, float qy, float qz, float qn, const float4 p) {
    const float dot = __fmaf_rn(qz, p.z, __fmaf_rn(qy, p.y, __fmul_rn(qx, p.x)));
    if (mode == 0) return __fsub_rn(__fadd_rn(qn, p.w), __fmul_rn(2.f, dot));   // graph.py:53-57
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), qn), p.w);                  // pointconv.py:21-24
}

// insert the candidates flagged in `cand` (one per lane: d, id) into the warp-resident best set
__device__ __forceinline__ void insert_candidates(unsigned cand, float d, int id, float& bd, int& bi, float& tau, int& tau_i) {
    while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        const float cd = __shfl_sync(kFull, d, src);
        const int cid = __shfl_sync(kFull, id, src);
        if (worse(cd, cid, tau, tau_i)) continue;   // tau may have tightened since the ballot
        if (bd == tau && bi == tau_i) { bd = cd; bi = cid; }   // replace the current worst (pairs are distinct)
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

__device__ __forceinline__ void write_result(const float* __restrict__ X, float qx, float qy, float qz, int lane, int k, int bi,
                                             size_t out_row, int32_t* __restrict__ out, float* __restrict__ rel) {
    if (lane < k) {
        const size_t o = out_row * k + lane;
        out[o] = bi;
        if (rel) {   // edge feature of Graph.construct_graph: neighbour - centre (graph.py:69-74)
            rel[o * 3 + 0] = __fsub_rn(__ldg(X + (size_t)bi * 3 + 0), qx);
            rel[o * 3 + 1] = __fsub_rn(__ldg(X + (size_t)bi * 3 + 1), qy);
            rel[o * 3 + 2] = __fsub_rn(__ldg(X + (size_t)bi * 3 + 2), qz);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// brute force (any N)
// ---------------------------------------------------------------------------------------------------------
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
    float bd = lane < k ? INFINITY : -INFINITY;
    int bi = lane < k ? 0x7fffffff - lane : -1;   // distinct sentinels: exactly one lane is "the worst"
    float tau = INFINITY;
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
            if (i < cnt) d = ref_distance(mode, qx, qy, qz, qn, s_pts[i]);
            const unsigned cand = __ballot_sync(kFull, i < cnt && !worse(d, id, tau, tau_i));
            insert_candidates(cand, d, id, bd, bi, tau, tau_i);
        }
    }
    if (live) write_result(X, qx, qy, qz, lane, k, bi, (size_t)b * S + q, out, rel);
}

// ---------------------------------------------------------------------------------------------------------
// x-sorted sweep (N <= kSortMaxN)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_sort_x(const float* __restrict__ xyz, int N, int NP /*pow2 >= N*/,
                                                  float4* __restrict__ sorted, int32_t* __restrict__ ids, float* __restrict__ max_norm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw);   // (orderable x key << 32) | id
    __shared__ float s_max[32];
    const int b = blockIdx.x;
    const float* X = xyz + (size_t)b * N * 3;
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        unsigned long long e = 0xFFFFFFFFFFFFFFFFull;   // pads sort to the end
        if (i < N) {
            const unsigned u = __float_as_uint(__ldg(X + (size_t)i * 3));
            const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            e = ((unsigned long long)key << 32) | (unsigned)i;
        }
        s[i] = e;
    }
    __syncthreads();
    for (int size = 2; size <= NP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < NP / 2; t += blockDim.x) {
                const int lo = (t / stride) * (stride << 1) + (t % stride);
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const unsigned long long a = s[lo], c = s[hi];
                if ((a > c) == asc) { s[lo] = c; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    float mx = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int id = (int)(s[i] & 0xFFFFFFFFull);
        const float x = __ldg(X + (size_t)id * 3), y = __ldg(X + (size_t)id * 3 + 1), z = __ldg(X + (size_t)id * 3 + 2);
        const float n2 = sqnorm(x, y, z);
        sorted[(size_t)b * N + i] = make_float4(x, y, z, n2);
        ids[(size_t)b * N + i] = id;
        mx = fmaxf(mx, n2);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0.f;
        v = warp_max(v);
        if (threadIdx.x == 0) max_norm[b] = v;
    }
}

__global__ void __launch_bounds__(kKnnThreads) k_knn_sweep(const float* __restrict__ xyz, const float4* __restrict__ sorted,
                                                           const int32_t* __restrict__ ids, const float* __restrict__ max_norm,
                                                           const float* __restrict__ query, int N, int S, int k, int mode,
                                                           int32_t* __restrict__ out, float* __restrict__ rel) {
    const int b = blockIdx.y;
    const int lane = lane_id(), w = warp_id();
    const int q = blockIdx.x * (kKnnThreads / 32) + w;
    if (q >= S) return;
    const float* X = xyz + (size_t)b * N * 3;
    const float4* P = sorted + (size_t)b * N;
    const int32_t* I = ids + (size_t)b * N;
    const float* Q = query + ((size_t)b * S + q) * 3;
    const float qx = __ldg(Q), qy = __ldg(Q + 1), qz = __ldg(Q + 2);
    const float qn = sqnorm(qx, qy, qz);
    // the expanded-form distance differs from the true squared distance by a few ulps of (|q|^2 + |x|^2):
    // stop a side only when (x - qx)^2 exceeds tau by more than that slack
    const float slack = 4e-6f * (qn + __ldg(max_norm + b)) + 1e-30f;
    int lo = 0, hi = N;   // first sorted position with x >= qx
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&P[mid].x) < qx) lo = mid + 1; else hi = mid;
    }
    int L = lo - 1, R = lo;   // next unvisited position on each side
    bool left_on = L >= 0, right_on = R < N;
    float bd = lane < k ? INFINITY : -INFINITY;
    int bi = lane < k ? 0x7fffffff - lane : -1;
    float tau = INFINITY;
    int tau_i = 0x7fffffff;
    const bool is_left = lane < 16;
    const int sub = lane & 15;
    while (left_on || right_on) {
        const int pos = is_left ? L - sub : R + sub;
        const bool ok = is_left ? (left_on && pos >= 0) : (right_on && pos < N);
        float d = INFINITY, dx2 = 0.f;
        int id = 0x7fffffff;
        if (ok) {
            const float4 p = __ldg(P + pos);
            id = __ldg(I + pos);
            d = ref_distance(mode, qx, qy, qz, qn, p);
            const float dx = p.x - qx;
            dx2 = dx * dx;
        }
        const unsigned cand = __ballot_sync(kFull, ok && !worse(d, id, tau, tau_i));
        insert_candidates(cand, d, id, bd, bi, tau, tau_i);
        // the farthest candidate scored on each side this step (lanes 15 and 31)
        const float far_l = __shfl_sync(kFull, dx2, 15), far_r = __shfl_sync(kFull, dx2, 31);
        L -= 16; R += 16;
        if (left_on && (L < 0 || far_l > tau + slack)) left_on = false;
        if (right_on && (R >= N || far_r > tau + slack)) right_on = false;
    }
    write_result(X, qx, qy, qz, lane, k, bi, (size_t)b * S + q, out, rel);
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int64_t pvraft_knn_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0 || N > kSortMaxN) return 0;   // brute force needs no workspace
    return (int64_t)B * N * (int64_t)(sizeof(float4) + sizeof(int32_t)) + (int64_t)B * sizeof(float) + 256;
}

extern "C" int pvraft_knn_fwd(const float* xyz, const float* query, int B, int N, int S, int k, int mode, int32_t* idx,
                              float* rel, void* workspace, void* stream) {
    if (!xyz || !query || !idx) return fail(PVRAFT_ERR_BAD_ARG, "knn: null pointer");
    if (B <= 0 || N <= 0 || S <= 0) return fail(PVRAFT_ERR_BAD_ARG, "knn: bad shape");
    if (k < 1 || k > 32 || k > N) return fail(PVRAFT_ERR_UNSUPPORTED, "knn: k=%d (need 1 <= k <= min(32, N=%d))", k, N);
    if (mode != 0 && mode != 1) return fail(PVRAFT_ERR_BAD_ARG, "knn: mode=%d", mode);
    if (B > 65535) return fail(PVRAFT_ERR_UNSUPPORTED, "knn: B=%d", B);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((S + kKnnThreads / 32 - 1) / (kKnnThreads / 32), B);
    if (workspace && N <= kSortMaxN && N >= 64) {
        int NP = 1;
        while (NP < N) NP <<= 1;
        // workspace layout: float4 sorted[B*N] | int32 ids[B*N] | float max_norm[B]
        float4* sorted = reinterpret_cast<float4*>(workspace);
        int32_t* ids = reinterpret_cast<int32_t*>(sorted + (size_t)B * N);
        float* max_norm = reinterpret_cast<float*>(ids + (size_t)B * N);
        const size_t smem = (size_t)NP * sizeof(unsigned long long);
        int rc;
        if ((rc = opt_in_smem(k_sort_x, smem))) return rc;
        k_sort_x<<<B, 1024, smem, st>>>(xyz, N, NP, sorted, ids, max_norm);
        if ((rc = check_launch("knn sort"))) return rc;
        k_knn_sweep<<<grid, kKnnThreads, 0, st>>>(xyz, sorted, ids, max_norm, query, N, S, k, mode, idx, rel);
        return check_launch("knn sweep");
    }
    k_knn<<<grid, kKnnThreads, 0, st>>>(xyz, query, N, S, k, mode, idx, rel);
    return check_launch("knn");
}
