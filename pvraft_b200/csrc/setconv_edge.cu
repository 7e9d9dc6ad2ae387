// SetConv edge stage (reference model/flot/gconv.py:65-80): the gather over the 32-NN graph, fc1 on
// [x_j - x_i, rel_xyz], the statistics of gn1 and the max-pool over neighbours, without ever
// materialising the reference's [B, C+3, 32, N] edge tensor.
//
// fc1 is linear and bias-free, so  W.[x_j - x_i, e] = P_j - P_i + W_e.e  with  P = W[:, :cin].x  computed
// once per point by pvraft_linear_fwd; this kernel only gathers the 32 rows P_j (L2-resident table).
// GroupNorm + LeakyReLU is monotone per channel, so max_e lrelu(GN(y_e)) = lrelu(GN(max_e y_e)) when the
// folded GN scale is >= 0 and lrelu(GN(min_e y_e)) otherwise: the kernel emits both the per-channel max
// and min of the raw y, plus the double-precision (sum, sum^2) of all N*32*C raw values per group.
//
// One warp per point; lane l owns the adjacent channel pairs (2l, 2l+1) + 64q: a neighbour row is one 8-byte load per lane
// and pair, and the per-edge scalars (neighbour id, edge vector) are read back as ONE broadcast 16-byte shared-memory load.
#include "common.cuh"

namespace pvraft {

constexpr int kEdgeThreads = 256;

// packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): two IEEE-rounded operations per instruction, bit-identical to
// the scalar forms
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk(unsigned long long v) {
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(v));
    return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int PAIRS>
__global__ void __launch_bounds__(kEdgeThreads, PAIRS == 1 ? 6 : 4) k_setconv_edge_pairs(const float* __restrict__ fc1p, const int32_t* __restrict__ nbr,
                                                                     const float* __restrict__ edge_feats, const float* __restrict__ w_fc1,
                                                                     int cin, int B, int N, int C, float* __restrict__ ymax,
                                                                     float* __restrict__ ymin, double* __restrict__ stats,
                                                                     const int32_t* __restrict__ order) {
    __shared__ double s_part[kEdgeThreads / 32][128][2];
    __shared__ __align__(16) float4 s_edge[kEdgeThreads / 32][32];   // (neighbour id bits, ex, ey, ez) of the warp's point
    pdl_trigger();   // the next kernel may be staged while this one drains
    pdl_wait();      // (launched with PDL: nothing above touches global memory)
    const int lane = lane_id(), w = warp_id(), nwarps = kEdgeThreads / 32;
    const int ld = cin + 3;
    float2 wx[PAIRS], wy[PAIRS], wz[PAIRS];
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
        const int c = min(2 * lane + 64 * q, C - 2);   // (lanes past C replicate the last pair; their results are dropped)
        wx[q] = make_float2(__ldg(w_fc1 + (size_t)c * ld + cin + 0), __ldg(w_fc1 + (size_t)(c + 1) * ld + cin + 0));
        wy[q] = make_float2(__ldg(w_fc1 + (size_t)c * ld + cin + 1), __ldg(w_fc1 + (size_t)(c + 1) * ld + cin + 1));
        wz[q] = make_float2(__ldg(w_fc1 + (size_t)c * ld + cin + 2), __ldg(w_fc1 + (size_t)(c + 1) * ld + cin + 2));
    }
    unsigned long long wx2[PAIRS], wy2[PAIRS], wz2[PAIRS];
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) { wx2[q] = pk(wx[q].x, wx[q].y); wy2[q] = pk(wy[q].x, wy[q].y); wz2[q] = pk(wz[q].x, wz[q].y); }
    const long long total = (long long)B * N;
    long long pt_begin, pt_end;
    split_range(total, gridDim.x, blockIdx.x, pt_begin, pt_end);   // (handing SM-mates adjacent ranges was measured: no gain)
    long long seg = pt_begin;
    while (seg < pt_end) {
        const int b = (int)(seg / N);
        long long seg_end = (long long)(b + 1) * N;
        if (seg_end > pt_end) seg_end = pt_end;
        double dS[PAIRS][2], dSS[PAIRS][2];
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) { dS[q][0] = dS[q][1] = 0.0; dSS[q][0] = dSS[q][1] = 0.0; }
        bool on[PAIRS];   // C need not fill the last group of 64 channels (encoder layers: 16, 48, 96)
        int coff[PAIRS];
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) { on[q] = 2 * lane + 64 * q < C; coff[q] = on[q] ? 2 * lane + 64 * q : 0; }
        const float* P = fc1p + (size_t)b * N * C;
        for (long long r = seg + w; r < seg_end; r += nwarps) {
            // processing order: with `order` (a space-filling-curve rank -> point table) the 8 warps of a CTA work on spatial
            // neighbours at the same time, whose 32-neighbourhoods overlap: their gathers hit the same rows in L1
            const int i = order ? __ldg(order + r) : (int)(r - (long long)b * N);
            const long long pt = (long long)b * N + i;
            // lane e parks neighbour e: id and edge feature x_j - x_i (graph.edge_feats, gconv.py:66)
            const float* ef = edge_feats + ((size_t)pt * 32 + lane) * 3;
            __syncwarp();
            s_edge[w][lane] = make_float4(__int_as_float(__ldg(nbr + pt * 32 + lane) * C), __ldg(ef), __ldg(ef + 1), __ldg(ef + 2));
            __syncwarp();
            unsigned long long pi[PAIRS], s1[PAIRS], s2[PAIRS];
            float2 mx[PAIRS], mn[PAIRS];
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(P + (size_t)i * C + coff[q]));
                pi[q] = pk(t.x, t.y);
                mx[q] = make_float2(-INFINITY, -INFINITY); mn[q] = make_float2(INFINITY, INFINITY);
                s1[q] = pk(0.f, 0.f); s2[q] = pk(0.f, 0.f);
            }
#pragma unroll 8
            for (int e = 0; e < 32; ++e) {
                const float4 ed = s_edge[w][e];
                const float* row = P + __float_as_int(ed.x);
                const unsigned long long ex = pk(ed.y, ed.y), ey = pk(ed.z, ed.z), ez = pk(ed.w, ed.w);
#pragma unroll
                for (int q = 0; q < PAIRS; ++q) {
                    const float2 pj = __ldg(reinterpret_cast<const float2*>(row + coff[q]));
                    // y = (P_j - P_i) + fma(w_z, e_z, fma(w_y, e_y, w_x * e_x)), both channels of the pair at once
                    const unsigned long long t = fma2(wz2[q], ez, fma2(wy2[q], ey, mul2(wx2[q], ex)));
                    const unsigned long long y2 = add2(sub2(pk(pj.x, pj.y), pi[q]), t);
                    const float2 y = upk(y2);
                    mx[q].x = fmaxf(mx[q].x, y.x); mx[q].y = fmaxf(mx[q].y, y.y);
                    mn[q].x = fminf(mn[q].x, y.x); mn[q].y = fminf(mn[q].y, y.y);
                    s1[q] = add2(s1[q], y2);
                    s2[q] = fma2(y2, y2, s2[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
                const size_t o = (size_t)pt * C + coff[q];
                if (on[q]) {
                    *reinterpret_cast<float2*>(ymax + o) = mx[q];
                    *reinterpret_cast<float2*>(ymin + o) = mn[q];
                }
                const float2 a1 = upk(s1[q]), a2 = upk(s2[q]);
                dS[q][0] += (double)a1.x; dS[q][1] += (double)a1.y;
                dSS[q][0] += (double)a2.x; dSS[q][1] += (double)a2.y;
            }
        }
        // block reduction of the per-channel partials -> per-group sums -> one atomic per (group, moment)
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (on[q]) {
                    s_part[w][2 * lane + 64 * q + h][0] = dS[q][h];
                    s_part[w][2 * lane + 64 * q + h][1] = dSS[q][h];
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            const int g = threadIdx.x >> 1, m = threadIdx.x & 1, gsz = C / PVRAFT_GN_GROUPS;
            double acc = 0.0;
            for (int c = g * gsz; c < (g + 1) * gsz; ++c)
                for (int ww = 0; ww < nwarps; ++ww) acc += s_part[ww][c][m];
            if (acc != 0.0) atomicAdd(stats + (size_t)b * 16 + threadIdx.x, acc);
        }
        seg = seg_end;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tiled variant for a spatially coherent processing order (`order` = Morton rank table, N % 32 == 0, N <= 16384).
// A CTA works on tiles of 32 consecutive points of the order.  Their 1024 neighbour references hit only a few hundred
// distinct rows (neighbourhoods of neighbours overlap), so the tile first builds the SET of referenced rows -- a bitmap over the
// sample's points, a block scan of its popcounts gives every row a slot -- copies those rows from L2 into shared memory once
// (coalesced 256-byte rows) and then gathers from shared memory: the edge loop's one global 8-byte gather per neighbour and
// channel pair becomes a conflict-free LDS.64, and the kernel leaves the L2 -> SM gather bandwidth it was bound by.
// A tile whose set does not fit (the curve jumped: two distant blobs) gathers from global memory like the kernel above.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTileP = 32;                 // points per tile
constexpr int kTabBytes = 80 * 1024;       // shared-memory row table

template <int PAIRS>
__global__ void __launch_bounds__(kEdgeThreads, 2) k_setconv_edge_tiled(const float* __restrict__ fc1p, const int32_t* __restrict__ nbr,
                                                                        const float* __restrict__ edge_feats, const float* __restrict__ w_fc1,
                                                                        int cin, int B, int N, int C, float* __restrict__ ymax,
                                                                        float* __restrict__ ymin, double* __restrict__ stats,
                                                                        const int32_t* __restrict__ order) {
    extern __shared__ __align__(16) unsigned char smem_e[];
    const int words = N >> 5;                                                     // bitmap words (<= 512)
    float* s_tab = reinterpret_cast<float*>(smem_e);                              // [cap][C]
    unsigned* s_bits = reinterpret_cast<unsigned*>(smem_e + kTabBytes);           // [512]
    int* s_pref = reinterpret_cast<int*>(s_bits + 512);                           // [512] rows before word i
    int* s_uid = s_pref + 512;                                                    // [cap <= 1280]
    float4* s_edge = reinterpret_cast<float4*>(s_uid + 1280);                     // [8 warps][32]
    double* s_part = reinterpret_cast<double*>(s_edge + 8 * 32);                  // [8 warps][128][2]
    __shared__ int s_wsum[kEdgeThreads / 32];
    __shared__ int s_total;
    pdl_trigger();
    pdl_wait();
    const int lane = lane_id(), w = warp_id(), nwarps = kEdgeThreads / 32, tid = threadIdx.x;
    const int ld = cin + 3;
    const int cap = min(1280, kTabBytes / (C * 4));
    unsigned long long wx2[PAIRS], wy2[PAIRS], wz2[PAIRS];
    bool on[PAIRS];
    int coff[PAIRS];
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
        const int c = min(2 * lane + 64 * q, C - 2);
        wx2[q] = pk(__ldg(w_fc1 + (size_t)c * ld + cin + 0), __ldg(w_fc1 + (size_t)(c + 1) * ld + cin + 0));
        wy2[q] = pk(__ldg(w_fc1 + (size_t)c * ld + cin + 1), __ldg(w_fc1 + (size_t)(c + 1) * ld + cin + 1));
        wz2[q] = pk(__ldg(w_fc1 + (size_t)c * ld + cin + 2), __ldg(w_fc1 + (size_t)(c + 1) * ld + cin + 2));
        on[q] = 2 * lane + 64 * q < C;
        coff[q] = on[q] ? 2 * lane + 64 * q : 0;
    }
    const int tiles_per_sample = N / kTileP;
    const long long n_tiles = (long long)B * tiles_per_sample;
    long long t_begin, t_end;
    split_range(n_tiles, gridDim.x, blockIdx.x, t_begin, t_end);
    double dS[PAIRS][2], dSS[PAIRS][2];
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) { dS[q][0] = dS[q][1] = 0.0; dSS[q][0] = dSS[q][1] = 0.0; }
    int cur_b = -1;
    auto flush = [&](int b) {   // per-channel partials of sample b -> per-group sums -> one atomic per (group, moment)
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (on[q]) {
                    s_part[((size_t)w * 128 + 2 * lane + 64 * q + h) * 2 + 0] = dS[q][h];
                    s_part[((size_t)w * 128 + 2 * lane + 64 * q + h) * 2 + 1] = dSS[q][h];
                }
                dS[q][h] = 0.0; dSS[q][h] = 0.0;
            }
        }
        __syncthreads();
        if (tid < 16) {
            const int g = tid >> 1, m = tid & 1, gsz = C / PVRAFT_GN_GROUPS;
            double acc = 0.0;
            for (int c = g * gsz; c < (g + 1) * gsz; ++c)
                for (int ww = 0; ww < nwarps; ++ww) acc += s_part[((size_t)ww * 128 + c) * 2 + m];
            if (acc != 0.0) atomicAdd(stats + (size_t)b * 16 + tid, acc);
        }
    };
    for (long long t = t_begin; t < t_end; ++t) {
        const int b = (int)(t / tiles_per_sample);
        if (b != cur_b) {
            if (cur_b >= 0) flush(cur_b);
            cur_b = b;
        }
        const float* P = fc1p + (size_t)b * N * C;
        const long long r0 = t * kTileP;                       // first rank of the tile (global: b*N + rank)
        // ---- the set of rows the tile references ----
        __syncthreads();                                       // previous tile's readers of the table / bitmap are done
        for (int i = tid; i < words; i += kEdgeThreads) s_bits[i] = 0u;
        __syncthreads();
        {
            const int pi = tid >> 3, e4 = (tid & 7) * 4;       // point of the tile, first of this thread's 4 edges
            const long long pt = (long long)b * N + __ldg(order + r0 + pi);
            const int4 id = __ldg(reinterpret_cast<const int4*>(nbr + pt * 32 + e4));
            atomicOr(&s_bits[id.x >> 5], 1u << (id.x & 31));
            atomicOr(&s_bits[id.y >> 5], 1u << (id.y & 31));
            atomicOr(&s_bits[id.z >> 5], 1u << (id.z & 31));
            atomicOr(&s_bits[id.w >> 5], 1u << (id.w & 31));
        }
        __syncthreads();
        {   // exclusive scan of the per-word popcounts: thread i owns words 2i, 2i+1
            const unsigned b0 = 2 * tid < words ? s_bits[2 * tid] : 0u, b1 = 2 * tid + 1 < words ? s_bits[2 * tid + 1] : 0u;
            const int c0 = __popc(b0), c1 = __popc(b1);
            int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += a;
            }
            if (lane == 31) s_wsum[w] = incl;
            __syncthreads();
            int base = incl - (c0 + c1);
            for (int ww = 0; ww < w; ++ww) base += s_wsum[ww];
            if (2 * tid < words) s_pref[2 * tid] = base;
            if (2 * tid + 1 < words) s_pref[2 * tid + 1] = base + c0;
            if (tid == kEdgeThreads - 1) s_total = base + c0 + c1;
            // the rows themselves, in slot order (only meaningful when they fit; harmless otherwise: bounded by cap)
            int at = base;
            unsigned m = b0;
            while (m) { const int bit = __ffs(m) - 1; m &= m - 1; if (at < cap) s_uid[at] = (2 * tid) * 32 + bit; ++at; }
            m = b1;
            while (m) { const int bit = __ffs(m) - 1; m &= m - 1; if (at < cap) s_uid[at] = (2 * tid + 1) * 32 + bit; ++at; }
        }
        __syncthreads();
        const int total = s_total;
        const bool in_smem = total <= cap;
        if (in_smem) {   // copy the referenced rows: a warp per row, 8 bytes per lane and channel pair
            for (int u = w; u < total; u += nwarps) {
                const float* row = P + (size_t)s_uid[u] * C;
#pragma unroll
                for (int q = 0; q < PAIRS; ++q)
                    if (on[q]) *reinterpret_cast<float2*>(s_tab + (size_t)u * C + coff[q]) = __ldg(reinterpret_cast<const float2*>(row + coff[q]));
            }
        }
        __syncthreads();
        // ---- the edge stage of the tile's points: warp w takes points w, w + 8, ... ----
        for (int pi = w; pi < kTileP; pi += nwarps) {
            const int i = __ldg(order + r0 + pi);
            const long long pt = (long long)b * N + i;
            const float* ef = edge_feats + ((size_t)pt * 32 + lane) * 3;
            const int id = __ldg(nbr + pt * 32 + lane);
            // where neighbour `lane` lives: its slot in the shared-memory table, or its row in global memory
            const int slot = s_pref[id >> 5] + __popc(s_bits[id >> 5] & ((1u << (id & 31)) - 1u));
            __syncwarp();
            s_edge[w * 32 + lane] = make_float4(__int_as_float((in_smem ? slot : id) * C), __ldg(ef), __ldg(ef + 1), __ldg(ef + 2));
            __syncwarp();
            unsigned long long pc[PAIRS], s1[PAIRS], s2[PAIRS];
            float2 mx[PAIRS], mn[PAIRS];
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
                const float2 tt = __ldg(reinterpret_cast<const float2*>(P + (size_t)i * C + coff[q]));
                pc[q] = pk(tt.x, tt.y);
                mx[q] = make_float2(-INFINITY, -INFINITY); mn[q] = make_float2(INFINITY, INFINITY);
                s1[q] = pk(0.f, 0.f); s2[q] = pk(0.f, 0.f);
            }
            // (two copies of the loop so that the common one reads the table with shared-memory loads, not generic ones)
#define PVRAFT_EDGE_LOOP(ROWPTR, LOAD)                                                                              \
            _Pragma("unroll 8") for (int e = 0; e < 32; ++e) {                                                      \
                const float4 ed = s_edge[w * 32 + e];                                                               \
                const float* row = (ROWPTR) + __float_as_int(ed.x);                                                 \
                const unsigned long long ex = pk(ed.y, ed.y), ey = pk(ed.z, ed.z), ez = pk(ed.w, ed.w);             \
                _Pragma("unroll") for (int q = 0; q < PAIRS; ++q) {                                                 \
                    const float2 pj = LOAD(reinterpret_cast<const float2*>(row + coff[q]));                         \
                    const unsigned long long tq = fma2(wz2[q], ez, fma2(wy2[q], ey, mul2(wx2[q], ex)));             \
                    const unsigned long long y2 = add2(sub2(pk(pj.x, pj.y), pc[q]), tq);                            \
                    const float2 y = upk(y2);                                                                       \
                    mx[q].x = fmaxf(mx[q].x, y.x); mx[q].y = fmaxf(mx[q].y, y.y);                                   \
                    mn[q].x = fminf(mn[q].x, y.x); mn[q].y = fminf(mn[q].y, y.y);                                   \
                    s1[q] = add2(s1[q], y2);                                                                        \
                    s2[q] = fma2(y2, y2, s2[q]);                                                                    \
                }                                                                                                   \
            }
#define PVRAFT_LD_SHARED(p) (*(p))
            if (in_smem) { PVRAFT_EDGE_LOOP(s_tab, PVRAFT_LD_SHARED) } else { PVRAFT_EDGE_LOOP(P, __ldg) }
#undef PVRAFT_EDGE_LOOP
#undef PVRAFT_LD_SHARED
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
                const size_t o = (size_t)pt * C + coff[q];
                if (on[q]) {
                    *reinterpret_cast<float2*>(ymax + o) = mx[q];
                    *reinterpret_cast<float2*>(ymin + o) = mn[q];
                }
                const float2 a1 = upk(s1[q]), a2 = upk(s2[q]);
                dS[q][0] += (double)a1.x; dS[q][1] += (double)a1.y;
                dSS[q][0] += (double)a2.x; dSS[q][1] += (double)a2.y;
            }
        }
    }
    if (cur_b >= 0) flush(cur_b);
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int pvraft_setconv_edge_fwd(const float* fc1p, const int32_t* nbr, const float* edge_feats, const float* w_fc1, int cin,
                                       int B, int N, int C, float* ymax, float* ymin, double* stats, const int32_t* order, void* stream) {
    if (!fc1p || !nbr || !edge_feats || !w_fc1 || !ymax || !ymin || !stats) return fail(PVRAFT_ERR_BAD_ARG, "setconv_edge: null pointer");
    if (B <= 0 || N <= 0 || cin <= 0) return fail(PVRAFT_ERR_BAD_ARG, "setconv_edge: bad shape");
    if (C <= 0 || C > 128 || C % PVRAFT_GN_GROUPS) return fail(PVRAFT_ERR_UNSUPPORTED, "setconv_edge: C=%d (multiple of 8, <= 128)", C);
    const long long total = (long long)B * N;
    long long g = (long long)sm_count() * 8;
    const long long need = (total + (kEdgeThreads / 32) - 1) / (kEdgeThreads / 32);
    if (g > need) g = need;
    const int grid = (int)(g < 1 ? 1 : g);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool tiled_on = []() { const char* e = getenv("PVRAFT_EDGE_TILED"); return !(e && atoi(e) == 0); }();
    if (tiled_on && order && N % kTileP == 0 && N <= 16384) {
        // shared-memory row table: spatially coherent tiles of 32 points gather their (few hundred) distinct neighbour rows once
        const size_t smem = (size_t)kTabBytes + 512 * 4 + 512 * 4 + 1280 * 4 + 8 * 32 * 16 + 8 * 128 * 2 * 8;
        const long long n_tiles = total / kTileP;
        long long gt = (long long)sm_count() * 2;
        if (gt > n_tiles) gt = n_tiles;
        int rc;
        if (C <= 64) {
            if ((rc = opt_in_smem(k_setconv_edge_tiled<1>, smem))) return rc;
            launch_pdl(k_setconv_edge_tiled<1>, (int)gt, kEdgeThreads, smem, st, fc1p, nbr, edge_feats, w_fc1, cin, B, N, C, ymax, ymin, stats, order);
        } else {
            if ((rc = opt_in_smem(k_setconv_edge_tiled<2>, smem))) return rc;
            launch_pdl(k_setconv_edge_tiled<2>, (int)gt, kEdgeThreads, smem, st, fc1p, nbr, edge_feats, w_fc1, cin, B, N, C, ymax, ymin, stats, order);
        }
        return check_launch("setconv_edge");
    }
    if (C <= 64) {
        launch_pdl(k_setconv_edge_pairs<1>, grid, kEdgeThreads, 0, st, fc1p, nbr, edge_feats, w_fc1, cin, B, N, C, ymax, ymin, stats, order);
        return check_launch("setconv_edge");
    }
    if (C <= 128) {
        launch_pdl(k_setconv_edge_pairs<2>, grid, kEdgeThreads, 0, st, fc1p, nbr, edge_feats, w_fc1, cin, B, N, C, ymax, ymin, stats, order);
        return check_launch("setconv_edge");
    }
    return fail(PVRAFT_ERR_UNSUPPORTED, "setconv_edge: C=%d", C);
}
