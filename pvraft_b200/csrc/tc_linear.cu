// Per-point linear layers on the tcgen05 tensor cores, fp32-accurate (3xTF32), with the GroupNorm / activation
// prologue and the bias / activation / residual / GroupNorm-statistics / GRU-gate epilogues fused around the MMA.
//
//   out[M x N] = epilogue( prologue(A)[M x K] . W[N x K]^T )          M = B*Npts points, N = cout, K = cin
//
// A CTA owns 128 consecutive points (= 128 TMEM lanes).  Per 32-channel k-block:
//   warp 0      TMA: the raw fp32 activation box [128 x 32] (SWIZZLE_128B; up to three source tensors are
//               concatenated along K, e.g. [h | inp | motion] for the GRU) and the pre-split weight boxes W_hi, W_lo
//   warps 2-5   transform: every 16-byte chunk of the raw box gets the folded GroupNorm affine + activation of its
//               channels (optionally choosing the max or the min input by the sign of the scale), is split into
//               hi = tf32(x), lo = tf32(x - hi) and written to the hi / lo boxes AT THE SAME swizzled offset (the split
//               is elementwise, so no swizzle arithmetic is needed); fence.proxy.async; arrive on the stage barrier
//   warp 1      one lane issues A_hi.W_hi + A_lo.W_hi + A_hi.W_lo (3 x 4 tcgen05.mma.kind::tf32, K = 8 each) into TMEM
//   warps 2-5   epilogue: tcgen05.ld (thread = point) -> bias / ReLU / residual / GRU gates -> global, GroupNorm
//               (sum, sum^2) per group reduced in the warp and accumulated with double atomics
// Replaces the k_linear / k_gru CUDA-core kernels whenever Npts % 128 == 0 and cin % 32 == 0.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace pvraft {

constexpr int kTcThreads = 192;
constexpr int kTcM = 128, kTcKB = 32;
constexpr int kTcABytes = kTcM * kTcKB * 4;   // 16 KB: one activation box
constexpr int kTcMaxStages = 4;

enum TcEpilogue { TC_EPI_PLAIN = 0, TC_EPI_GRU_ZR = 1, TC_EPI_GRU_Q = 2 };

struct TcParams {
    // prologue (per input channel, per sample): x = act(raw * scale + shift); raw = max or min input by sign(scale)
    const double* in_stats;   // [B,8,2] or null (plain)
    const float* in_gamma;
    const float* in_beta;
    double in_count;
    int in_act;
    float in_slope;
    int minmax;               // 1: second raw source holds the per-channel minima
    // epilogue
    int epi;
    const float* bias;        // [N] or null
    const float* bias2;       // GRU_ZR: bias of r
    int out_act;
    const float* residual;    // [M,N] or null
    float* out;               // [M,N]   (GRU_ZR: z [M,64]; GRU_Q: new hidden state [M,64])
    float* out2;              // GRU_ZR: r*h [M,64]
    const float* h;           // GRU: previous hidden state [M,64]
    const float* z;           // GRU_Q: update gate [M,64]
    double* out_stats;        // [B,8,2] or null
    int M, N, K, cout, pts_per_sample;
    int seg_kb[3];            // k-blocks contributed by each activation source
    int stages;               // depth of the shared-memory ring (1..4), chosen by the host for occupancy
};

__device__ __forceinline__ unsigned tsu32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tsu32(bar)), "r"(count));
}
__device__ __forceinline__ void tmbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tsu32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tsu32(bar)) : "memory");
}
__device__ __forceinline__ void tmbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(tsu32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void ttma_load_2d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(tsu32(dst)),
                 "l"(map), "r"(tsu32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ unsigned long long tumma_desc(const void* smem_tile) {   // K-major, SWIZZLE_128B (see corr_gemm.cu)
    unsigned long long d = 0;
    d |= (unsigned long long)((tsu32(smem_tile) >> 4) & 0x3FFFu);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)64 << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)2 << 61;
    return d;
}
__device__ __forceinline__ void tumma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tumma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tsu32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tsigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kTcThreads, 3)
k_tc_linear(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_amin,
            const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // stage layout: [A hi (raw in, hi out) 16K][A lo 16K][W hi N*128][W lo N*128][A min raw 16K, only with minmax]
    const int w_bytes = p.N * kTcKB * 4;
    const int w_off = 2 * kTcABytes, min_off = 2 * kTcABytes + 2 * w_bytes;
    const int stage_bytes = (p.minmax ? 3 : 2) * kTcABytes + 2 * w_bytes;
    const int kTcStages = p.stages;
    float* s_scale = reinterpret_cast<float*>(tiles + (size_t)kTcStages * stage_bytes);   // [K]
    float* s_shift = s_scale + p.K;
    float* s_bias = s_shift + p.K;                                                        // [2 * N]
    __shared__ __align__(8) unsigned long long s_full[kTcMaxStages], s_ready[kTcMaxStages], s_empty[kTcMaxStages], s_tmem_full;
    __shared__ unsigned s_tmem_base;
    const int warp = warp_id(), lane = lane_id();
    const int tile = blockIdx.x;
    const int row0 = tile * kTcM;
    const int sample = row0 / p.pts_per_sample;
    const int num_kb = p.K / kTcKB;
    const unsigned tmem_cols = p.N <= 32 ? 32u : p.N <= 64 ? 64u : p.N <= 128 ? 128u : 256u;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTcStages; ++s) { tmbar_init(&s_full[s], 1); tmbar_init(&s_ready[s], 128); tmbar_init(&s_empty[s], 1); }
        tmbar_init(&s_tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tsu32(&s_tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && p.in_stats != nullptr) {   // folded GroupNorm affine of every input channel of this sample
        const int gsz = p.K / PVRAFT_GN_GROUPS;
        for (int k = threadIdx.x - 64; k < p.K; k += 128) {
            const GnAffine af = gn_affine(p.in_stats + (size_t)sample * 16 + (k / gsz) * 2, p.in_count, __ldg(p.in_gamma + k), __ldg(p.in_beta + k));
            s_scale[k] = af.scale;
            s_shift[k] = af.shift;
        }
    }
    if (warp >= 2) {
        for (int c = threadIdx.x - 64; c < p.N; c += 128) {
            s_bias[c] = (p.bias != nullptr && c < p.cout) ? __ldg(p.bias + c) : 0.f;
            s_bias[p.N + c] = (p.bias2 != nullptr && c < p.cout) ? __ldg(p.bias2 + c) : 0.f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                const unsigned phase = (unsigned)(kb / kTcStages) & 1u;
                tmbar_wait(&s_empty[s], phase ^ 1u);
                unsigned char* st = tiles + (size_t)s * stage_bytes;
                tmbar_expect_tx(&s_full[s], (unsigned)(kTcABytes * (p.minmax ? 2 : 1) + 2 * w_bytes));
                // which source tensor does this k-block come from?
                int seg = 0, kk = kb;
                if (kk >= p.seg_kb[0]) { kk -= p.seg_kb[0]; seg = 1; if (kk >= p.seg_kb[1]) { kk -= p.seg_kb[1]; seg = 2; } }
                const CUtensorMap* ma = seg == 0 ? &map_a0 : seg == 1 ? &map_a1 : &map_a2;
                ttma_load_2d(st, ma, &s_full[s], kk * kTcKB, row0);
                if (p.minmax) ttma_load_2d(st + min_off, &map_amin, &s_full[s], kk * kTcKB, row0);
                ttma_load_2d(st + w_off, &map_w_hi, &s_full[s], kb * kTcKB, 0);
                ttma_load_2d(st + w_off + w_bytes, &map_w_lo, &s_full[s], kb * kTcKB, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(p.N >> 3) << 17) | ((unsigned)(kTcM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                const unsigned phase = (unsigned)(kb / kTcStages) & 1u;
                tmbar_wait(&s_ready[s], phase);   // transformed activations (and, transitively, the TMA data) are in place
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char* st = tiles + (size_t)s * stage_bytes;
                const unsigned long long a_hi = tumma_desc(st), a_lo = tumma_desc(st + kTcABytes);
                const unsigned long long b_hi = tumma_desc(st + w_off), b_lo = tumma_desc(st + w_off + w_bytes);
#pragma unroll
                for (int k = 0; k < kTcKB / 8; ++k) {
                    const unsigned long long off = (unsigned long long)(k * 2);
                    tumma_tf32(tmem, a_hi + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
                    tumma_tf32(tmem, a_lo + off, b_hi + off, idesc, 1u);
                    tumma_tf32(tmem, a_hi + off, b_lo + off, idesc, 1u);
                }
                tumma_commit(&s_empty[s]);
                if (kb == num_kb - 1) tumma_commit(&s_tmem_full);
            }
        }
    } else {
        // ===== transform (prologue + hi/lo split), then epilogue =====
        const int t = threadIdx.x - 64;   // 0..127
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kTcStages;
            const unsigned phase = (unsigned)(kb / kTcStages) & 1u;
            tmbar_wait(&s_full[s], phase);
            unsigned char* st = tiles + (size_t)s * stage_bytes;
#pragma unroll 2
            for (int c = t; c < kTcABytes / 16; c += 128) {
                float4 x = *reinterpret_cast<const float4*>(st + (size_t)c * 16);
                if (p.in_stats != nullptr) {
                    const int r = c >> 3, lc = (c & 7) ^ (r & 7);      // logical 16-byte chunk of row r under SWIZZLE_128B
                    const int k = kb * kTcKB + lc * 4;
                    const float4 sc = *reinterpret_cast<const float4*>(s_scale + k);
                    const float4 sh = *reinterpret_cast<const float4*>(s_shift + k);
                    if (p.minmax) {
                        const float4 mn = *reinterpret_cast<const float4*>(st + min_off + (size_t)c * 16);
                        x.x = sc.x < 0.f ? mn.x : x.x; x.y = sc.y < 0.f ? mn.y : x.y;
                        x.z = sc.z < 0.f ? mn.z : x.z; x.w = sc.w < 0.f ? mn.w : x.w;
                    }
                    x.x = apply_act(fmaf(x.x, sc.x, sh.x), p.in_act, p.in_slope);
                    x.y = apply_act(fmaf(x.y, sc.y, sh.y), p.in_act, p.in_slope);
                    x.z = apply_act(fmaf(x.z, sc.z, sh.z), p.in_act, p.in_slope);
                    x.w = apply_act(fmaf(x.w, sc.w, sh.w), p.in_act, p.in_slope);
                }
                float4 hi, lo;
                hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
                lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
                *reinterpret_cast<float4*>(st + (size_t)c * 16) = hi;
                *reinterpret_cast<float4*>(st + kTcABytes + (size_t)c * 16) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
            tmbar_arrive(&s_ready[s]);
        }
        // ---- epilogue: thread = point (TMEM lane) ----
        tmbar_wait(&s_tmem_full, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quad = warp & 3;
        const int row = row0 + quad * 32 + lane;
        const bool live = row < p.M;
        const int gsz = p.cout / PVRAFT_GN_GROUPS;
        if (p.epi == TC_EPI_PLAIN) {
            const bool vec = (p.cout & 3) == 0;
            for (int c0 = 0; c0 < p.N; c0 += 32) {
                unsigned v[32];
                tmem_ld32(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)c0, v);
                float y[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + q * 4);
                    y[q * 4 + 0] = apply_act(__uint_as_float(v[q * 4 + 0]) + bv.x, p.out_act, 0.f);
                    y[q * 4 + 1] = apply_act(__uint_as_float(v[q * 4 + 1]) + bv.y, p.out_act, 0.f);
                    y[q * 4 + 2] = apply_act(__uint_as_float(v[q * 4 + 2]) + bv.z, p.out_act, 0.f);
                    y[q * 4 + 3] = apply_act(__uint_as_float(v[q * 4 + 3]) + bv.w, p.out_act, 0.f);
                }
                if (live) {
                    float* o = p.out + (size_t)row * p.cout + c0;
                    if (vec) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (c0 + q * 4 < p.cout) {
                                if (p.residual != nullptr) {
                                    const float4 rv = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)row * p.cout + c0 + q * 4));
                                    y[q * 4] += rv.x; y[q * 4 + 1] += rv.y; y[q * 4 + 2] += rv.z; y[q * 4 + 3] += rv.w;
                                }
                                *reinterpret_cast<float4*>(o + q * 4) = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (c0 + i < p.cout) {
                                if (p.residual != nullptr) y[i] += __ldg(p.residual + (size_t)row * p.cout + c0 + i);
                                o[i] = y[i];
                            }
                        }
                    }
                }
                if (p.out_stats != nullptr) {
                    // 8 partial (sum, sum^2) pairs per thread: sub-group j covers columns [4j, 4j+4) of this step; every
                    // GroupNorm group is a union of whole sub-groups (group size is a multiple of 4)
                    float s1[8], s2[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        s1[j] = 0.f; s2[j] = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float a = live ? y[j * 4 + i] : 0.f;
                            s1[j] += a;
                            s2[j] = fmaf(a, a, s2[j]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            s1[j] += __shfl_xor_sync(kFull, s1[j], o);
                            s2[j] += __shfl_xor_sync(kFull, s2[j], o);
                        }
                    }
                    if (lane < 8) {   // lane j publishes sub-group j
                        float a1 = 0.f, a2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { a1 = lane == j ? s1[j] : a1; a2 = lane == j ? s2[j] : a2; }
                        const int c = c0 + lane * 4;
                        if (c < p.cout) {
                            const int grp = c / gsz;
                            atomicAdd(p.out_stats + (size_t)sample * 16 + grp * 2 + 0, (double)a1);
                            atomicAdd(p.out_stats + (size_t)sample * 16 + grp * 2 + 1, (double)a2);
                        }
                    }
                }
            }
        } else if (p.epi == TC_EPI_GRU_ZR) {
            // accumulator columns 0..63 = z pre-activation, 64..127 = r pre-activation (model/update.py:34-35)
            for (int c0 = 0; c0 < 64; c0 += 32) {
                unsigned vz[32], vr[32];
                tmem_ld32(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)c0, vz);
                tmem_ld32(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)(64 + c0), vr);
                if (live) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int c = c0 + q * 4;
                        const float4 hv = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)row * 64 + c));
                        const float4 bz = *reinterpret_cast<const float4*>(s_bias + c);
                        const float4 br = *reinterpret_cast<const float4*>(s_bias + p.N + c);
                        float4 z, rh;
                        z.x = tsigmoid(__uint_as_float(vz[q * 4 + 0]) + bz.x); z.y = tsigmoid(__uint_as_float(vz[q * 4 + 1]) + bz.y);
                        z.z = tsigmoid(__uint_as_float(vz[q * 4 + 2]) + bz.z); z.w = tsigmoid(__uint_as_float(vz[q * 4 + 3]) + bz.w);
                        rh.x = tsigmoid(__uint_as_float(vr[q * 4 + 0]) + br.x) * hv.x; rh.y = tsigmoid(__uint_as_float(vr[q * 4 + 1]) + br.y) * hv.y;
                        rh.z = tsigmoid(__uint_as_float(vr[q * 4 + 2]) + br.z) * hv.z; rh.w = tsigmoid(__uint_as_float(vr[q * 4 + 3]) + br.w) * hv.w;
                        *reinterpret_cast<float4*>(p.out + (size_t)row * 64 + c) = z;
                        *reinterpret_cast<float4*>(p.out2 + (size_t)row * 64 + c) = rh;
                    }
                }
            }
        } else {
            // q = tanh(acc + b); h' = (1 - z) h + z q   (model/update.py:37-39)
            for (int c0 = 0; c0 < 64; c0 += 32) {
                unsigned vq[32];
                tmem_ld32(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)c0, vq);
                if (live) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int c = c0 + q * 4;
                        const float4 hv = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)row * 64 + c));
                        const float4 zv = __ldg(reinterpret_cast<const float4*>(p.z + (size_t)row * 64 + c));
                        const float4 bq = *reinterpret_cast<const float4*>(s_bias + c);
                        float4 o;
                        o.x = (1.f - zv.x) * hv.x + zv.x * tanhf(__uint_as_float(vq[q * 4 + 0]) + bq.x);
                        o.y = (1.f - zv.y) * hv.y + zv.y * tanhf(__uint_as_float(vq[q * 4 + 1]) + bq.y);
                        o.z = (1.f - zv.z) * hv.z + zv.z * tanhf(__uint_as_float(vq[q * 4 + 2]) + bq.z);
                        o.w = (1.f - zv.w) * hv.w + zv.w * tanhf(__uint_as_float(vq[q * 4 + 3]) + bq.w);
                        *reinterpret_cast<float4*>(p.out + (size_t)row * 64 + c) = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

// hi = tf32(w), lo = tf32(w - hi) of a [rows, ld] weight window [rows, cols] written as [rows_pad, cols_pad] (zero padded)
__global__ void k_weight_split(const float* __restrict__ w, int rows, int cols, int ld, int col0, int rows_pad, int cols_pad,
                               float* __restrict__ hi, float* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_pad * cols_pad) return;
    const int r = i / cols_pad, c = i - r * cols_pad;
    float x = 0.f;
    if (r < rows && c < cols) x = __ldg(w + (size_t)r * ld + col0 + c);
    const float h = tf32_rna(x);
    hi[i] = h;
    lo[i] = tf32_rna(x - h);
}

typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeFn tc_encode_fn() {
    static TcEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TcEncodeFn>(p);
    }
    return fn;
}
// [rows, cols] fp32 row-major (row stride ld floats), box = box_rows x 32 columns, 128-byte swizzle
static int tc_make_map(CUtensorMap* m, const float* base, long long rows, int cols, long long ld, int box_rows) {
    TcEncodeFn fn = tc_encode_fn();
    if (!fn) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int pvraft_tc_weight_split(const float* w, int rows, int cols, int ld, int col0, int rows_pad, int cols_pad, float* hi,
                                      float* lo, void* stream) {
    if (!w || !hi || !lo || rows <= 0 || cols <= 0 || rows_pad < rows || cols_pad < cols) return fail(PVRAFT_ERR_BAD_ARG, "tc_weight_split: bad argument");
    const int n = rows_pad * cols_pad;
    k_weight_split<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, ld > 0 ? ld : cols, col0, rows_pad, cols_pad, hi, lo);
    return check_launch("tc_weight_split");
}

extern "C" int pvraft_tc_linear_fwd(const pvraft_tc_linear_args* a, void* stream) {
    if (!a || !a->in[0] || !a->w_hi || !a->w_lo || !a->out) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: null pointer");
    if (a->B <= 0 || a->N <= 0 || a->n_pad < 16 || a->n_pad > 256 || a->n_pad % 16) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: bad shape (n_pad=%d)", a->n_pad);
    if (a->N % kTcM) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: points per sample (%d) must be a multiple of 128", a->N);
    int K = 0;
    for (int s = 0; s < 3; ++s) {
        if (a->in[s] && (a->in_channels[s] <= 0 || a->in_channels[s] % kTcKB)) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: source %d has %d channels (multiple of 32 needed)", s, a->in_channels[s]);
        if (a->in[s]) K += a->in_channels[s];
    }
    if (K <= 0 || K > 512) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: K=%d", K);
    if (a->in_stats && (K % PVRAFT_GN_GROUPS || !a->in_gamma || !a->in_beta || a->in[1])) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GroupNorm prologue needs a single source with K %% 8 == 0");
    if (a->in_min && !a->in_stats) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: in_min needs the GroupNorm prologue");
    if (a->out_stats && (a->epilogue != TC_EPI_PLAIN || a->cout % PVRAFT_GN_GROUPS || (a->cout / PVRAFT_GN_GROUPS) % 4)) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: out_stats needs a GroupNorm group size that is a multiple of 4 (cout=%d)", a->cout);
    if (a->epilogue != TC_EPI_PLAIN && (a->cout != 64 || !a->h || !a->bias)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GRU epilogues need cout=64, h and bias");
    if (a->epilogue == TC_EPI_GRU_ZR && (a->n_pad != 128 || !a->bias2 || !a->out2)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GRU zr epilogue needs n_pad=128, bias2, out2");
    if (a->epilogue == TC_EPI_GRU_Q && (a->n_pad != 64 || !a->z)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GRU q epilogue needs n_pad=64 and z");
    const long long M = (long long)a->B * a->N;
    TcParams p{};
    p.in_stats = a->in_stats; p.in_gamma = a->in_gamma; p.in_beta = a->in_beta; p.in_count = a->in_count; p.in_act = a->in_act;
    p.in_slope = a->in_slope; p.minmax = a->in_min != nullptr;
    p.epi = a->epilogue; p.bias = a->bias; p.bias2 = a->bias2; p.out_act = a->out_act; p.residual = a->residual; p.out = a->out;
    p.out2 = a->out2; p.h = a->h; p.z = a->z; p.out_stats = a->out_stats;
    p.M = (int)M; p.N = a->n_pad; p.K = K; p.cout = a->cout; p.pts_per_sample = a->N;
    CUtensorMap maps[4], mw_hi, mw_lo;
    int rc;
    for (int s = 0; s < 3; ++s) {
        const float* src = a->in[s] ? a->in[s] : a->in[0];
        const int ch = a->in[s] ? a->in_channels[s] : a->in_channels[0];
        p.seg_kb[s] = a->in[s] ? a->in_channels[s] / kTcKB : 0;
        if ((rc = tc_make_map(&maps[s], src, M, ch, ch, kTcM))) return rc;
    }
    if ((rc = tc_make_map(&maps[3], a->in_min ? a->in_min : a->in[0], M, a->in_channels[0], a->in_channels[0], kTcM))) return rc;
    if ((rc = tc_make_map(&mw_hi, a->w_hi, a->n_pad, K, K, a->n_pad)) || (rc = tc_make_map(&mw_lo, a->w_lo, a->n_pad, K, K, a->n_pad))) return rc;
    const size_t stage = (size_t)(a->in_min ? 3 : 2) * kTcABytes + (size_t)2 * a->n_pad * kTcKB * 4;
    const size_t fixed = (size_t)(2 * K + 2 * a->n_pad) * sizeof(float) + 1024 + 64;
    // ring depth: deep enough to overlap TMA with the MMA, shallow enough that 3 CTAs share an SM (their phases --
    // load, transform, MMA, epilogue -- then overlap across CTAs); env PVRAFT_TC_STAGES overrides for experiments
    int stages = (int)(((size_t)kSmemBudget / 3 - fixed) / stage);
    stages = stages < 1 ? 1 : (stages > kTcMaxStages ? kTcMaxStages : stages);
    if (stages > K / kTcKB) stages = K / kTcKB;
    if (const char* e = getenv("PVRAFT_TC_STAGES")) { const int v = atoi(e); if (v >= 1 && v <= kTcMaxStages) stages = v; }
    p.stages = stages;
    const size_t smem = stages * stage + fixed;
    if ((rc = opt_in_smem(k_tc_linear, smem))) return rc;
    k_tc_linear<<<(unsigned)((M + kTcM - 1) / kTcM), kTcThreads, smem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], mw_hi, mw_lo, p);
    return check_launch("tc_linear");
}
