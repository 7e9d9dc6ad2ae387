// Per-point linear layers on the tcgen05 tensor cores, fp32-accurate (3xTF32), with the GroupNorm / activation
// prologue and the bias / activation / residual / GroupNorm-statistics / GRU-gate epilogues fused around the MMA.
//
//   out[M x N] = epilogue( prologue(A)[M x K] . W[N x K]^T )          M = B*Npts points, N = cout, K = cin
//
// Persistent kernel, one CTA (18 warps) per SM, tiles of 128 consecutive points (= 128 TMEM lanes), 32-channel k-blocks:
//   warp 0       TMA producer: raw fp32 activation boxes [128 x 32] (SWIZZLE_128B; up to three source tensors concatenated
//                along K, e.g. [h | inp | motion] for the GRU) into a ring of 2..6 stages; the pre-split weight boxes
//                W_hi, W_lo once per CTA when they fit next to the ring, else with every k-block
//   warps 2-9    transform, two groups on alternate k-blocks: every 16-byte chunk of the raw box gets the folded GroupNorm
//                affine + activation of its channels (optionally choosing the max or the min input by the sign of the
//                scale), is split into hi = tf32(x), lo = tf32(x - hi) and written in place / next to it AT THE SAME swizzled
//                offset (the split is elementwise); fence.proxy.async; one mbarrier arrival per warp
//   warp 1       MMA issuer (the warp walks the loop, one elected lane issues): A_hi.W_hi + A_lo.W_hi + A_hi.W_lo =
//                3 x 4 tcgen05.mma.kind::tf32 (K = 8 each) per k-block into one of two TMEM accumulators
//   warps 10-17  epilogue, two per TMEM lane quadrant, one accumulator behind the MMA: tcgen05.ld -> bias / activation /
//                residual -> transpose through shared memory -> coalesced stores; GroupNorm (sum, sum^2) of the output
//                combined across the warps into one double atomic per (group, moment) and tile; ConvGRU-gate, cat-tail and
//                flow-head (64 -> 3 + RAFT coordinate update) variants
// Launched with programmatic stream serialization: the prologue overlaps the previous kernel's tail (griddepcontrol.wait
// precedes the first global read).  Replaces the k_linear / k_gru / k_corrfeat / k_flowout CUDA-core kernels whenever
// Npts % 128 == 0 and every source has a multiple of 32 channels.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace pvraft {

constexpr int kTcThreads = 576;   // TMA producer | MMA | 8 transform warps (two groups) | 8 epilogue warps
constexpr int kTcM = 128, kTcKB = 32;
constexpr int kTcABytes = kTcM * kTcKB * 4;   // 16 KB: one activation box
constexpr int kTcMaxStages = 6;

// In-kernel timeline of CTA 0 (how the pipeline was tuned, tools/tc_clock.py): compiled in only with -DPVRAFT_TC_TIMELINE
// and then recorded when PVRAFT_TC_DBG=1; the product build carries none of it.
#ifdef PVRAFT_TC_TIMELINE
__device__ unsigned long long g_tc_clock[64];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_MARK(cond, slot) do { if (clk && (cond)) g_tc_clock[(slot)] = gtimer(); } while (0)
#else
#define TC_MARK(cond, slot) do { } while (0)
#endif

enum TcEpilogue { TC_EPI_PLAIN = 0, TC_EPI_GRU_ZR = 1, TC_EPI_GRU_Q = 2, TC_EPI_FLOW = 3 };

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
    const float* src[3];      // activation sources [M, 32*seg_kb[i]] row-major
    const float* src_min;     // per-channel minima paired with src[0] (minmax prologue)
    int seg_kb[3];            // k-blocks contributed by each activation source
    int stages;               // depth of the shared-memory ring (1..4)
    int w_resident;           // 1: all weight boxes are loaded once per CTA and stay in shared memory
    int settled;              // 1: weights / biases may be read before griddepcontrol.wait
    int* done;                // [B] or null: every epilogue warp adds 1 per finished tile of the sample (release)
    const int* wait_on;       // [B] or null: the `done` counters of the previous launch, which produced this layer's inputs.
                              // Non-null: the grid-wide dependency wait is replaced by per-sample waits in the TMA producer
    unsigned wait_target;     // 8 * tiles per sample
    int dbg;                  // 1: CTA 0 records a globaltimer timeline into g_tc_clock
    int gn_kb;                // k-blocks (from the start: source 0) that go through the GroupNorm prologue
    int out_ld;               // row stride of `out` in floats (cout, or cout + 3 with a tail)
    const float* tail;        // [M,3] copied into output columns cout..cout+2, or nullptr
    const float *w3, *b3, *coords1, *coords2;   // FLOW epilogue
    float *coords2_out, *flow_out, *flow_user;
    const int32_t* row_map;
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
// round-to-nearest (ties away from zero) to the 10-bit TF32 mantissa with two full-rate integer ops; identical to
// cvt.rna.tf32.f32 for finite values (the conversion instruction runs at a fraction of the ALU rate)
__device__ __forceinline__ float tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
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
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// thread-per-row values of a [32 rows x 16 columns] block -> shared-memory transpose -> stores of 8 rows x 64 B per instruction
// (thread-per-row stores would scatter 32 half-filled sectors per instruction).  stg: this warp's [32][20] staging tile.
constexpr int kTcPitch16 = 20;
__device__ __forceinline__ void stage_store16(float* __restrict__ stg, int lane, const float (&y)[16], float* __restrict__ gbase, int ld) {
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(stg + lane * kTcPitch16 + q * 4) = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
    __syncwarp();
    const int rsub = lane >> 2, cq = lane & 3;
    float4 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = *reinterpret_cast<const float4*>(stg + (j * 8 + rsub) * kTcPitch16 + cq * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(gbase + (size_t)(j * 8 + rsub) * ld + cq * 4) = o[j];
}
// the inverse: a [32 rows x 16 columns] block of a row-major global tensor, loaded as 8 rows x 64 B per instruction and handed
// to the thread that owns each row (thread-per-row loads would touch 32 sectors per instruction)
// COHERENT: the operand may have been written by the previous launch while this one was already running (chained launches):
// ld.global.cg (L2) instead of the non-coherent path.
template <bool COHERENT = false>
__device__ __forceinline__ void stage_load16(float* __restrict__ stg, int lane, const float* gbase, int ld, float (&x)[16]) {
    const int rsub = lane >> 2, cq = lane & 3;
    float4 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4* src = reinterpret_cast<const float4*>(gbase + (size_t)(j * 8 + rsub) * ld + cq * 4);
        o[j] = COHERENT ? __ldcg(src) : __ldg(src);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(stg + (j * 8 + rsub) * kTcPitch16 + cq * 4) = o[j];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(stg + lane * kTcPitch16 + q * 4);
        x[q * 4 + 0] = v.x; x[q * 4 + 1] = v.y; x[q * 4 + 2] = v.z; x[q * 4 + 3] = v.w;
    }
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const int* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(int* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ bool telect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float tsigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// epilogue of one accumulator buffer: thread = point (TMEM lane)
__device__ __forceinline__ void tc_epilogue(const TcParams& p, unsigned tacc, int quad, int half, int lane, int row0, int sample,
                                            const float* __restrict__ s_bias, float* __restrict__ s_stage, float* __restrict__ s_part) {
    const int row = row0 + quad * 32 + lane;
    const int gsz = p.cout / PVRAFT_GN_GROUPS;
    const unsigned tl = tacc + ((unsigned)(quad * 32) << 16);
    if (p.epi == TC_EPI_PLAIN) {
        // Every tile is full (M is a multiple of 128).  This code runs on one warp per scheduler, so it is written for
        // latency: no per-element branches, addresses hoisted, loads batched ahead of their consumers.
        const bool vec = (p.out_ld & 3) == 0;
        const ActCoef oact = act_coef(p.out_act, 0.f);
        float* stg = s_stage + (size_t)(half * 4 + quad) * 32 * 36;   // this warp's [32 rows][36] staging tile
        const int rsub = lane >> 3, cq = lane & 7;
        float* obase = p.out + (size_t)(row0 + quad * 32 + rsub) * p.out_ld + cq * 4;
        const size_t ostep = (size_t)4 * p.out_ld;
        const bool want_stats = p.out_stats != nullptr;
        for (int c0 = half * 32; c0 < p.N; c0 += 64) {
            unsigned v[32];
            tmem_ld32(tl + (unsigned)c0, v);
            float y[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + q * 4);
                y[q * 4 + 0] = apply_act(__uint_as_float(v[q * 4 + 0]) + bv.x, oact);
                y[q * 4 + 1] = apply_act(__uint_as_float(v[q * 4 + 1]) + bv.y, oact);
                y[q * 4 + 2] = apply_act(__uint_as_float(v[q * 4 + 2]) + bv.z, oact);
                y[q * 4 + 3] = apply_act(__uint_as_float(v[q * 4 + 3]) + bv.w, oact);
            }
            if (p.residual != nullptr) {   // (rare: FlotRefine.fc) thread-per-row loads, before the statistics
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < p.cout) y[i] += __ldg(p.residual + (size_t)row * p.cout + c0 + i);
            }
            if (p.tail != nullptr && c0 + 32 == p.out_ld) {   // last three columns of the row carry the tail (cat([out, flow]))
                const float* tp = p.tail + (size_t)row * 3;
                y[29] = __ldg(tp); y[30] = __ldg(tp + 1); y[31] = __ldg(tp + 2);
            }
            if (vec) {
                // transpose through shared memory so that a store instruction writes 4 rows x 128 contiguous bytes
                // (thread-per-row stores would scatter 32 half-filled sectors per instruction)
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(stg + lane * 36 + q * 4) = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
                __syncwarp();
                float4 t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(stg + (i * 4 + rsub) * 36 + cq * 4);
                if (c0 + cq * 4 < p.out_ld) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(obase + i * ostep + c0) = t[i];
                }
                if (want_stats) {
                    // lane = column: (sum, sum^2) of this warp's 32 rows, read down the staged tile (conflict-free)
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float a = stg[r * 36 + lane];
                        s1 += a;
                        s2 = fmaf(a, a, s2);
                    }
                    *reinterpret_cast<float2*>(s_part + (size_t)(quad * 128 + c0 + lane) * 2) = make_float2(s1, s2);
                }
            } else {
                float* o = p.out + (size_t)row * p.out_ld + c0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (c0 + i < p.out_ld) o[i] = y[i];
                }
            }
        }
        if (want_stats) {
            // combine the 4 epilogue warps and the columns of each GroupNorm group: thread t owns column t
            asm volatile("bar.sync 3, 256;" ::: "memory");
            const int t = (half * 4 + quad) * 32 + lane;   // 0..255; threads 0..N-1 own one column each
            double a1 = 0.0, a2 = 0.0;
            if (t < p.N) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float2 pr = *reinterpret_cast<const float2*>(s_part + (size_t)(w * 128 + t) * 2);
                    a1 += (double)pr.x;
                    a2 += (double)pr.y;
                }
            }
            if ((gsz & (gsz - 1)) == 0 && gsz <= 32) {   // groups are aligned runs of gsz lanes
                for (int o = gsz >> 1; o > 0; o >>= 1) {
                    a1 += __shfl_xor_sync(kFull, a1, o);
                    a2 += __shfl_xor_sync(kFull, a2, o);
                }
                if ((t & (gsz - 1)) == 0 && t < p.cout) {
                    atomicAdd(p.out_stats + (size_t)sample * 16 + (t / gsz) * 2 + 0, a1);
                    atomicAdd(p.out_stats + (size_t)sample * 16 + (t / gsz) * 2 + 1, a2);
                }
            } else if (t < p.cout) {
                atomicAdd(p.out_stats + (size_t)sample * 16 + (t / gsz) * 2 + 0, a1);
                atomicAdd(p.out_stats + (size_t)sample * 16 + (t / gsz) * 2 + 1, a2);
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
        }
    } else if (p.epi == TC_EPI_GRU_ZR) {
        // accumulator columns 0..63 = z pre-activation, 64..127 = r pre-activation (model/update.py:34-35)
        float* stg = s_stage + (size_t)(half * 4 + quad) * 32 * kTcPitch16;
        for (int c = half * 32; c < half * 32 + 32; c += 16) {
            unsigned vz[16], vr[16];
            tmem_ld16(tl + (unsigned)c, vz);
            tmem_ld16(tl + (unsigned)(64 + c), vr);
            float z[16], rh[16], hh[16];
            stage_load16(stg, lane, p.h + (size_t)(row0 + quad * 32) * 64 + c, 64, hh);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 hv = make_float4(hh[q * 4], hh[q * 4 + 1], hh[q * 4 + 2], hh[q * 4 + 3]);
                float4 bz = *reinterpret_cast<const float4*>(s_bias + c + q * 4);
                float4 br = *reinterpret_cast<const float4*>(s_bias + p.N + c + q * 4);
                if (p.residual != nullptr) {   // per-point pre-activation term [M,128] = [z | r] (the constant context part)
                    const float4 az = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)row * 128 + c + q * 4));
                    const float4 ar = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)row * 128 + 64 + c + q * 4));
                    bz.x += az.x; bz.y += az.y; bz.z += az.z; bz.w += az.w;
                    br.x += ar.x; br.y += ar.y; br.z += ar.z; br.w += ar.w;
                }
                z[q * 4 + 0] = tsigmoid(__uint_as_float(vz[q * 4 + 0]) + bz.x); z[q * 4 + 1] = tsigmoid(__uint_as_float(vz[q * 4 + 1]) + bz.y);
                z[q * 4 + 2] = tsigmoid(__uint_as_float(vz[q * 4 + 2]) + bz.z); z[q * 4 + 3] = tsigmoid(__uint_as_float(vz[q * 4 + 3]) + bz.w);
                rh[q * 4 + 0] = tsigmoid(__uint_as_float(vr[q * 4 + 0]) + br.x) * hv.x; rh[q * 4 + 1] = tsigmoid(__uint_as_float(vr[q * 4 + 1]) + br.y) * hv.y;
                rh[q * 4 + 2] = tsigmoid(__uint_as_float(vr[q * 4 + 2]) + br.z) * hv.z; rh[q * 4 + 3] = tsigmoid(__uint_as_float(vr[q * 4 + 3]) + br.w) * hv.w;
            }
            stage_store16(stg, lane, z, p.out + (size_t)(row0 + quad * 32) * 64 + c, 64);
            stage_store16(stg, lane, rh, p.out2 + (size_t)(row0 + quad * 32) * 64 + c, 64);
        }
    } else if (p.epi == TC_EPI_FLOW) {
        // y = relu(acc + b) (flow_head.out_conv.0/1); delta = w3 . y + b3 (out_conv.2); RAFT update of the coordinates.
        // The two warps of a lane quadrant hold 32 of the 64 columns each: half 1 parks its partial dot products.
        const float* s_w3 = s_part + 512;   // [3][64], staged at kernel start; s_part[0..511] = [128 rows][4] exchange
        unsigned v[32];
        const int c0 = half * 32;
        tmem_ld32(tl + (unsigned)c0, v);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + q * 4);
            const float4 wa = *reinterpret_cast<const float4*>(s_w3 + 0 * 64 + c0 + q * 4);
            const float4 wb = *reinterpret_cast<const float4*>(s_w3 + 1 * 64 + c0 + q * 4);
            const float4 wc = *reinterpret_cast<const float4*>(s_w3 + 2 * 64 + c0 + q * 4);
            const float y0 = fmaxf(__uint_as_float(v[q * 4 + 0]) + bv.x, 0.f), y1 = fmaxf(__uint_as_float(v[q * 4 + 1]) + bv.y, 0.f);
            const float y2 = fmaxf(__uint_as_float(v[q * 4 + 2]) + bv.z, 0.f), y3 = fmaxf(__uint_as_float(v[q * 4 + 3]) + bv.w, 0.f);
            d0 = fmaf(wa.w, y3, fmaf(wa.z, y2, fmaf(wa.y, y1, fmaf(wa.x, y0, d0))));
            d1 = fmaf(wb.w, y3, fmaf(wb.z, y2, fmaf(wb.y, y1, fmaf(wb.x, y0, d1))));
            d2 = fmaf(wc.w, y3, fmaf(wc.z, y2, fmaf(wc.y, y1, fmaf(wc.x, y0, d2))));
        }
        float* xch = s_part + (size_t)(quad * 32 + lane) * 4;
        if (half == 1) *reinterpret_cast<float4*>(xch) = make_float4(d0, d1, d2, 0.f);
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (half == 0) {
            const float4 o = *reinterpret_cast<const float4*>(xch);
            const float dd[3] = {(d0 + o.x) + s_bias[p.N + 0], (d1 + o.y) + s_bias[p.N + 1], (d2 + o.z) + s_bias[p.N + 2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const size_t g = (size_t)row * 3 + k;
                p.out[g] = dd[k];
                if (p.coords2_out != nullptr) {
                    const float c2 = p.coords2[g] + dd[k];   // RAFTSceneFlow.py:45
                    p.coords2_out[g] = c2;
                    if (p.flow_out != nullptr) {
                        const float fl = c2 - __ldg(p.coords1 + g);   // RAFTSceneFlow.py:46
                        p.flow_out[g] = fl;
                        if (p.flow_user != nullptr) p.flow_user[(size_t)__ldg(p.row_map + row) * 3 + k] = fl;
                    }
                }
            }
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");   // the exchange buffer is free for the next tile
    } else {
        // q = tanh(acc + b); h' = (1 - z) h + z q   (model/update.py:37-39)
        float* stg = s_stage + (size_t)(half * 4 + quad) * 32 * kTcPitch16;
        for (int c = half * 32; c < half * 32 + 32; c += 16) {
            unsigned vq[16];
            tmem_ld16(tl + (unsigned)c, vq);
            float o[16], hh[16], zz[16];
            stage_load16(stg, lane, p.h + (size_t)(row0 + quad * 32) * 64 + c, 64, hh);
            stage_load16<true>(stg, lane, p.z + (size_t)(row0 + quad * 32) * 64 + c, 64, zz);   // z comes from the launch before
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 hv = make_float4(hh[q * 4], hh[q * 4 + 1], hh[q * 4 + 2], hh[q * 4 + 3]);
                const float4 zv = make_float4(zz[q * 4], zz[q * 4 + 1], zz[q * 4 + 2], zz[q * 4 + 3]);
                float4 bq = *reinterpret_cast<const float4*>(s_bias + c + q * 4);
                if (p.residual != nullptr) {   // per-point pre-activation term [M,64]
                    const float4 aq = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)row * 64 + c + q * 4));
                    bq.x += aq.x; bq.y += aq.y; bq.z += aq.z; bq.w += aq.w;
                }
                o[q * 4 + 0] = (1.f - zv.x) * hv.x + zv.x * tanhf(__uint_as_float(vq[q * 4 + 0]) + bq.x);
                o[q * 4 + 1] = (1.f - zv.y) * hv.y + zv.y * tanhf(__uint_as_float(vq[q * 4 + 1]) + bq.y);
                o[q * 4 + 2] = (1.f - zv.z) * hv.z + zv.z * tanhf(__uint_as_float(vq[q * 4 + 2]) + bq.z);
                o[q * 4 + 3] = (1.f - zv.w) * hv.w + zv.w * tanhf(__uint_as_float(vq[q * 4 + 3]) + bq.w);
            }
            stage_store16(stg, lane, o, p.out + (size_t)(row0 + quad * 32) * 64 + c, 64);
        }
    }
}

// A CTA walks tiles blockIdx.x, +gridDim.x, ...; the operand ring and the two accumulators run across tile boundaries.
constexpr int kTcXform = 256;   // transform threads

// Position of one pipeline role in the (tile, k-block) walk and in the operand ring; advanced without divisions (a
// runtime integer division is a ~100-cycle dependent chain, paid per step by warps that have nothing to hide it behind)
struct TcCursor {
    int ti = 0, kb = 0, s = 0;
    unsigned phase = 0;
    __device__ __forceinline__ void next(int num_kb, int S) {
        if (++kb == num_kb) { kb = 0; ++ti; }
        if (++s == S) { s = 0; phase ^= 1u; }
    }
};

__global__ void __launch_bounds__(kTcThreads, 1)
k_tc_linear(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
            const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_min, const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared array: an integer round trip would lose the address space
    // and turn every shared-memory access below into a generic LD/ST
    unsigned char* tiles = smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    // stage layout: [A hi 16K][A lo 16K][W hi N*128][W lo N*128, only when the weights are streamed];
    // resident weights live behind the ring as num_kb x [W hi][W lo]
    const int w_bytes = p.N * kTcKB * 4;
    const int w_off = 2 * kTcABytes;
    const int stage_bytes = w_off + (p.w_resident ? 0 : 2 * w_bytes);
    const int S = p.stages;
    const int num_kb = p.K / kTcKB;
    unsigned char* w_res = tiles + (size_t)S * stage_bytes;
    float* s_scale = reinterpret_cast<float*>(w_res + (p.w_resident ? (size_t)num_kb * 2 * w_bytes : 0));   // [K]
    float* s_bias = s_scale + 4 * p.K;                                            // ([2 groups][scale K | shift K] above)                                                // [2 * N]
    float* s_estage = s_bias + 2 * p.N;                                           // [4 warps][32][36] epilogue staging
    float* s_part = s_estage + (p.epi == TC_EPI_PLAIN ? 8 * 32 * 36 : (p.epi == TC_EPI_FLOW ? 0 : 8 * 32 * kTcPitch16));   // staging tiles by epilogue                                       // [4 warps][128 columns][2]
    __shared__ __align__(8) unsigned long long s_full[kTcMaxStages], s_ready[kTcMaxStages], s_empty[kTcMaxStages];
    __shared__ __align__(8) unsigned long long s_acc_full[2], s_acc_empty[2], s_w_full;
    __shared__ unsigned s_tmem_base;
    const int warp = warp_id(), lane = lane_id();
    const int n_tiles = (p.M + kTcM - 1) / kTcM;
    const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int total_steps = my_tiles * num_kb;
    const unsigned acc_cols = p.N <= 32 ? 32u : p.N <= 64 ? 64u : 128u;   // columns per accumulator buffer
    const unsigned tmem_cols = acc_cols * 2;
#ifdef PVRAFT_TC_TIMELINE
    const bool clk = p.dbg && blockIdx.x == 0;
#endif
    TC_MARK(threadIdx.x == 0, 0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) { tmbar_init(&s_full[s], 1); tmbar_init(&s_ready[s], kTcXform / 64); tmbar_init(&s_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tmbar_init(&s_acc_full[a], 1); tmbar_init(&s_acc_empty[a], 8); }
        tmbar_init(&s_w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tsu32(&s_tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Programmatic dependent launch: everything above (shared-memory carve-up, barrier init, TMEM allocation, descriptor
    // prefetch) may overlap the tail of the previous kernel on this stream.  The layer's PARAMETERS (hi/lo weights, biases)
    // are fetched in that window too when the caller vouches that they are settled (p.settled: last written several
    // launches ago -- the steady state of a forward, whose weights are split once): the SMs that finished the previous
    // kernel early then hold their weights when the dependency resolves.  Every read of an ACTIVATION is below the wait.
    auto load_weights = [&]() {   // the whole weight matrix (hi and lo) once per CTA
        tmbar_expect_tx(&s_w_full, (unsigned)(num_kb * 2 * w_bytes));
        for (int kb = 0; kb < num_kb; ++kb) {
            ttma_load_2d(w_res + (size_t)kb * 2 * w_bytes, &map_w_hi, &s_w_full, kb * kTcKB, 0);
            ttma_load_2d(w_res + (size_t)kb * 2 * w_bytes + w_bytes, &map_w_lo, &s_w_full, kb * kTcKB, 0);
        }
    };
    auto load_bias = [&]() {
        for (int c = threadIdx.x - 320; c < p.N; c += 256) {
            s_bias[c] = (p.bias != nullptr && c < p.cout) ? __ldg(p.bias + c) : 0.f;
            s_bias[p.N + c] = (p.bias2 != nullptr && c < p.cout) ? __ldg(p.bias2 + c) : 0.f;
        }
        if (p.epi == TC_EPI_FLOW) {   // out_conv.2: weights behind the exchange buffer, bias in the bias2 slots
            for (int i = threadIdx.x - 320; i < 192; i += 256) s_part[512 + i] = __ldg(p.w3 + i);
            if (threadIdx.x - 320 < 3) s_bias[p.N + threadIdx.x - 320] = __ldg(p.b3 + threadIdx.x - 320);
        }
    };
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.settled) {
        __syncthreads();   // barrier init visible to the TMA issuer
        if (warp == 0 && lane == 0 && p.w_resident) load_weights();
        if (warp >= 10) load_bias();
    }
    // Chained on the previous launch (p.wait_on): no grid-wide wait -- the TMA producer waits per sample on that launch's
    // `done` counters instead, so this CTA starts on the tiles whose inputs are complete while the stragglers of the
    // previous launch still run on other SMs.  Everything else this kernel reads is older than the previous launch,
    // which itself only signals after its own dependencies resolved.
    if (p.wait_on == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!p.settled && warp >= 10) load_bias();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem_base;

    if (warp == 0) {
        // ===== TMA producer: raw activation boxes (and the weights) =====
        if (lane == 0) {
            if (p.w_resident && !p.settled) load_weights();
            const unsigned tx = (unsigned)(kTcABytes * (p.minmax ? 2 : 1) + (p.w_resident ? 0 : 2 * w_bytes));
            TcCursor cw;
            int ready_sample = -1;
            for (int step = 0; step < total_steps; ++step, cw.next(num_kb, S)) {
                const int s = cw.s, kb = cw.kb;
                const int row0 = (blockIdx.x + cw.ti * gridDim.x) * kTcM;
                if (p.wait_on != nullptr && kb == 0) {
                    const int sample = row0 / p.pts_per_sample;
                    if (sample != ready_sample) {   // all tiles of this sample (rows and GroupNorm sums) are out
                        while (ld_acquire_gpu(p.wait_on + sample) < p.wait_target) __nanosleep(40);
                        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> the TMA reads below
                        ready_sample = sample;
                    }
                }
                tmbar_wait(&s_empty[s], cw.phase ^ 1u);   // the MMAs that read this stage last time have retired
                unsigned char* st = tiles + (size_t)s * stage_bytes;
                tmbar_expect_tx(&s_full[s], tx);
                // raw fp32 rows of the source this k-block belongs to land where the hi operand will be (the transform works
                // in place); the min array of a (max, min) pair lands in the lo half
                if (kb < p.seg_kb[0]) ttma_load_2d(st, &map_a0, &s_full[s], kb * kTcKB, row0);
                else if (kb < p.seg_kb[0] + p.seg_kb[1]) ttma_load_2d(st, &map_a1, &s_full[s], (kb - p.seg_kb[0]) * kTcKB, row0);
                else ttma_load_2d(st, &map_a2, &s_full[s], (kb - p.seg_kb[0] - p.seg_kb[1]) * kTcKB, row0);
                if (p.minmax) ttma_load_2d(st + kTcABytes, &map_min, &s_full[s], kb * kTcKB, row0);
                if (!p.w_resident) {
                    ttma_load_2d(st + w_off, &map_w_hi, &s_full[s], kb * kTcKB, 0);
                    ttma_load_2d(st + w_off + w_bytes, &map_w_lo, &s_full[s], kb * kTcKB, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The whole warp walks the loop (convergent control flow keeps the descriptor arithmetic on the uniform datapath)
        // and one elected lane issues; under `if (lane == 0)` the compiler wraps every UTCHMMA in an election loop.
        {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(p.N >> 3) << 17) | ((unsigned)(kTcM >> 4) << 24);
            if (p.w_resident) tmbar_wait(&s_w_full, 0u);
            int step = 0;
            TcCursor cm;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                const unsigned acc_phase = (unsigned)(ti >> 1) & 1u;
                tmbar_wait(&s_acc_empty[acc], acc_phase ^ 1u);   // the epilogue has drained this accumulator buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tacc = tmem + (unsigned)acc * acc_cols;
                for (int kb = 0; kb < num_kb; ++kb, ++step, cm.next(num_kb, S)) {
                    const int s = cm.s;
                    const unsigned phase = cm.phase;
                    tmbar_wait(&s_ready[s], phase);               // transformed activations are in place
                    tmbar_wait(&s_full[s], phase);                // (already complete: the transform waited on it)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    unsigned char* st = tiles + (size_t)s * stage_bytes;
                    const unsigned long long a_hi = tumma_desc(st), a_lo = tumma_desc(st + kTcABytes);
                    const unsigned char* wb = p.w_resident ? w_res + (size_t)kb * 2 * w_bytes : st + w_off;
                    const unsigned long long b_hi = tumma_desc(wb), b_lo = tumma_desc(wb + w_bytes);
                    if (telect_one()) {
#pragma unroll
                        for (int k = 0; k < kTcKB / 8; ++k) {
                            const unsigned long long off = (unsigned long long)(k * 2);
                            tumma_tf32(tacc, a_hi + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
                            tumma_tf32(tacc, a_lo + off, b_hi + off, idesc, 1u);
                            tumma_tf32(tacc, a_hi + off, b_lo + off, idesc, 1u);
                        }
                        tumma_commit(&s_empty[s]);
                        if (kb == num_kb - 1) tumma_commit(&s_acc_full[acc]);
                        TC_MARK(step < 8, 16 + step);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 10) {
        // ===== transform: raw fp32 box -> GroupNorm affine + activation -> tf32 hi/lo operand tiles, in place =====
        // Two groups of four warps take alternate k-blocks, so the load -> math -> store -> fence chain of one step
        // overlaps the next step's; each group keeps its own copy of the per-sample GroupNorm table.
        const int grp = (warp - 2) >> 2;
        const int t = (threadIdx.x - 64) & 127;   // chunk c = t + 128*i, i < 8, of the 1024 16-byte chunks of a k-block
        float* g_scale = s_scale + grp * 2 * p.K;
        float* g_shift = g_scale + p.K;
        const ActCoef iact = act_coef(p.in_act, p.in_slope);
        TcCursor cp;
        const int tiles_per_sample = p.pts_per_sample / kTcM;
        int table_first = -1, table_end = -1;   // tile range [first, end) of the sample whose table this group holds
        if (grp == 1) cp.next(num_kb, S);
        for (int step = grp; step < total_steps; step += 2) {
            const int kb = cp.kb, s = cp.s;
            auto gn_table = [&]() {
                const int tile = blockIdx.x + cp.ti * gridDim.x;
                if (tile < table_first || tile >= table_end) {   // folded GroupNorm affine of every input channel of this sample
                    const int sample = tile / tiles_per_sample;
                    table_first = sample * tiles_per_sample;
                    table_end = table_first + tiles_per_sample;
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the group is done with the previous table
                    const int gn_k = p.gn_kb * kTcKB, gsz = gn_k / PVRAFT_GN_GROUPS;
                    for (int k = t; k < gn_k; k += 128) {
                        const double* sp = p.in_stats + (size_t)sample * 16 + (k / gsz) * 2;
                        const double st[2] = {__ldcg(sp), __ldcg(sp + 1)};   // (written by the launch before: L2, not the nc path)
                        const GnAffine af = gn_affine(st, p.in_count, __ldg(p.in_gamma + k), __ldg(p.in_beta + k));
                        g_scale[k] = af.scale;
                        g_shift[k] = af.shift;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                }
            };
            // the sums are complete once the previous launch is: before the box wait normally (the table overlaps the TMA
            // latency), after it when chained (the producer issued this box only after the sample's `done` count was reached)
            if (p.in_stats != nullptr && p.wait_on == nullptr) gn_table();
            tmbar_wait(&s_full[s], cp.phase);   // the raw box(es) of this k-block have landed
            if (p.in_stats != nullptr && p.wait_on != nullptr) gn_table();
            unsigned char* st = tiles + (size_t)s * stage_bytes;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = t + i * 128, r = c >> 3, lc = c & 7;
                // K-major SWIZZLE_128B: 16-byte chunk lc of row r lives at chunk (lc ^ (r & 7)) of the row's 128 bytes
                const int off = r * 128 + ((lc ^ (r & 7)) << 4);
                float4 x = *reinterpret_cast<const float4*>(st + off);
                if (kb < p.gn_kb) {
                    const int k = kb * kTcKB + lc * 4;
                    const float4 sc = *reinterpret_cast<const float4*>(g_scale + k);
                    const float4 sh = *reinterpret_cast<const float4*>(g_shift + k);
                    if (p.minmax) {
                        const float4 mn = *reinterpret_cast<const float4*>(st + kTcABytes + off);
                        x.x = sc.x < 0.f ? mn.x : x.x; x.y = sc.y < 0.f ? mn.y : x.y;
                        x.z = sc.z < 0.f ? mn.z : x.z; x.w = sc.w < 0.f ? mn.w : x.w;
                    }
                    x.x = apply_act(fmaf(x.x, sc.x, sh.x), iact);
                    x.y = apply_act(fmaf(x.y, sc.y, sh.y), iact);
                    x.z = apply_act(fmaf(x.z, sc.z, sh.z), iact);
                    x.w = apply_act(fmaf(x.w, sc.w, sh.w), iact);
                }
                float4 hi, lo;
                hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
                lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
                *reinterpret_cast<float4*>(st + off) = hi;   // in place: raw -> hi; the lo half of the stage held the min array
                *reinterpret_cast<float4*>(st + kTcABytes + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) tmbar_arrive(&s_ready[s]);   // one arrival per warp of the group
            TC_MARK(t == 0 && step < 8, 8 + step);
            cp.next(num_kb, S);
            cp.next(num_kb, S);
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global, one accumulator buffer behind the MMA =====
        const int quad = warp & 3;          // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int half = (warp - 10) >> 2;  // two warps per lane quadrant: 32-column chunks c0 = 32*half, +64, ...
        // Completion signal of a tile (p.done: one release-add per epilogue warp and tile).  The release has to wait for the
        // warp's outstanding stores, ~1 us if issued right behind them -- exposed, because this warp is alone on its
        // scheduler.  So the signal of tile i is sent when the accumulator of tile i+1 arrives (its stores have long
        // landed: the fence is free), and only the last tile signals at once.  A consumer loses nothing: it cannot run on
        // this SM before this CTA exits, and samples that complete in an earlier round are not needed sooner.
        int unsignalled = -1;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int acc = ti & 1;
            const unsigned acc_phase = (unsigned)(ti >> 1) & 1u;
            const int row0 = (blockIdx.x + ti * gridDim.x) * kTcM;
            tmbar_wait(&s_acc_full[acc], acc_phase);
            if (unsignalled >= 0 && lane == 0) red_release_gpu(p.done + unsignalled);
            TC_MARK(threadIdx.x == 320 && ti < 4, 24 + ti);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int sample = row0 / p.pts_per_sample;
            tc_epilogue(p, tmem + (unsigned)acc * acc_cols, quad, half, lane, row0, sample, s_bias, s_estage, s_part);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();   // (also orders every lane's stores of this tile before lane 0's later release)
            if (lane == 0) tmbar_arrive(&s_acc_empty[acc]);   // one arrival per epilogue warp
            if (p.done != nullptr) unsignalled = sample;
            TC_MARK(threadIdx.x == 320 && ti < 4, 28 + ti);
        }
        if (unsignalled >= 0 && lane == 0) red_release_gpu(p.done + unsignalled);
    }
    __syncwarp();   // the producer / MMA roles run on one lane: re-converge those warps before the CTA-wide barrier
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
    TC_MARK(threadIdx.x == 64, 1);
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

#ifdef PVRAFT_TC_TIMELINE
extern "C" __attribute__((visibility("default"))) int pvraft_tc_debug_clock(unsigned long long* host64) {
    return (int)cudaMemcpyFromSymbol(host64, g_tc_clock, sizeof(unsigned long long) * 64);
}
#endif

extern "C" int pvraft_tc_weight_split(const float* w, int rows, int cols, int ld, int col0, int rows_pad, int cols_pad, float* hi,
                                      float* lo, void* stream) {
    if (!w || !hi || !lo || rows <= 0 || cols <= 0 || rows_pad < rows || cols_pad < cols) return fail(PVRAFT_ERR_BAD_ARG, "tc_weight_split: bad argument");
    const int n = rows_pad * cols_pad;
    k_weight_split<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, ld > 0 ? ld : cols, col0, rows_pad, cols_pad, hi, lo);
    return check_launch("tc_weight_split");
}

extern "C" int pvraft_tc_linear_fwd(const pvraft_tc_linear_args* a, void* stream) {
    if (!a || !a->in[0] || !a->w_hi || !a->w_lo || !a->out) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: null pointer");
    if (a->B <= 0 || a->N <= 0 || a->n_pad < 16 || a->n_pad > 128 || a->n_pad % 16) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: bad shape (n_pad=%d)", a->n_pad);
    if (a->N % kTcM) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: points per sample (%d) must be a multiple of 128", a->N);
    int K = 0;
    for (int s = 0; s < 3; ++s) {
        if (a->in[s] && (a->in_channels[s] <= 0 || a->in_channels[s] % kTcKB)) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: source %d has %d channels (multiple of 32 needed)", s, a->in_channels[s]);
        if (a->in[s]) K += a->in_channels[s];
    }
    if (K <= 0 || K > 512) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: K=%d", K);
    if (a->in_stats && (a->in_channels[0] % PVRAFT_GN_GROUPS || !a->in_gamma || !a->in_beta)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GroupNorm prologue needs gamma, beta and in_channels[0] %% 8 == 0");
    if (a->tail && (a->epilogue != TC_EPI_PLAIN || a->cout + 3 != a->n_pad || a->n_pad % 32 || a->residual || a->out_stats)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: a tail needs the plain epilogue and cout + 3 == n_pad (multiple of 32)");
    if (a->in_min && !a->in_stats) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: in_min needs the GroupNorm prologue");
    if (a->out_stats && (a->epilogue != TC_EPI_PLAIN || a->cout % PVRAFT_GN_GROUPS || (a->cout / PVRAFT_GN_GROUPS) % 4)) return fail(PVRAFT_ERR_UNSUPPORTED, "tc_linear: out_stats needs a GroupNorm group size that is a multiple of 4 (cout=%d)", a->cout);
    if (a->epilogue < TC_EPI_PLAIN || a->epilogue > TC_EPI_FLOW) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: unknown epilogue %d", a->epilogue);
    if ((a->epilogue == TC_EPI_GRU_ZR || a->epilogue == TC_EPI_GRU_Q) && (a->cout != 64 || !a->h || (!a->bias && !a->residual))) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GRU epilogues need cout=64, h and a bias or a pre-activation term");
    if (a->epilogue == TC_EPI_FLOW && (a->cout != 64 || a->n_pad != 64 || !a->bias || !a->w3 || !a->b3 || (a->coords2_out && !a->coords2) || (a->flow_out && (!a->coords2_out || !a->coords1)) || (a->flow_user && (!a->flow_out || !a->row_map)))) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: flow epilogue needs cout = n_pad = 64, bias, w3, b3 and consistent coordinate pointers");
    if (a->epilogue == TC_EPI_GRU_ZR && (a->n_pad != 128 || (!a->bias2 && !a->residual) || !a->out2)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GRU zr epilogue needs n_pad=128, bias2, out2");
    if (a->epilogue == TC_EPI_GRU_Q && (a->n_pad != 64 || !a->z)) return fail(PVRAFT_ERR_BAD_ARG, "tc_linear: GRU q epilogue needs n_pad=64 and z");
    const long long M = (long long)a->B * a->N;
    TcParams p{};
    p.in_stats = a->in_stats; p.in_gamma = a->in_gamma; p.in_beta = a->in_beta; p.in_count = a->in_count; p.in_act = a->in_act;
    p.in_slope = a->in_slope; p.minmax = a->in_min != nullptr;
    p.epi = a->epilogue; p.bias = a->bias; p.bias2 = a->bias2; p.out_act = a->out_act; p.residual = a->residual; p.out = a->out;
    p.out2 = a->out2; p.h = a->h; p.z = a->z; p.out_stats = a->out_stats;
    p.M = (int)M; p.N = a->n_pad; p.K = K; p.cout = a->cout; p.pts_per_sample = a->N;
    CUtensorMap mw_hi, mw_lo;
    int rc;
    for (int s = 0; s < 3; ++s) {
        p.src[s] = a->in[s];
        p.seg_kb[s] = a->in[s] ? a->in_channels[s] / kTcKB : 0;
    }
    p.src_min = a->in_min;
    p.gn_kb = a->in_stats ? a->in_channels[0] / kTcKB : 0;
    p.tail = a->tail;
    p.w3 = a->w3; p.b3 = a->b3; p.coords1 = a->coords1; p.coords2 = a->coords2; p.coords2_out = a->coords2_out; p.flow_out = a->flow_out; p.flow_user = a->flow_user; p.row_map = a->row_map;
    p.out_ld = a->tail ? a->cout + 3 : a->cout;
    p.settled = a->params_settled ? 1 : 0;
    p.done = a->done;
    p.wait_target = 8u * (unsigned)(a->N / kTcM);
    if ((rc = tc_make_map(&mw_hi, a->w_hi, a->n_pad, K, K, a->n_pad)) || (rc = tc_make_map(&mw_lo, a->w_lo, a->n_pad, K, K, a->n_pad))) return rc;
    CUtensorMap ma[3], mmin;
    for (int s = 0; s < 3; ++s) {   // unused slots repeat source 0 (a tensor map must be valid even if never dereferenced)
        const int q = a->in[s] ? s : 0;
        if ((rc = tc_make_map(&ma[s], a->in[q], M, a->in_channels[q], a->in_channels[q], kTcM))) return rc;
    }
    if ((rc = tc_make_map(&mmin, a->in_min ? a->in_min : a->in[0], M, a->in_channels[0], a->in_channels[0], kTcM))) return rc;
    const size_t a_stage = (size_t)2 * kTcABytes;
    const size_t w_all = (size_t)(K / kTcKB) * 2 * a->n_pad * kTcKB * 4;          // hi + lo of the whole weight matrix
    const size_t fixed = (size_t)(4 * K + 2 * a->n_pad + (a->epilogue == TC_EPI_PLAIN ? 8 * 32 * 36 : (a->epilogue == TC_EPI_FLOW ? 0 : 8 * 32 * 20)) + 4 * 128 * 2) * sizeof(float) + 1024 + 64;
    const size_t budget = (size_t)kSmemBudget - 2048 /* static barriers */ - fixed;
    // Weights stay resident in shared memory when that still leaves a ring of >= 3 activation stages (re-streaming the
    // same few KB per tile from every SM hot-spots a handful of L2 slices); otherwise they travel with the k-blocks.
    const size_t w_kb = (size_t)2 * a->n_pad * kTcKB * 4;
    const int stages_res = w_all < budget ? (int)((budget - w_all) / a_stage) : 0;
    const int stages_str = (int)(budget / (a_stage + w_kb));
    p.w_resident = (stages_res >= 3 || stages_res >= stages_str) ? 1 : 0;
    if (const char* e = getenv("PVRAFT_TC_WRES")) p.w_resident = (atoi(e) != 0 && stages_res >= 2) ? 1 : 0;   // (debug override)
    int stages = p.w_resident ? stages_res : stages_str;
    stages = stages < 1 ? 1 : (stages > kTcMaxStages ? kTcMaxStages : stages);
    const size_t stage = a_stage + (p.w_resident ? 0 : w_kb);
    if (const char* e = getenv("PVRAFT_TC_STAGES")) { const int v = atoi(e); if (v >= 1 && v <= stages) stages = v; }
    if (stages < 2) return fail(PVRAFT_ERR_SMEM, "tc_linear: K=%d, n_pad=%d leave room for only %d operand stage(s) (2 needed)", K, a->n_pad, stages);
    p.stages = stages;
    if (const char* e = getenv("PVRAFT_TC_DBG")) p.dbg = atoi(e);
    const size_t smem = stages * stage + (p.w_resident ? w_all : 0) + fixed;
    if ((rc = opt_in_smem(k_tc_linear, smem))) return rc;
    const long long n_tiles = (M + kTcM - 1) / kTcM;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    // launched with programmatic stream serialization: the kernel's prologue may start while the previous kernel drains
    static const bool pdl = []() { const char* e = getenv("PVRAFT_TC_PDL"); return !(e && atoi(e) == 0); }();
    // chaining replaces the grid-wide wait, so it needs the early start PDL gives and parameters that are already in place
    p.wait_on = (a->wait_on && pdl && p.settled) ? a->wait_on : nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, k_tc_linear, mw_hi, mw_lo, ma[0], ma[1], ma[2], mmin, p);
    if (le != cudaSuccess) return fail((int)le, "tc_linear: launch failed: %s", cudaGetErrorString(le));
    return check_launch("tc_linear");
}
