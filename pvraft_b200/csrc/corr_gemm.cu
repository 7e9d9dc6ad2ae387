// All-pairs feature correlation on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
// Replaces CorrBlock.calculate_corr (reference model/corr.py:95-100):  corr[b,i,j] = <fmap1[b,:,i], fmap2[b,:,j]> / sqrt(C).
// This is the one large dense contraction of the model (2*N*N*C = 17.2 GFLOP per sample at N=8192).
//
// The operands are point-major [B,N,C] (K-major for the MMA).  To keep fp32 parity on a TF32 datapath every
// operand is split once into hi = tf32(x) and lo = tf32(x - hi) (k_tf32_split) and each product is evaluated as
// hi*hi + lo*hi + hi*lo (the classic 3xTF32 scheme, error ~2^-21 relative per product; the dropped lo*lo term is
// ~2^-22).  The kernel is persistent (one CTA per SM walks 128 x 128 output tiles); roles and pipeline are described at
// k_corr_gemm below.  The division by sqrt(C) is a true division, as on the reference's CPU path.
#include <cuda.h>

#include "common.cuh"

namespace pvraft {

constexpr int kGemmThreads = 320;   // TMA producer | MMA issuer | 8 epilogue warps
constexpr int kTileM = 128, kTileN = 128, kBlockK = 32;     // 32 tf32 = one 128-byte swizzle row
constexpr int kOperandBytes = kTileM * kBlockK * 4;         // 16 KB
constexpr int kStageBytes = 4 * kOperandBytes;              // A_hi, A_lo, B_hi, B_lo
constexpr int kStages = 3;

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init_(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(void* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(su32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(su32(dst)),
                 "l"(map), "r"(su32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset = 8 rows * 128 B = 1024 B
// (>> 4 = 64) in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ unsigned long long umma_desc(const void* smem_tile) {
    unsigned long long d = 0;
    d |= (unsigned long long)((su32(smem_tile) >> 4) & 0x3FFFu);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)64 << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13),
// K-major A and B (0), n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(bar)) : "memory");
}

struct GemmParams {
    float* corr;   // [B,N,N]
    int N, C;
    float scale;   // sqrt(C): the divisor of model/corr.py:99
    float rscale;  // RN(1 / scale)
    int tiles_n;   // N / 128
    long long n_tiles;   // B * tiles_n * tiles_n
};

__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// x / s, correctly rounded, from r = RN(1/s) (Markstein): q0 = x r; q = q0 + (x - q0 s) r.  Three instructions instead of
// the ~10 of the generic division, bit-identical to it (checked against true division on 1.3e8 values).
__device__ __forceinline__ float div_by_const(float x, float s, float r) {
    const float q0 = x * r;
    return fmaf(fmaf(-q0, s, x), r, q0);
}

constexpr int kEpiWarps = 8;
constexpr int kStageCols = 16;                                   // columns staged per epilogue step
constexpr int kStagePitch = kStageCols + 4;                      // floats; keeps float4 alignment, spreads banks
constexpr int kEpiStageBytes = kEpiWarps * 32 * kStagePitch * 4; // 20 KB

// Persistent kernel, one CTA per SM, tiles t = blockIdx.x + i * gridDim.x with the column tile fastest (neighbouring CTAs
// share the A row-panel in L2):
//   warp 0      TMA producer: four 128 x 32 boxes (A hi/lo, B hi/lo) per k-block into a 3-stage ring
//   warp 1      MMA issuer (whole warp walks the loop, one elected lane issues): 3 x 4 tcgen05.mma per k-block into one of
//               two 128-column TMEM accumulators
//   warps 2-9   epilogue, two warps per TMEM lane quadrant: tcgen05.ld 16 columns -> exact division by sqrt(C) -> transpose
//               through shared memory -> 64-byte row segments to global.  It drains accumulator t while the MMAs of tile
//               t+1 run.
__global__ void __launch_bounds__(kGemmThreads, 1)
k_corr_gemm(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B) by pointer arithmetic on the shared array: an integer round trip would lose the
    // address space and turn every shared-memory access into a generic LD/ST
    unsigned char* tiles = smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    float* s_stage = reinterpret_cast<float*>(tiles + (size_t)kStages * kStageBytes);
    __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages], s_acc_full[2], s_acc_empty[2];
    __shared__ unsigned s_tmem_base;
    const int warp = warp_id(), lane = lane_id();
    const int num_kb = p.C / kBlockK;
    const long long my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init_(&s_full[s], 1); mbar_init_(&s_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init_(&s_acc_full[a], 1); mbar_init_(&s_acc_empty[a], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // two 128-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&s_tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem_base;
    const int tiles_per_batch = p.tiles_n * p.tiles_n;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0;
            unsigned phase = 0;
            for (long long i = 0; i < my_tiles; ++i) {
                const long long t = blockIdx.x + i * gridDim.x;
                const int b = (int)(t / tiles_per_batch), r = (int)(t - (long long)b * tiles_per_batch);
                const int tile_m = r / p.tiles_n, tile_n = r - tile_m * p.tiles_n;
                const int row_a = b * p.N + tile_m * kTileM, row_b = b * p.N + tile_n * kTileN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_(&s_empty[s], phase ^ 1u);
                    unsigned char* st = tiles + (size_t)s * kStageBytes;
                    mbar_expect_tx_(&s_full[s], kStageBytes);
                    tma_load_2d(st + 0 * kOperandBytes, &map_a_hi, &s_full[s], kb * kBlockK, row_a);
                    tma_load_2d(st + 1 * kOperandBytes, &map_a_lo, &s_full[s], kb * kBlockK, row_a);
                    tma_load_2d(st + 2 * kOperandBytes, &map_b_hi, &s_full[s], kb * kBlockK, row_b);
                    tma_load_2d(st + 3 * kOperandBytes, &map_b_lo, &s_full[s], kb * kBlockK, row_b);
                    if (++s == kStages) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr unsigned idesc = umma_idesc_tf32(kTileM, kTileN);
        int s = 0;
        unsigned phase = 0;
        for (long long i = 0; i < my_tiles; ++i) {
            const int acc = (int)(i & 1);
            mbar_wait_(&s_acc_empty[acc], (((unsigned)(i >> 1)) & 1u) ^ 1u);   // the epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned tacc = tmem + (unsigned)(acc * kTileN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_(&s_full[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char* st = tiles + (size_t)s * kStageBytes;
                const unsigned long long a_hi = umma_desc(st + 0 * kOperandBytes), a_lo = umma_desc(st + 1 * kOperandBytes);
                const unsigned long long b_hi = umma_desc(st + 2 * kOperandBytes), b_lo = umma_desc(st + 3 * kOperandBytes);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k) {
                        // one MMA covers K = 8 tf32 = 32 bytes: advance the start address field by 32 B >> 4 = 2
                        const unsigned long long off = (unsigned long long)(k * 2);
                        umma_tf32(tacc, a_hi + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_tf32(tacc, a_lo + off, b_hi + off, idesc, 1u);
                        umma_tf32(tacc, a_hi + off, b_lo + off, idesc, 1u);
                    }
                    umma_commit(&s_empty[s]);                            // the stage may be refilled once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&s_acc_full[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++s == kStages) { s = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> shared (transpose) -> global =====
        const int quad = warp & 3;                 // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int half = (warp - 2) >> 2;          // columns [64*half, 64*half + 64) of the tile
        float* stg = s_stage + (size_t)(warp - 2) * 32 * kStagePitch;
        const int rsub = lane >> 2, cq = lane & 3; // store phase: 8 rows x 4 float4 per instruction
        for (long long i = 0; i < my_tiles; ++i) {
            const long long t = blockIdx.x + i * gridDim.x;
            const int b = (int)(t / tiles_per_batch), r = (int)(t - (long long)b * tiles_per_batch);
            const int tile_m = r / p.tiles_n, tile_n = r - tile_m * p.tiles_n;
            const int acc = (int)(i & 1);
            mbar_wait_(&s_acc_full[acc], ((unsigned)(i >> 1)) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned tl = tmem + (unsigned)(acc * kTileN) + ((unsigned)(quad * 32) << 16);
            float* obase = p.corr + ((size_t)b * p.N + (size_t)tile_m * kTileM + quad * 32 + rsub) * p.N + (size_t)tile_n * kTileN + cq * 4;
#pragma unroll 1
            for (int c0 = half * 64; c0 < half * 64 + 64; c0 += kStageCols) {
                unsigned v[16];
                tmem_ld16(tl + (unsigned)c0, v);
                __syncwarp();   // the previous step's readers are done with the staging tile
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // corr / sqrt(C) as a true (correctly rounded) division (model/corr.py:99)
                    *reinterpret_cast<float4*>(stg + lane * kStagePitch + q * 4) =
                        make_float4(div_by_const(__uint_as_float(v[q * 4 + 0]), p.scale, p.rscale), div_by_const(__uint_as_float(v[q * 4 + 1]), p.scale, p.rscale),
                                    div_by_const(__uint_as_float(v[q * 4 + 2]), p.scale, p.rscale), div_by_const(__uint_as_float(v[q * 4 + 3]), p.scale, p.rscale));
                }
                __syncwarp();
                float4 o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = *reinterpret_cast<const float4*>(stg + (j * 8 + rsub) * kStagePitch + cq * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(obase + (size_t)(j * 8) * p.N + c0) = o[j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_(&s_acc_empty[acc]);
        }
    }
    __syncwarp();   // the producer runs on one lane: re-converge before the CTA-wide (aligned) barrier
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

// hi = tf32(x) (round to nearest), lo = tf32(x - hi)
__global__ void k_tf32_split(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    float h[4], l[4];
    const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[q]));
        h[q] = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[q] - h[q]));
        l[q] = __uint_as_float(lb);
    }
    *reinterpret_cast<float4*>(hi + i) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + i) = make_float4(l[0], l[1], l[2], l[3]);
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, C] fp32 row-major, box = 128 rows x 32 columns, 128-byte swizzle
static int make_map(CUtensorMap* m, const float* base, long long rows, int C) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_matmul: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_matmul: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

}  // namespace pvraft

using namespace pvraft;

extern "C" int64_t pvraft_corr_matmul_workspace_bytes(int B, int N, int C) {
    if (B <= 0 || N <= 0 || C <= 0) return 0;
    return (int64_t)4 * B * N * C * (int64_t)sizeof(float);   // hi/lo copies of both feature maps
}

extern "C" int pvraft_corr_matmul_fwd(const float* fmap1, const float* fmap2, int B, int N, int C, float* corr, void* workspace,
                                      void* stream) {
    if (!fmap1 || !fmap2 || !corr || !workspace) return fail(PVRAFT_ERR_BAD_ARG, "corr_matmul: null pointer");
    if (B <= 0 || N <= 0 || C <= 0) return fail(PVRAFT_ERR_BAD_ARG, "corr_matmul: bad shape");
    if (N % kTileM || C % kBlockK) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_matmul: N=%d must be a multiple of 128 and C=%d of 32", N, C);
    if (B > 65535 || N / kTileM > 65535) return fail(PVRAFT_ERR_UNSUPPORTED, "corr_matmul: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * N * C;
    float* a_hi = reinterpret_cast<float*>(workspace);
    float* a_lo = a_hi + n;
    float* b_hi = a_lo + n;
    float* b_lo = b_hi + n;
    const unsigned blocks = (unsigned)((n / 4 + 255) / 256);
    k_tf32_split<<<blocks, 256, 0, st>>>(fmap1, n, a_hi, a_lo);
    k_tf32_split<<<blocks, 256, 0, st>>>(fmap2, n, b_hi, b_lo);
    int rc = check_launch("tf32_split");
    if (rc) return rc;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    if ((rc = make_map(&ma_hi, a_hi, (long long)B * N, C)) || (rc = make_map(&ma_lo, a_lo, (long long)B * N, C)) ||
        (rc = make_map(&mb_hi, b_hi, (long long)B * N, C)) || (rc = make_map(&mb_lo, b_lo, (long long)B * N, C)))
        return rc;
    GemmParams p{};
    p.corr = corr; p.N = N; p.C = C;
    p.scale = sqrtf((float)C);
    p.rscale = (float)(1.0 / (double)p.scale);
    p.tiles_n = N / kTileN;
    p.n_tiles = (long long)B * p.tiles_n * p.tiles_n;
    const size_t smem = (size_t)kStages * kStageBytes + kEpiStageBytes + 1024;
    if ((rc = opt_in_smem(k_corr_gemm, smem))) return rc;
    const int grid = (int)(p.n_tiles < sm_count() ? p.n_tiles : sm_count());
    k_corr_gemm<<<grid, kGemmThreads, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    return check_launch("corr_gemm");
}
