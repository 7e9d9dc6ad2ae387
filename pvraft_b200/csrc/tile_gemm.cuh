// Register-tiled fp32 point-MLP building block shared by the dense kernels.
//
// A CTA of 256 threads owns a tile of TP = 64 points.  Activations sit in shared memory
// point-major ([point][k], row stride AS floats, AS % 4 == 0 and (AS/4) odd so that the 16-byte
// words of consecutive rows land in different banks), weights k-major ([k][channel], stride WS).
// Thread (tx = tid & 15, ty = tid >> 4) accumulates a 4-point x (4*CM4)-channel register tile:
// points ty*4 .. ty*4+3, channels tx*4 + 64*g .. +3 for g < CM4.  The k loop runs in chunks of 4
// with 128-bit shared loads for both operands (activation words are warp-broadcast: a warp spans
// only two ty values).
#pragma once
#include "common.cuh"

namespace pvraft {

constexpr int kTP = 64;          // points per tile
constexpr int kMlpThreads = 256; // threads per CTA of the dense kernels

template <int CM4>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ act, int AS, const float* __restrict__ w, int WS,
                                          int KD, float (&acc)[4][CM4 * 4]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* a0 = act + (ty * 4) * AS;
    const float* w0 = w + tx * 4;
    // software pipeline: the shared-memory fragments of chunk k+4 are fetched before the FMAs of chunk k
    // (a CTA of 8 warps leaves only 2 warps per scheduler, too few to hide the LDS latency otherwise)
    float4 a[4], wv[4][CM4];
#pragma unroll
    for (int p = 0; p < 4; ++p) a[p] = *reinterpret_cast<const float4*>(a0 + p * AS);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int g = 0; g < CM4; ++g) wv[kk][g] = *reinterpret_cast<const float4*>(w0 + kk * WS + g * 64);
#pragma unroll 1
    for (int k = 0; k < KD; k += 4) {
        float4 an[4], wn[4][CM4];
        const int kn = k + 4 < KD ? k + 4 : k;   // the last iteration re-reads its own chunk (harmless)
#pragma unroll
        for (int p = 0; p < 4; ++p) an[p] = *reinterpret_cast<const float4*>(a0 + p * AS + kn);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int g = 0; g < CM4; ++g) wn[kk][g] = *reinterpret_cast<const float4*>(w0 + (kn + kk) * WS + g * 64);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int g = 0; g < CM4; ++g) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float av = kk == 0 ? a[p].x : kk == 1 ? a[p].y : kk == 2 ? a[p].z : a[p].w;
                    acc[p][g * 4 + 0] = fmaf(av, wv[kk][g].x, acc[p][g * 4 + 0]);
                    acc[p][g * 4 + 1] = fmaf(av, wv[kk][g].y, acc[p][g * 4 + 1]);
                    acc[p][g * 4 + 2] = fmaf(av, wv[kk][g].z, acc[p][g * 4 + 2]);
                    acc[p][g * 4 + 3] = fmaf(av, wv[kk][g].w, acc[p][g * 4 + 3]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) a[p] = an[p];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int g = 0; g < CM4; ++g) wv[kk][g] = wn[kk][g];
    }
}

// Stage a [cout][cin_total] row-major global weight (columns col0 .. col0+cin-1) as k-major [KD][WS] in shared
// memory.  Global reads are coalesced (source order); the transposed shared-memory stores hit 8 banks because
// WS = padded_cout + 4 (wstride()), i.e. a 4-way conflict instead of 32-way.  Rows cin..KD-1 and columns
// cout..CP-1 of the [KD][CP] window are zero filled.
__device__ __forceinline__ void stage_weight(float* __restrict__ dst, int KD, int WS, int CP, const float* __restrict__ src,
                                             int cout, int cin_total, int col0, int cin) {
    for (int i = threadIdx.x; i < cout * cin; i += blockDim.x) {
        const int c = i / cin, k = i - c * cin;
        dst[k * WS + c] = __ldg(src + (size_t)c * cin_total + col0 + k);
    }
    if (KD > cin || CP > cout) {
        for (int i = threadIdx.x; i < KD * CP; i += blockDim.x) {
            const int k = i / CP, c = i - k * CP;
            if (k >= cin || c >= cout) dst[k * WS + c] = 0.f;
        }
    }
}

__device__ __forceinline__ void stage_vector(float* __restrict__ dst, int n_pad, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) dst[i] = (src != nullptr && i < n) ? __ldg(src + i) : 0.f;
}

// padded activation row stride: multiple of 4 floats with an odd number of 16-byte words
__host__ __device__ __forceinline__ int act_stride(int k_pad) {
    int s = k_pad + 4;
    if (((s / 4) & 1) == 0) s += 4;
    return s;
}

// row stride of a k-major weight tile with `cp` (multiple of 32) padded output channels
__host__ __device__ __forceinline__ int wstride(int cp) { return cp + 4; }
__host__ __device__ __forceinline__ int pad4(int x) { return (x + 3) & ~3; }
__host__ __device__ __forceinline__ int pad64(int x) { return (x + 63) & ~63; }

}  // namespace pvraft
