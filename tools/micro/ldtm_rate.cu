// Micro-benchmark: latency / throughput of tcgen05.ld 32x32b.x32 and of a staged epilogue chunk.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_rate ldtm_rate.cu && ./ldtm_rate
#include <cstdio>
#include <cuda_runtime.h>
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
__global__ void __launch_bounds__(128) k(int reps, int mode, float* out, long long* cyc) {
    __shared__ unsigned s_base;
    __shared__ __align__(16) float stage[4][32][36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_base)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tl = s_base + ((unsigned)(warp * 32) << 16);
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        unsigned v[32];
        tmem_ld32(tl + (unsigned)((r & 3) * 32), v);
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i] & 0x3fffffffu);
        } else {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(&stage[warp][lane][q * 4]) = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
            __syncwarp();
            const int cq = lane & 7, rsub = lane >> 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + rsub;
                const float4 x = *reinterpret_cast<const float4*>(&stage[warp][rr][cq * 4]);
                if (mode == 2) *reinterpret_cast<float4*>(out + ((size_t)blockIdx.x * 128 + warp * 32 + rr) * 64 + (r & 1) * 32 + cq * 4) = x;
                else acc += x.x + x.y + x.z + x.w;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = (t1 - t0) / reps;
    if (acc == 123.456f) out[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(128u) : "memory");
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)148 * 128 * 64 * 4); cudaMalloc(&cyc, 64);
    for (int mode = 0; mode < 3; ++mode) { k<<<148, 128>>>(8, mode, out, cyc); k<<<148, 128>>>(64, mode, out, cyc); }
    cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    printf("cycles per chunk: ldtm+sum %lld, +smem transpose %lld, +global stores %lld  (%s)\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
