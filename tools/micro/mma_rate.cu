// Micro-benchmark: sustained rate of legacy warp-level mma.sync (TF32 m16n8k8, BF16 m16n8k16) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_tf32(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = threadIdx.x + 2, a3 = threadIdx.x + 3, b0 = 5, b1 = 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = threadIdx.x + 2, a3 = threadIdx.x + 3, b0 = 5, b1 = 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
    float c[16]; for (int i = 0; i < 16; ++i) c[i] = i;
    float a = threadIdx.x * 1e-3f, b = 1.0001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], b, a);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        float ms;
        k_tf32<<<148, warps * 32>>>(out, 100); cudaEventRecord(e0); k_tf32<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 148.0 * warps * iters * 8 * (16.0 * 8 * 8 * 2) / (ms * 1e-3) / 1e12;
        k_bf16<<<148, warps * 32>>>(out, 100); cudaEventRecord(e0); k_bf16<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms2; cudaEventElapsedTime(&ms2, e0, e1);
        double bf = 148.0 * warps * iters * 8 * (16.0 * 8 * 16 * 2) / (ms2 * 1e-3) / 1e12;
        k_ffma<<<148, warps * 32>>>(out, 100); cudaEventRecord(e0); k_ffma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms3; cudaEventElapsedTime(&ms3, e0, e1);
        double ff = 148.0 * warps * 32 * iters * 16 * 2.0 / (ms3 * 1e-3) / 1e12;
        printf("warps/SM %2d: mma.sync tf32 %.1f TFLOP/s, bf16 %.1f TFLOP/s, ffma %.1f TFLOP/s\n", warps, tf, bf, ff);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
