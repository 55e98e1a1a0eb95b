// Microbenchmark: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int CH>
__global__ void k1(float* out, float a, float b, int iters) {
    float acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k2(uint64_t* out, uint64_t a, uint64_t b, int iters) {
    uint64_t acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = (uint64_t)(threadIdx.x + i) * 0x3f8000003f800000ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = ffma2(acc[i], a, b);
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* o1; uint64_t* o2;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    cudaMalloc(&o1, blocks * threads * 4); cudaMalloc(&o2, blocks * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k1<8><<<blocks, threads>>>(o1, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)blocks * threads / 32 * iters * 8;
        printf("FFMA : %.3f ms  %.1f Gwarp-inst/s  %.2f TFLOP/s\n", ms, inst / ms / 1e6, inst * 64 / ms / 1e9);
        uint64_t a = 0x3f8000413f800041ull, b = 0x3f0000003f000000ull;
        cudaEventRecord(e0); k2<8><<<blocks, threads>>>(o2, a, b, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2: %.3f ms  %.1f Gwarp-inst/s  %.2f TFLOP/s\n", ms, inst / ms / 1e6, inst * 128 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
