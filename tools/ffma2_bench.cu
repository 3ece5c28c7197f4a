// ffma2_bench.cu -- microbenchmark (not product code): issue rate of scalar FFMA vs packed FFMA2
// (fma.rn.f32x2, new on sm_100) per SM.  8 independent accumulator chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench tools/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[i]) : "f"(a), "f"(b));
        }
    } else {
        unsigned long long p[8], pa, pb;
        float2 ta = make_float2(a, a), tb = make_float2(b, b);
        pa = *reinterpret_cast<unsigned long long*>(&ta); pb = *reinterpret_cast<unsigned long long*>(&tb);
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 t = make_float2(x[2 * i], x[2 * i + 1]); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], pa, pb);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); x[2 * i] = t.x; x[2 * i + 1] = t.y; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, 1.0001f, 0.5f, iters); else k<1><<<148 * 8, 256>>>(d, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = 148.0 * 8 * 256 * 16 * (double)iters;
        printf("%s: %.3f ms, %.1f TFMA/s (%.2f fp32 TFLOP/s)\n", mode == 0 ? "FFMA  (scalar, 16 chains)" : "FFMA2 (packed, 8 chains) ", ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9);
    }
    return 0;
}
