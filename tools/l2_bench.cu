// l2_bench.cu -- microbenchmark (not product code): L2-resident read and write bandwidth of the B200,
// the second roofline level SURVEY 8(d) asks for (MEASURED_PEAKS.json carries only the HBM figure).
// A buffer of S MB (S << 126 MB L2) is read / written REPS times by one launch of 148*8 CTAs with
// 128-bit accesses; the first pass warms L2, the timed launch then runs out of L2 alone.  Also prints
// the same kernels on a 4 GB buffer (HBM) as a cross-check of MEASURED_PEAKS.json.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bench tools/l2_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(512) rd(const float4* __restrict__ p, size_t n, int reps, float* sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n; i += 4 * stride) {
            float4 a = __ldcg(p + i), b = __ldcg(p + i + stride), c = __ldcg(p + i + 2 * stride), d = __ldcg(p + i + 3 * stride);
            acc += a.x + b.y + c.z + d.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}
__global__ void __launch_bounds__(512) wr(float4* __restrict__ p, size_t n, int reps)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        const float4 v = make_float4((float)r, 1.f, 2.f, 3.f);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcg(p + i, v);
    }
}
__global__ void __launch_bounds__(512) cp(const float4* __restrict__ a, float4* __restrict__ b, size_t n, int reps)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcg(b + i, __ldcg(a + i));
}

template <class F> static float timed(F f)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4;
    float* sink; CK(cudaMalloc(&sink, 4));
    printf("{\"sms\": %d", sms);
    const size_t sizes_mb[] = {16, 32, 64, 4096};
    for (size_t s : sizes_mb) {
        const size_t bytes = s << 20, n = bytes / 16;
        float4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
        const int reps = s >= 1024 ? 1 : 64;
        float t_r = timed([&] { rd<<<grid, 512>>>(a, n, reps, sink); });
        float t_w = timed([&] { wr<<<grid, 512>>>(a, n, reps); });
        float t_c = timed([&] { cp<<<grid, 512>>>(a, b, n / 2 * 1, reps); });
        printf(", \"%s_%zuMB\": {\"read_gbs\": %.1f, \"write_gbs\": %.1f, \"copy_gbs_rw\": %.1f}", s >= 1024 ? "hbm" : "l2", s,
               bytes * (double)reps / t_r / 1e6, bytes * (double)reps / t_w / 1e6, (double)(n / 2) * 16 * 2 * reps / t_c / 1e6);
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    printf("}\n");
    return 0;
}
