// red_bench.cu -- microbenchmark (not product code): cost of the backward scatter on B200.
// Each thread = one pixel, 8 taps; every tap adds 4 bilinear corner weights into a plane.
//   mode 0: 4 scalar REDG.F32 per tap into a row-major plane (what iter_bwd_kernel v1 does)
//   mode 1: 2 REDG.F32x2 per tap into two row-major planes shifted by one float
//   mode 2: 1 REDG.F32x4 per tap into four 2x2-blocked planes with different block phases
//   mode 3: plain stores (no atomics; wrong sums) -- LSU/L2 store ceiling for the same pattern
//   mode 4: no memory writes at all (address/weight math + offset loads only)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench tools/red_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) scatter(const float* __restrict__ off, const float* __restrict__ gy,
                                               float* __restrict__ S, int H, int W, long plane_stride)
{
    const int P = H * W;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const float g = gy[b * P + r];
    const float* ob = off + b * 16 * P + r;
    float* sb = S + b * plane_stride * 4;
    const int Wb = W / 2 + 2;
    float sink = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int tt = t < 4 ? t : t + 1;
        const float h_im = (float)(h - 1 + tt / 3) + ob[(long)(2 * t) * P];
        const float w_im = (float)(w - 1 + tt % 3) + ob[(long)(2 * t + 1) * P];
        if (!(h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W)) continue;
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int hl = (int)hf, wl = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf;
        const float w1 = (1 - lh) * (1 - lw) * g, w2 = (1 - lh) * lw * g, w3 = lh * (1 - lw) * g, w4 = lh * lw * g;
        if (MODE == 0) {
            float* p = sb + (long)(hl + 1) * (W + 2) + (wl + 1);   // padded plane: no guards needed
            atomicAdd(p, w1); atomicAdd(p + 1, w2); atomicAdd(p + W + 2, w3); atomicAdd(p + W + 3, w4);
        } else if (MODE == 1) {
            const int sx = (wl + 1) & 1;                           // plane sx is shifted by sx floats
            float* p = sb + sx * plane_stride + (long)(hl + 1) * (W + 4) + (wl + 1) + sx;
            atomicAdd((float2*)p, make_float2(w1, w2));
            atomicAdd((float2*)(p + W + 4), make_float2(w3, w4));
        } else if (MODE == 2) {
            const int sy = (hl + 1) & 1, sx = (wl + 1) & 1;
            const int by = (hl + 1 + sy) >> 1, bx = (wl + 1 + sx) >> 1;
            float* p = sb + (sy * 2 + sx) * plane_stride + ((long)by * Wb + bx) * 4;
            atomicAdd((float4*)p, make_float4(w1, w2, w3, w4));
        } else if (MODE == 3) {
            const int sy = (hl + 1) & 1, sx = (wl + 1) & 1;
            const int by = (hl + 1 + sy) >> 1, bx = (wl + 1 + sx) >> 1;
            float* p = sb + (sy * 2 + sx) * plane_stride + ((long)by * Wb + bx) * 4;
            *(float4*)p = make_float4(w1, w2, w3, w4);
        } else {
            sink += w1 + w2 + w3 + w4;
        }
    }
    if (MODE == 4 && sink == 123.456f) S[0] = sink;
}

int main(int argc, char** argv)
{
    const int B = 8, H = 352, W = 1216, P = H * W;
    const float sigma = argc > 1 ? atof(argv[1]) : 2.0f;
    const int smooth = argc > 2 ? atoi(argv[2]) : 0;
    std::vector<float> off((size_t)B * 16 * P), gy((size_t)B * P);
    std::mt19937 rng(7240);
    std::normal_distribution<float> nd(0.f, sigma);
    if (!smooth) for (auto& v : off) v = nd(rng);
    else {   // spatially coherent offsets: one random value per 16x16 cell + small jitter
        for (int b = 0; b < B; ++b) for (int c = 0; c < 16; ++c) {
            std::vector<float> cell((H / 16 + 1) * (W / 16 + 1));
            for (auto& v : cell) v = nd(rng);
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
                off[((size_t)b * 16 + c) * P + y * W + x] = cell[(y / 16) * (W / 16 + 1) + x / 16] + 0.05f * nd(rng);
        }
    }
    for (auto& v : gy) v = nd(rng);
    float *d_off, *d_gy, *d_S;
    const long plane_stride = (long)(H + 4) * (W + 8);
    CK(cudaMalloc(&d_off, off.size() * 4)); CK(cudaMalloc(&d_gy, gy.size() * 4));
    CK(cudaMalloc(&d_S, (size_t)B * 4 * plane_stride * 4));
    CK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gy, gy.data(), gy.size() * 4, cudaMemcpyHostToDevice));
    dim3 grid((P + 255) / 256, B);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[5] = {"4x RED.F32", "2x RED.F32x2 (shifted planes)", "1x RED.F32x4 (blocked planes)", "1x ST.128 (no atomics)", "math only"};
    float* d_flush; const size_t flush_bytes = 512u << 20;
    CK(cudaMalloc(&d_flush, flush_bytes));
    const int cold = argc > 3 ? atoi(argv[3]) : 0;   // 1: evict the scatter planes from L2 before each run
    for (int mode = 0; mode < 5; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaMemsetAsync(d_S, 0, (size_t)B * 4 * plane_stride * 4));
            if (cold) CK(cudaMemsetAsync(d_flush, 1, flush_bytes));
            cudaEventRecord(e0);
            switch (mode) {
            case 0: scatter<0><<<grid, 256>>>(d_off, d_gy, d_S, H, W, plane_stride); break;
            case 1: scatter<1><<<grid, 256>>>(d_off, d_gy, d_S, H, W, plane_stride); break;
            case 2: scatter<2><<<grid, 256>>>(d_off, d_gy, d_S, H, W, plane_stride); break;
            case 3: scatter<3><<<grid, 256>>>(d_off, d_gy, d_S, H, W, plane_stride); break;
            default: scatter<4><<<grid, 256>>>(d_off, d_gy, d_S, H, W, plane_stride); break;
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("cold=%d sigma=%.1f smooth=%d  %-34s %8.3f ms  %7.2f Gpix/s  %7.1f G corner-adds/s\n", cold, sigma, smooth,
               names[mode], best, (double)B * P / best / 1e6, (double)B * P * 32 / best / 1e6);
    }
    return 0;
}
