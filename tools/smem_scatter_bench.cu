// smem_scatter_bench.cu -- microbenchmark (not product code): the backward scatter accumulated in a
// shared-memory tile and flushed with ONE TMA tensor reduction (cp.reduce.async.bulk.tensor ... .add)
// per CTA, against the global vector-RED scatter of red_bench.cu.
//   mode 0: 4 scalar global REDs per tap into a plain plane (correctness reference)
//   mode 1: smem tile, 4x atomicAdd(float) on shared (ATOMS.CAST.SPIN loops) per tap, 1 tensor reduce
//   mode 2: two column-phase smem tiles, 2x 64-bit CAS loops per tap (a row pair per CAS), 2 tensor reduces
// Same workload as red_bench.cu: 8 KITTI frames, 8 taps, iid N(0, sigma^2) offsets.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_scatter_bench tools/smem_scatter_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void add_pair(float* p, float a, float b)
{
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *q, assumed;
    do {
        assumed = old;
        float2 o = *reinterpret_cast<float2*>(&assumed);
        o.x += a; o.y += b;
        old = atomicCAS(q, assumed, *reinterpret_cast<unsigned long long*>(&o));
    } while (old != assumed);
}

template <int MODE, int TH>
__global__ void __launch_bounds__(32 * TH) tile_scatter(const __grid_constant__ CUtensorMap map, const float* __restrict__ off,
                                                         const float* __restrict__ gy, float* __restrict__ S, int H, int W, int Wp, int Hp)
{
    // S: NBOX padded planes per image, pixel (y, x) of phase k at [(y + R) * Wp + x + R + k]: box origins are
    // never negative and always 16-byte aligned (TMA stores/reductions trap on negative or unaligned x).
    constexpr int R = 8, BW = 32 + 2 * R, BH = TH + 2 * R, NBOX = MODE == 2 ? 2 : 1;
    __shared__ __align__(128) float box[NBOX][BH * BW];
    const int P = H * W;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TH;
    const long b = blockIdx.z;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < NBOX * BH * BW; i += 32 * TH) (&box[0][0])[i] = 0.f;
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    float oh[8], ow[8];
    const float* ob = off + b * 16 * P + r;
#pragma unroll
    for (int t = 0; t < 8; ++t) { oh[t] = ob[(long)(2 * t) * P]; ow[t] = ob[(long)(2 * t + 1) * P]; }
    const float g = gy[b * P + r];
    __syncthreads();
    if (inside) {
        float* sb = S + b * NBOX * (long)Hp * Wp + (long)R * Wp + R;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int tt = t < 4 ? t : t + 1;
            const float h_im = (float)(h - 1 + tt / 3) + oh[t];
            const float w_im = (float)(w - 1 + tt % 3) + ow[t];
            if (!(h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W)) continue;
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int hl = (int)hf, wl = (int)wf;
            const float lh = h_im - hf, lw = w_im - wf;
            const float w1 = (1 - lh) * (1 - lw) * g, w2 = (1 - lh) * lw * g, w3 = lh * (1 - lw) * g, w4 = lh * lw * g;
            const int ty = hl - (y0 - R), tx = wl - (x0 - R);
            if ((unsigned)ty < (unsigned)(BH - 1) && (unsigned)tx < (unsigned)(BW - 1)) {
                if (MODE == 1) {
                    float* p = &box[0][ty * BW + tx];
                    atomicAdd(p, w1); atomicAdd(p + 1, w2); atomicAdd(p + BW, w3); atomicAdd(p + BW + 1, w4);
                } else {
                    const int ph = tx & 1;
                    float* p = &box[ph][ty * BW + tx + ph];
                    add_pair(p, w1, w2);
                    add_pair(p + BW, w3, w4);
                }
            } else {
                if (hl >= 0 && wl >= 0) atomicAdd(sb + hl * Wp + wl, w1);
                if (hl >= 0 && wl + 1 < W) atomicAdd(sb + hl * Wp + wl + 1, w2);
                if (hl + 1 < H && wl >= 0) atomicAdd(sb + (hl + 1) * Wp + wl, w3);
                if (hl + 1 < H && wl + 1 < W) atomicAdd(sb + (hl + 1) * Wp + wl + 1, w4);
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NBOX; ++k)
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                         :: "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&box[k][0])), "r"(x0), "r"(y0), "r"((int)b * NBOX + k) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) ref_scatter(const float* __restrict__ off, const float* __restrict__ gy, float* __restrict__ S, int H, int W)
{
    const int P = H * W;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const float g = gy[b * P + r];
    const float* ob = off + b * 16 * P + r;
    float* sb = S + b * P;
    for (int t = 0; t < 8; ++t) {
        const int tt = t < 4 ? t : t + 1;
        const float h_im = (float)(h - 1 + tt / 3) + ob[(long)(2 * t) * P];
        const float w_im = (float)(w - 1 + tt % 3) + ob[(long)(2 * t + 1) * P];
        if (!(h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W)) continue;
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int hl = (int)hf, wl = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf;
        const float w1 = (1 - lh) * (1 - lw) * g, w2 = (1 - lh) * lw * g, w3 = lh * (1 - lw) * g, w4 = lh * lw * g;
        if (hl >= 0 && wl >= 0) atomicAdd(sb + hl * W + wl, w1);
        if (hl >= 0 && wl + 1 < W) atomicAdd(sb + hl * W + wl + 1, w2);
        if (hl + 1 < H && wl >= 0) atomicAdd(sb + (hl + 1) * W + wl, w3);
        if (hl + 1 < H && wl + 1 < W) atomicAdd(sb + (hl + 1) * W + wl + 1, w4);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(float* base, int planes, int H, int W, int bw, int bh)
{
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

template <int MODE, int TH>
static float run(const float* d_off, const float* d_gy, float* d_S, int B, int H, int W, int Wp, int Hp)
{
    constexpr int NBOX = MODE == 2 ? 2 : 1;
    CUtensorMap map = make_map(d_S, B * NBOX, Hp, Wp, 48, TH + 16);
    dim3 grid((W + 31) / 32, (H + TH - 1) / TH, B), block(32, TH);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaMemsetAsync(d_S, 0, (size_t)B * NBOX * Hp * Wp * 4));
        cudaEventRecord(e0);
        tile_scatter<MODE, TH><<<grid, block>>>(map, d_off, d_gy, d_S, H, W, Wp, Hp);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv)
{
    const int B = 8, H = 352, W = 1216, P = H * W;
    const float sigma = argc > 1 ? atof(argv[1]) : 2.0f;
    std::vector<float> off((size_t)B * 16 * P), gy((size_t)B * P);
    std::mt19937 rng(7240);
    std::normal_distribution<float> nd(0.f, sigma);
    for (auto& v : off) v = nd(rng);
    for (auto& v : gy) v = nd(rng);
    float *d_off, *d_gy, *d_S, *d_ref;
    CK(cudaMalloc(&d_off, off.size() * 4)); CK(cudaMalloc(&d_gy, gy.size() * 4));
    const int R = 8, Wp = (W + 2 * R + 4 + 3) / 4 * 4, Hp = H + 2 * R + 1;
    CK(cudaMalloc(&d_S, (size_t)B * 2 * Hp * Wp * 4)); CK(cudaMalloc(&d_ref, (size_t)B * P * 4));
    CK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gy, gy.data(), gy.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_ref, 0, (size_t)B * P * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ref_scatter<<<dim3((P + 255) / 256, B), 256>>>(d_off, d_gy, d_ref, H, W);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms0; cudaEventElapsedTime(&ms0, e0, e1);
    std::vector<float> ref((size_t)B * P), got((size_t)B * 2 * Hp * Wp);
    CK(cudaMemcpy(ref.data(), d_ref, ref.size() * 4, cudaMemcpyDeviceToHost));
    printf("sigma=%.1f  mode 0 (4x global RED.F32, first run)          %8.3f ms\n", sigma, ms0);
    auto check = [&](const char* name, float ms, int nbox = 1) {
        CK(cudaMemcpy(got.data(), d_S, got.size() * 4, cudaMemcpyDeviceToHost));
        double md = 0, mr = 0;
        for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            double v = 0;
            for (int k = 0; k < nbox; ++k) v += got[((size_t)b * nbox + k) * Hp * Wp + (size_t)(y + R) * Wp + x + R + k];
            const double rf = ref[(size_t)b * P + y * W + x];
            md = fmax(md, fabs(v - rf)); mr = fmax(mr, fabs(rf));
        }
        printf("sigma=%.1f  %-52s %8.3f ms  %7.2f Gpix/s  max|diff| %.2e (max|ref| %.2f)\n", sigma, name, ms, (double)B * P / ms / 1e6, md, mr);
    };
    check("mode 1 smem atomicAdd.f32 x4, TH=4", run<1, 4>(d_off, d_gy, d_S, B, H, W, Wp, Hp));
    check("mode 1 smem atomicAdd.f32 x4, TH=8", run<1, 8>(d_off, d_gy, d_S, B, H, W, Wp, Hp));
    check("mode 1 smem atomicAdd.f32 x4, TH=16", run<1, 16>(d_off, d_gy, d_S, B, H, W, Wp, Hp));
    check("mode 2 smem CAS.64 x2 (column-phase tiles), TH=4", run<2, 4>(d_off, d_gy, d_S, B, H, W, Wp, Hp), 2);
    check("mode 2 smem CAS.64 x2 (column-phase tiles), TH=8", run<2, 8>(d_off, d_gy, d_S, B, H, W, Wp, Hp), 2);
    check("mode 2 smem CAS.64 x2 (column-phase tiles), TH=16", run<2, 16>(d_off, d_gy, d_S, B, H, W, Wp, Hp), 2);
    return 0;
}
