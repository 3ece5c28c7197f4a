// wgrad_probe.cu -- standalone development harness (tool) for the weight gradient of a 3x3 convolution on tcgen05:
//     dW[n, c, dy, dx] = sum over (b, y, x) of  g[b, n, y, x] * X[b, c, y + dy - 1, x + dx - 1]        (X zero-padded)
// as a split-K GEMM over PIXELS: every CTA walks over 32-pixel chunks (b, y', x'0), D[(tap, n), c] += G_tap[(n), px] * X[c, px],
// M = 128 = 4 taps x 32 outputs (three MMAs cover the nine taps), N = 64 channels, K = 32 pixels = 4 x (K = 8), both operands
// K-major (pixels are contiguous in NCHW) in the 128-byte swizzle, written by TMA.  The +-1 pixel tap shift along K cannot
// be a descriptor or TMA start (4 bytes), so the caller passes three copies of g shifted by -1 / 0 / +1 pixels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o wgrad_probe tools/wgrad_probe.cu -lcuda && ./wgrad_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

constexpr int C = 64, NP = 32;                      // channels of X, padded outputs per tap
constexpr int XT = C * 128, GT = NP * 128;          // bytes of an X tile [64][32 px], of one tap tile [32][32 px]
constexpr int STAGE = XT + 12 * GT;                 // X + 12 tap slots (9 used)
constexpr int RING = 3;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, 128-byte swizzle: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO); LBO unused (1); version 1; layout type 2
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)1 << 16 | (uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61;
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (long spin = 0; spin < (1L << 26) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done) __trap();
}

// grid = min(items, SMs), block = 192: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (once, at the end)
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, int B, int H, int W, int n_out,
             float *__restrict__ dW)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[RING], empty[RING], done_bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = (W + 31) / 32;
    const long items = (long)B * H * chunks;
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        uint32_t g = 0;
        for (long it = blockIdx.x; it < items; it += gridDim.x, ++g) {
            const int xc = (int)(it % chunks), y = (int)((it / chunks) % H), b = (int)(it / ((long)chunks * H));
            const uint32_t slot = g % RING;
            mbar_wait(&empty[slot], ((g / RING) & 1u) ^ 1u);
            const uint32_t sx = ring + slot * STAGE, sg = sx + XT;
            if (elect_one()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(XT + 9 * GT) : "memory");
                // X tile: dims (x, channel, row, image), box {32, 64, 1, 1} -> [64 ch][32 px]
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sx), "l"(reinterpret_cast<uint64_t>(&map_x)), "r"(smem_u32(&full[slot])), "r"(xc * 32), "r"(0), "r"(y), "r"(b) : "memory");
                // g tiles: copy dxi (0: shifted so that gs[x] = g[x + 1], 1: g, 2: gs[x] = g[x - 1]), rows y - 1 .. y + 1:
                // box {32, 32 ch, 3 rows, 1} -> slots dxi * 3 + r, r = 0..2 <-> dy = 2 - r
                for (int dxi = 0; dxi < 3; ++dxi)
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(sg + dxi * 3 * GT), "l"(reinterpret_cast<uint64_t>(&map_g)), "r"(smem_u32(&full[slot])), "r"(xc * 32), "r"(0),
                                 "r"(y - 1), "r"(dxi * B + b) : "memory");
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // D fp32, A = B = TF32, both K-major, N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t g = 0;
        for (long it = blockIdx.x; it < items; it += gridDim.x, ++g) {
            const uint32_t slot = g % RING;
            mbar_wait(&full[slot], (g / RING) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sx = ring + slot * STAGE, sg = sx + XT;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t db = desc_k_sw128(sx + kk * 32);
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const uint64_t da = desc_k_sw128(sg + m * 4 * GT + kk * 32);
                        const uint32_t acc = (g | (uint32_t)kk) != 0 ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                     ::"r"(tmem + m * 64), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
        __syncwarp();
    } else {
        // epilogue: accumulator m, lane (4 tap slots x 32 outputs), 64 columns = channels
        mbar_wait(&done_bar, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;                       // TMEM lane quarter of this warp
        const int row = quarter * 32 + lane, tl = row >> 5, n = row & 31;
        for (int m = 0; m < 3; ++m) {
            const int s = m * 4 + tl;                       // tap slot = dxi * 3 + r
            uint32_t v[64];
            for (int cb = 0; cb < 64; cb += 16) {
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * 64 + cb);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[cb + 0]), "=r"(v[cb + 1]), "=r"(v[cb + 2]), "=r"(v[cb + 3]), "=r"(v[cb + 4]), "=r"(v[cb + 5]),
                               "=r"(v[cb + 6]), "=r"(v[cb + 7]), "=r"(v[cb + 8]), "=r"(v[cb + 9]), "=r"(v[cb + 10]), "=r"(v[cb + 11]),
                               "=r"(v[cb + 12]), "=r"(v[cb + 13]), "=r"(v[cb + 14]), "=r"(v[cb + 15])
                             : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (s < 9 && n < n_out && items > 0) {
                const int dxi = s / 3, r = s % 3, dy = 2 - r, dx = dxi;
                for (int c = 0; c < 64; ++c) atomicAdd(dW + ((long)(n * C + c) * 3 + dy) * 3 + dx, __uint_as_float(v[c]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

int main(int argc, char **argv)
{
    const int B = argc > 1 ? atoi(argv[1]) : 2, H = argc > 2 ? atoi(argv[2]) : 7, W = argc > 3 ? atoi(argv[3]) : 44, n = argc > 4 ? atoi(argv[4]) : 5;
    const bool check = (long)B * H * W <= 200000;
    const long P = (long)H * W;
    std::vector<float> X((size_t)B * C * P), G((size_t)B * n * P), GS((size_t)3 * B * n * P, 0.f), dW((size_t)n * C * 9), R((size_t)n * C * 9);
    srand(5);
    for (auto &v : X) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto &v : G) v = (rand() % 2001 - 1000) / 1000.0f;
    // gs[dxi][.., x] = g[.., x - dxi + 1] (zero outside the row)
    for (int d = 0; d < 3; ++d)
        for (long r = 0; r < (long)B * n * H; ++r)
            for (int x = 0; x < W; ++x) {
                const int xs = x - d + 1;
                GS[((size_t)d * B * n * H + r) * W + x] = (xs >= 0 && xs < W) ? G[r * W + xs] : 0.f;
            }
    if (check)
        for (int o = 0; o < n; ++o)
            for (int c = 0; c < C; ++c)
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        double s = 0;
                        for (int b = 0; b < B; ++b)
                            for (int y = 0; y < H; ++y)
                                for (int x = 0; x < W; ++x) {
                                    const int yy = y + dy - 1, xx = x + dx - 1;
                                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                                    s += (double)tf32(G[((size_t)(b * n + o) * H + y) * W + x]) * tf32(X[((size_t)(b * C + c) * H + yy) * W + xx]);
                                }
                        R[((size_t)(o * C + c) * 3 + dy) * 3 + dx] = (float)s;
                    }
    float *dX, *dG, *dD;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dG, GS.size() * 4); cudaMalloc(&dD, dW.size() * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dG, GS.data(), GS.size() * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto make = [&](CUtensorMap *m, float *base, int chans, int imgs, int box_c, int box_r) {
        const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)chans, (cuuint64_t)H, (cuuint64_t)imgs};
        const cuuint64_t strides[3] = {(cuuint64_t)P * 4, (cuuint64_t)W * 4, (cuuint64_t)chans * P * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)box_c, (cuuint32_t)box_r, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        return ((EncodeTiledFn)fp)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap mx, mg;
    if (make(&mx, dX, C, B, 64, 1) != CUDA_SUCCESS || make(&mg, dG, n, 3 * B, 32, 3) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const size_t smem = (size_t)RING * STAGE + 1024;
    cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long items = (long)B * H * ((W + 31) / 32);
    const int grid = (int)(items < sms ? items : sms);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(dD, 0, dW.size() * 4);
        cudaEventRecord(e0);
        wgrad_kernel<<<grid, 192, smem>>>(mx, mg, B, H, W, n, dD);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cuda: %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = fminf(best, ms);
    }
    cudaMemcpy(dW.data(), dD, dW.size() * 4, cudaMemcpyDeviceToHost);
    double maxd = 0, maxr = 0;
    if (check)
        for (size_t i = 0; i < dW.size(); ++i) { maxd = fmax(maxd, fabs(dW[i] - R[i])); maxr = fmax(maxr, fabs(R[i])); }
    printf("B %d H %d W %d n %d: %.3f ms  max|dW - ref| = %.4g (max|ref| %.4g)  dW[0..2] = %g %g %g  ref %g %g %g\n", B, H, W, n, best, maxd, maxr,
           dW[0], dW[1], dW[2], R[0], R[1], R[2]);
    return (!check || maxd < 2e-2 * fmax(maxr, 1.0)) ? 0 : 1;
}
