// umma_mn_probe.cu -- probe (tool, not product): the pieces kernels_head2.cuh is built from, on the actual B200.
//   * A operand MN-major (pixels contiguous), 128-byte swizzle with 32-byte atoms, written by TMA: box {32 px, 8 channels, ROWS rows} of a
//     [8 ch][ROWS][W] fp32 tensor (dims ordered x, channel, row), four boxes side by side = 128 pixels;
//     descriptor: start = base + row * 1024, LBO = bytes between the 32-pixel atoms (one box), SBO = 512 (second group of
//     four channels), layout type SWIZZLE_128B_BASE32B; tensor map CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
//   * box start at a NEGATIVE, 16-byte aligned column (zero fill = the convolution's padding)
//   * B operand K-major un-swizzled (as in kernels_head.cuh), N = 80
//   * several accumulators side by side in one TMEM allocation (column offsets r * 80), and a narrow N = 16 MMA that
//     accumulates into columns 64..79 of accumulator 0
// Variant 0: LBO = box bytes, SBO = 512; variant 1: the two swapped.  Prints max |C - ref| per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_mn_probe tools/umma_mn_probe.cu -lcuda && ./umma_mn_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 80, NN = 16, ROWS = 3, W = 256, XS = -4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 |
           (uint64_t)1 << 46;
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return desc_kmajor(saddr, lbo, sbo) | (uint64_t)1 << 61;        // layout type 1 = SWIZZLE_128B_BASE32B (the only MN-major layout of tf32 operands)
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, const float *B, const float *B2, float *C,
                                             int *status, int variant, float *dbgA)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    float *sA = (float *)base;                              // [4 boxes][ROWS][8 ch][32 px] (swizzled by TMA)
    float *sB = sA + 4 * ROWS * 256;                        // [2][N/8][8][4]
    float *sB2 = sB + N * 8;                                // [2][NN/8][8][4]
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < N * 8; i += 128) {
        const int n = i / 8, k = i % 8;
        sB[((k / 4) * (N / 8) + n / 8) * 32 + (n % 8) * 4 + (k % 4)] = B[i];
    }
    for (int i = tid; i < NN * 8; i += 128) {
        const int n = i / 8, k = i % 8;
        sB2[((k / 4) * (NN / 8) + n / 8) * 32 + (n % 8) * 4 + (k % 4)] = B2[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    uint32_t done = 0;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(4 * ROWS * 1024) : "memory");
        for (int w = 0; w < 4; ++w)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(sA + w * ROWS * 256)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar_tma)),
                         "r"(XS + 32 * w), "r"(0), "r"(0) : "memory");
    }
    for (long spin = 0; spin < (1L << 24) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar_tma)) : "memory");
    if (!done) { if (tid == 0) *status = 2; return; }
    for (int i = tid; i < 4 * ROWS * 256; i += 128) dbgA[i] = sA[i];
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t box = ROWS * 1024;
        const uint32_t lbo = variant == 0 ? box : 512, sbo = variant == 0 ? 512 : box;
        // D = F32, A = B = TF32, A MN-major (bit 15), B K-major
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int r = 0; r < ROWS; ++r) {
            const uint64_t da = desc_mn_sw128(smem_u32(sA) + r * 1024, lbo, sbo);
            const uint64_t db = desc_kmajor(smem_u32(sB), (N / 8) * 128, 128);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem + r * N), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        }
        {   // narrow MMA: accumulator 0, columns 64..79 += A[row 1] * B2^T
            const uint64_t da = desc_mn_sw128(smem_u32(sA) + 1 * 1024, lbo, sbo);
            const uint64_t db = desc_kmajor(smem_u32(sB2), (NN / 8) * 128, 128);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem + 64), "l"(da), "l"(db), "r"(idesc2), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    done = 0;
    for (long spin = 0; spin < (1L << 24) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar_mma)) : "memory");
    if (!done) { if (tid == 0) *status = 1; return; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int cb = 0; cb < ROWS * N; cb += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) C[(long)m * ROWS * N + cb + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    if (tid == 0) *status = 0;
}

static float tf32(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    std::vector<float> A(8 * ROWS * W), B(N * 8), B2(NN * 8), C(M * ROWS * N), R(M * ROWS * N);
    srand(11);
    for (auto &x : A) x = (rand() % 2001 - 1000) / 500.0f;
    for (auto &x : B) x = (rand() % 2001 - 1000) / 500.0f;
    for (auto &x : B2) x = (rand() % 2001 - 1000) / 500.0f;
    auto a_at = [&](int k, int r, int x) { return (x < 0 || x >= W) ? 0.f : A[(k * ROWS + r) * W + x]; };
    for (int m = 0; m < M; ++m)
        for (int r = 0; r < ROWS; ++r)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < 8; ++k) s += (double)tf32(a_at(k, r, XS + m)) * tf32(B[n * 8 + k]);
                if (r == 0 && n >= 64)
                    for (int k = 0; k < 8; ++k) s += (double)tf32(a_at(k, 1, XS + m)) * tf32(B2[(n - 64) * 8 + k]);
                R[(m * ROWS + r) * N + n] = (float)s;
            }
    float *dA, *dB, *dB2, *dC;
    int *dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dB2, B2.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)W, 8, (cuuint64_t)ROWS};
    const cuuint64_t strides[2] = {(cuuint64_t)ROWS * W * 4, (cuuint64_t)W * 4};
    const cuuint32_t box[3] = {32, 8, (cuuint32_t)ROWS};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult cr = ((EncodeTiledFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
    const size_t smem = 1024 + 4 * ROWS * 1024 + (N + NN) * 32 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float *dDbg; cudaMalloc(&dDbg, 4 * ROWS * 256 * 4);
    std::vector<float> dbg(4 * ROWS * 256);
    int ok_any = 0;
    for (int variant = 0; variant < 2; ++variant) {
        int st = -1;
        cudaMemset(dC, 0, C.size() * 4);
        cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
        probe<<<1, 128, smem>>>(map, dB, dB2, dC, dS, variant, dDbg);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: cuda: %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(dbg.data(), dDbg, dbg.size() * 4, cudaMemcpyDeviceToHost);
        if (variant == 0) {
            // where did A[k][r][x] land?  expected: box w = (x - XS) / 32, offset ((r * 8 + k) * 32 + (x - XS) % 32) floats, 16-byte chunk index ^ (k % 8)
            long bad_plain = 0, bad_swz = 0, nz = 0;
            for (int w = 0; w < 4; ++w) for (int r = 0; r < ROWS; ++r) for (int k = 0; k < 8; ++k) for (int p = 0; p < 32; ++p) {
                const float want = a_at(k, r, XS + 32 * w + p);
                const int row = r * 8 + k, chunk = p / 4;
                const float plain = dbg[w * ROWS * 256 + row * 32 + p];
                const float swz = dbg[w * ROWS * 256 + row * 32 + (((p / 8) ^ (row % 4)) * 8) + p % 8];
                bad_plain += plain != want; bad_swz += swz != want; nz += plain != 0.f;
            }
            printf("smem dump: nonzero %ld  mismatches vs plain layout %ld  vs 128B/32B-atom swizzled layout %ld (of %d)\n", nz, bad_plain, bad_swz, 4 * ROWS * 256);
        }
        printf("C[lane 5][0..3] = %g %g %g %g   ref = %g %g %g %g\n", C[5 * ROWS * N], C[5 * ROWS * N + 1], C[5 * ROWS * N + 2], C[5 * ROWS * N + 3],
               R[5 * ROWS * N], R[5 * ROWS * N + 1], R[5 * ROWS * N + 2], R[5 * ROWS * N + 3]);
        double maxd = 0, maxr = 0, maxd_r[ROWS] = {0}, maxd_narrow = 0, maxd_left = 0;
        for (int m = 0; m < M; ++m)
            for (int r = 0; r < ROWS; ++r)
                for (int n = 0; n < N; ++n) {
                    const int i = (m * ROWS + r) * N + n;
                    const double d = fabs(C[i] - R[i]);
                    maxd = fmax(maxd, d);
                    maxr = fmax(maxr, fabs(R[i]));
                    maxd_r[r] = fmax(maxd_r[r], d);
                    if (r == 0 && n >= 64) maxd_narrow = fmax(maxd_narrow, d);
                    if (m < 4) maxd_left = fmax(maxd_left, fabs(C[i]));
                }
        printf("variant %d: status %d  max|C - ref| = %.4g (max|ref| %.4g)  per row %.3g %.3g %.3g  narrow cols %.3g  zero-fill lanes max|C| %.3g\n",
               variant, st, maxd, maxr, maxd_r[0], maxd_r[1], maxd_r[2], maxd_narrow, maxd_left);
        if (st == 0 && maxd < 2e-2 * maxr) ok_any |= 1 << variant;
    }
    printf("ok variants mask: %d\n", ok_any);
    return ok_any ? 0 : 1;
}
