// umma_probe.cu -- smallest possible tcgen05 TF32 GEMM (tool): C[128 x 32] = A[128 x KT] * B[32 x KT]^T with both
// operands in shared memory in the K-major, un-swizzled core-matrix layout, accumulator in TMEM, one CTA.
// It pins down, on the actual B200, the pieces kernels_head.cuh is built from: shared-memory descriptors (LBO / SBO),
// the instruction descriptor of kind::tf32, tcgen05.alloc / mma / commit / ld.  Prints max |C - C_ref|.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 32, KT = 64;        // KT / 8 MMAs of K = 8

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 16-byte unit (row r, k-column j) at  (r % 8) + (r / 8) * SBO + j * LBO   [units of 16 B]
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version 1 (Blackwell)
    return d;                                    // base offset 0, layout type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *C, int *status)
{
    // A: [KT/4 k-columns][16 row groups][8 rows][4 floats]   LBO = 16 * 128 B, SBO = 128 B
    // B: [KT/4 k-columns][ 4 row groups][8 rows][4 floats]   LBO =  4 * 128 B, SBO = 128 B
    __shared__ __align__(128) float sA[M * KT];
    __shared__ __align__(128) float sB[N * KT];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * KT; i += 128) {
        const int m = i / KT, k = i % KT;
        sA[((k / 4) * (M / 8) + m / 8) * 32 + (m % 8) * 4 + (k % 4)] = A[i];
    }
    for (int i = tid; i < N * KT; i += 128) {
        const int n = i / KT, k = i % KT;
        sB[((k / 4) * (N / 8) + n / 8) * 32 + (n % 8) * 4 + (k % 4)] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3, M >> 4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k8 = 0; k8 < KT / 8; ++k8) {
            const uint64_t da = make_desc(smem_u32(sA) + k8 * 2 * (M / 8) * 128, (M / 8) * 128, 128);
            const uint64_t db = make_desc(smem_u32(sB) + k8 * 2 * (N / 8) * 128, (N / 8) * 128, 128);
            const uint32_t acc = k8 > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db),
                "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait (a wrong descriptor must not hang the box)
    uint32_t done = 0;
    for (long spin = 0; spin < (1L << 24) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (!done) {
        if (tid == 0) *status = 1;
        return;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);      // lane field in bits 16.., column in the low bits
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int n = 0; n < N; ++n) C[m * N + n] = __uint_as_float(v[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
    if (tid == 0) *status = 0;
}

static float tf32(float x)      // round to 10 mantissa bits (nearest even); the tensor core truncates or rounds: both
{                               // stay inside the tolerance below
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

int main()
{
    std::vector<float> A(M * KT), B(N * KT), C(M * N), R(M * N);
    srand(7);
    for (auto &x : A) x = (rand() % 2001 - 1000) / 500.0f;
    for (auto &x : B) x = (rand() % 2001 - 1000) / 500.0f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < KT; ++k) s += (double)tf32(A[m * KT + k]) * tf32(B[n * KT + k]);
            R[m * N + n] = (float)s;
        }
    float *dA, *dB, *dC;
    int *dS, st = -1;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0, C.size() * 4);
    cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxd = 0, maxr = 0;
    for (int i = 0; i < M * N; ++i) { maxd = fmax(maxd, fabs(C[i] - R[i])); maxr = fmax(maxr, fabs(R[i])); }
    printf("cuda: %s  status: %d  max|C - ref| = %.4g (max|ref| = %.4g)  C[0..3] = %g %g %g %g  ref = %g %g %g %g\n",
           cudaGetErrorString(e), st, maxd, maxr, C[0], C[1], C[2], C[3], R[0], R[1], R[2], R[3]);
    return (e == cudaSuccess && st == 0 && maxd < 2e-2 * maxr) ? 0 : 1;
}
