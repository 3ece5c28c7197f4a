// kernels_head.cuh -- the three final head convolutions in front of the propagation as ONE tcgen05 implicit GEMM
// (first half; sm_100a only).
//
// Reference (nlspnmodel.py:69-86,297,301,313; builders common.py:45-67):
//     pred_init  = relu   (conv3x3(cat(id_fd1, fe1), 128 -> 1))
//     off_aff    =         conv3x3(cat(oa_fd1, fe1), 128 -> 3N)          ("guidance": 2N offsets | N raw affinities)
//     confidence = sigmoid(conv3x3(cat(cf_fd1, fe1), 128 -> 1))
// Stock PyTorch on B200 (KITTI 352x1216, B = 8, TF32 allowed, NCHW): 1.7 ms for the three torch.cat, 2.4 ms for the
// 128 -> 24 convolution alone, 6.75 ms in total (tools/head_baseline.py) -- longer than our whole propagation
// forward + backward.
//
// Here: no concatenation (the four 64-channel tensors are read where they lie), one GEMM
//     D[pixel, out] = sum over (source s, tap, channel c)  X_s[c, y + dy, x + dx] * Wp[(s, tap, c), out]
// with M = 128 pixels of one image row per CTA, N = NP outputs (3N + 2 padded to a multiple of 16; the weight matrix
// is block-sparse: the init / confidence columns only see their own decoder branch and the shared fe1), K = 4 x 9 x 64
// = 2304, issued as tcgen05.mma.cta_group::1.kind::tf32 (M 128 x N NP x K 8) with the fp32 accumulator in tensor
// memory -- TF32 is also what cuDNN computes these convolutions in by default (torch.backends.cudnn.allow_tf32).
//
// A operand: for every (source, 8-channel chunk) stage the 256 threads build the NINE tap-shifted [128 px x 8 ch]
// tiles straight from global memory in the K-major un-swizzled core-matrix layout (8 rows x 16 B per core matrix):
// lane = pixel, so the four channel loads of a thread are warp-coalesced 128-byte rows (unaligned by one pixel for
// dx = +-1; 8 of the 9 taps hit L1) and its one 16-byte shared-memory store lands conflict-free.  Image borders are
// predicated loads (zero padding).  B operand: the host packs the weights once per call into the same layout, one
// contiguous block per stage.  Two stage buffers: the tensor core works on stage i while the threads build stage
// i + 1; tcgen05.commit -> mbarrier hands a buffer back.  Epilogue: warps 0-3 read their 32 TMEM lanes (tcgen05.ld
// 32x32b), add the bias, apply relu / sigmoid and write pred_init, guidance, confidence as coalesced NCHW rows.
//
// tools/umma_probe.cu pins the descriptor encodings on the actual GPU.
#pragma once
#include <mutex>

#include "common.cuh"
#include "tma.cuh"

namespace nlspn {

constexpr int kHeadCin = 64;             // channels of each of the four source tensors
constexpr int kHeadTM = 128;             // pixels per CTA (one image row segment) = MMA M
constexpr int kHeadThreads = 256;
constexpr int kHeadChunk = 8;            // channels per stage = MMA K (tf32)
constexpr int kHeadStages = 4 * (kHeadCin / kHeadChunk);      // 32

// outputs padded to the MMA's N granularity for M = 128
__host__ __device__ constexpr int head_np(int K) { return ((3 * (K * K - 1) + 2) + 15) / 16 * 16; }
// floats of the packed weight matrix: [stage][tap][k-column (2)][NP / 8][8][4]
__host__ __device__ constexpr long head_packed_floats(int K) { return (long)kHeadStages * 9 * kHeadChunk * head_np(K); }

// K-major, no swizzle (cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE): the 16-byte unit of row r,
// k-column j sits at (r % 8) + (r / 8) * SBO + j * LBO, in units of 16 bytes
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

// Packs the three weight tensors [1,128,3,3], [3N,128,3,3], [1,128,3,3] into the B-operand blocks.  Output column
// 0 = init, 1 .. 3N = off_aff, 3N + 1 = confidence, the rest zero.  Input channels 0..63 of each head are its own
// decoder branch (sources 0, 1, 2), 64..127 the shared fe1 (source 3): torch.cat((x_fd1, fe1), dim=1).
__global__ void head_pack_weights_kernel(const float *__restrict__ w_id, const float *__restrict__ w_oa,
                                         const float *__restrict__ w_cf, int N3, int NP, float *__restrict__ packed)
{
    const long total = (long)kHeadStages * 9 * kHeadChunk * NP;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        // i -> (stage, tap, j, ng, nr, kk)
        long t = i;
        const int kk = t % 4; t /= 4;
        const int nr = t % 8; t /= 8;
        const int ng = t % (NP / 8); t /= (NP / 8);
        const int j = t % 2; t /= 2;
        const int tap = t % 9; t /= 9;
        const int stage = (int)t;
        const int s = stage / (kHeadCin / kHeadChunk), c = (stage % (kHeadCin / kHeadChunk)) * kHeadChunk + j * 4 + kk;
        const int n = ng * 8 + nr;
        float v = 0.f;
        const int cin = s == 3 ? kHeadCin + c : c;           // channel inside the 128-channel concatenation
        if (n == 0) {
            if (s == 0 || s == 3) v = w_id[(long)cin * 9 + tap];
        } else if (n <= N3) {
            if (s == 1 || s == 3) v = w_oa[((long)(n - 1) * 2 * kHeadCin + cin) * 9 + tap];
        } else if (n == N3 + 1) {
            if (s == 2 || s == 3) v = w_cf[(long)cin * 9 + tap];
        }
        packed[i] = v;
    }
}

template <int NP>
struct HeadSmem {
    static constexpr int kATile = kHeadTM * kHeadChunk;                 // floats per tap tile (4 KB)
    static constexpr int kA = 9 * kATile;                               // floats per stage
    static constexpr int kB = 9 * kHeadChunk * NP;                      // floats per stage
    static constexpr size_t bytes = sizeof(float) * 2 * (kA + kB) + 64;
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(tma::smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// bounded: a tensor-core fault must surface as an error, not as a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    for (long spin = 0; spin < (1L << 26); ++spin)      // seconds: longer than any legitimate wait (persistent kernels included)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// epilogue shared by both kernels: warps 0-3 own TMEM lanes 32 w .. 32 w + 31 = pixels; 32 columns per tcgen05.ld
template <int NP>
__device__ __forceinline__ void head_epilogue(uint32_t tmem, int warp, int tid, int x0, int y, long b, long P, int N3, int W,
                                              const float *__restrict__ bias, float *__restrict__ pred_init,
                                              float *__restrict__ guidance, float *__restrict__ confidence)
{
    if (warp < 4) {
        const int px = x0 + warp * 32 + (tid & 31);
        const bool valid = px < W;
        const long q = b * P + (long)y * W + px;
        const int NOUT = N3 + 2;
#pragma unroll
        for (int cb = 0; cb < NP; cb += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
            if (NP - cb >= 32) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                      "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                      "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
            } else {     // 16 columns left (NP = 80: 32 + 32 + 16)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0u;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = cb + i;
                    if (n >= NOUT) break;
                    const float o = __uint_as_float(v[i]) + __ldg(bias + n);
                    if (n == 0) pred_init[q] = fmaxf(o, 0.f);                                   // nlspnmodel.py:68 (relu)
                    else if (n <= N3) guidance[(b * N3 + (n - 1)) * P + (long)y * W + px] = o;  // :81 (no activation)
                    else confidence[q] = 1.f / (1.f + expf(-o));                                // :83-86 (sigmoid)
                }
            }
        }
    }
}

// grid = (ceil(W / 128), H, B), block = 256, dynamic shared memory HeadSmem<NP>::bytes
template <int NP>
__global__ void __launch_bounds__(kHeadThreads, 2)
head_fused_kernel(const float *__restrict__ id_fd1, const float *__restrict__ oa_fd1,
                  const float *__restrict__ cf_fd1, const float *__restrict__ fe1,
                  const float *__restrict__ packed, const float *__restrict__ bias, int N3, int H, int W,
                  float *__restrict__ pred_init, float *__restrict__ guidance, float *__restrict__ confidence)
{
    using S = HeadSmem<NP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sA = reinterpret_cast<float *>(smem_raw);                    // [2][9][128 x 8]
    float *sB = sA + 2 * S::kA;                                         // [2][9][NP x 8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + 2 * S::kB);      // empty[0], empty[1], done
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int x0 = blockIdx.x * kHeadTM, y = blockIdx.y;
    const long b = blockIdx.z;
    const long P = (long)H * W;
    constexpr uint32_t kCols = NP <= 32 ? 32 : NP <= 64 ? 64 : NP <= 128 ? 128 : 256;     // TMEM allocation
    if (tid == 0) {
        tma::mbar_init(&bars[0], 1);
        tma::mbar_init(&bars[1], 1);
        tma::mbar_init(&bars[2], 1);
        tma::fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(&tmem_base_s)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A = B = TF32, both K-major, N >> 3, M >> 4
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(kHeadTM >> 4) << 24);

    const int m = tid & (kHeadTM - 1);         // pixel of the tile this thread stages
    const int j = tid >> 7;                    // k-column: channels 4j .. 4j + 3 of the chunk
    const int xm = x0 + m;
    const float *srcs[4] = {id_fd1, oa_fd1, cf_fd1, fe1};
    uint32_t wait_parity[2] = {0u, 0u};        // parity the next wait on empty[s] expects

    for (int stage = 0; stage < kHeadStages; ++stage) {
        const int buf = stage & 1;
        if (stage >= 2) {                      // the MMAs of stage - 2 must have read this buffer
            mbar_wait_bounded(&bars[buf], wait_parity[buf]);
            wait_parity[buf] ^= 1u;
        }
        const int s = stage / (kHeadCin / kHeadChunk), c0 = (stage % (kHeadCin / kHeadChunk)) * kHeadChunk + 4 * j;
        const float *xs = srcs[s] + (b * kHeadCin + c0) * P;
        float *a_buf = sA + buf * S::kA;
        // ---- A: nine tap-shifted tiles; this thread: pixel m, channels c0 .. c0 + 3 -> one 16-byte store per tap
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = xm + tap % 3 - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
                const float *p = xs + (long)yy * W + xx;
                v.x = __ldg(p);
                v.y = __ldg(p + P);
                v.z = __ldg(p + 2 * P);
                v.w = __ldg(p + 3 * P);
            }
            // tile [k-column j][row group m / 8][8 rows][4 floats]
            *reinterpret_cast<float4 *>(a_buf + tap * S::kATile + ((j * (kHeadTM / 8) + (m >> 3)) * 8 + (m & 7)) * 4) = v;
        }
        // ---- B: this stage's packed block, copied as it lies
        {
            const float4 *gsrc = reinterpret_cast<const float4 *>(packed + (long)stage * S::kB);
            float4 *bdst = reinterpret_cast<float4 *>(sB + buf * S::kB);
            for (int i = tid; i < S::kB / 4; i += kHeadThreads) bdst[i] = __ldg(gsrc + i);
        }
        tma::fence_proxy_async_smem();         // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = tma::smem_u32(a_buf), b0 = tma::smem_u32(sB + buf * S::kB);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint64_t da = umma_desc_kmajor(a0 + tap * S::kATile * 4, (kHeadTM / 8) * 128, 128);
                const uint64_t db = umma_desc_kmajor(b0 + tap * kHeadChunk * NP * 4, (NP / 8) * 128, 128);
                const uint32_t acc = (stage | tap) != 0 ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db),
                             "r"(idesc), "r"(acc) : "memory");
            }
            // commit: the barrier completes when every MMA issued so far has finished (also with its smem reads)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             tma::smem_u32(stage + 1 < kHeadStages ? &bars[buf] : &bars[2])) : "memory");
        }
    }
    mbar_wait_bounded(&bars[2], 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    head_epilogue<NP>(tmem, warp, tid, x0, y, b, P, N3, W, bias, pred_init, guidance, confidence);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols));
}

// ======================================================================================
// The same GEMM with the activations delivered by TMA (W % 4 == 0).  ncu on head_fused_kernel at KITTI B = 8: 9.5 ms,
// DRAM 5 %, every warp parked on the shared-memory store that waits for its global loads -- one exposed memory round
// trip per stage and only 16 warps per SM to hide it.  Here one thread keeps a ring of kHeadRing stages in flight:
// per stage one 4-D TMA box {136 px, 3 rows, 8 channels} of the source tensor (zero fill outside the image = the
// convolution's padding) and one bulk copy of the stage's packed weights straight into their MMA layout; the 256
// threads only re-pack the box into the nine K-major tap tiles (shared -> shared) and the tensor core reads the weights
// from the ring slot.  One mbarrier per ring slot (bytes landed), one for "MMAs of the previous stage retired".
// grid = (ceil(W / 128), H, B), block = 256, dynamic shared memory HeadSmemTma<NP>::bytes
// ======================================================================================
constexpr int kHeadRawW = 136;           // 4 + 128 + 4 pixels: the box must START on a 16-byte boundary (x0 - 4, not x0 - 1)
                                         // and a box row must be a multiple of 32 bytes
                                         // (528-byte rows make cp.async.bulk.tensor trap with error 715 on B200, like the
                                         // 208-byte rows of DESIGN.md decision 6)
constexpr int kHeadRing = 3;

template <int NP>
struct HeadSmemTma {
    static constexpr int kRaw = kHeadChunk * 3 * kHeadRawW;            // floats per stage: [8 ch][3 rows][136 px]
    static constexpr int kB = 9 * kHeadChunk * NP;
    static constexpr int kATile = kHeadTM * kHeadChunk;
    static constexpr int kA = 9 * kATile;
    static constexpr uint32_t kStageBytes = sizeof(float) * (kRaw + kB);
    static constexpr size_t bytes = sizeof(float) * (kHeadRing * (kRaw + kB) + kA) + 8 * (kHeadRing + 1) + 128;
    static_assert((kRaw * 4) % 128 == 0 && (kB * 4) % 128 == 0, "ring slots stay 128-byte aligned");
};

template <int NP>
__global__ void __launch_bounds__(kHeadThreads, 2)
head_fused_tma_kernel(const __grid_constant__ CUtensorMap map_id, const __grid_constant__ CUtensorMap map_oa,
                      const __grid_constant__ CUtensorMap map_cf, const __grid_constant__ CUtensorMap map_fe,
                      const float *__restrict__ packed, const float *__restrict__ bias, int N3, int H, int W,
                      float *__restrict__ pred_init, float *__restrict__ guidance, float *__restrict__ confidence)
{
    using S = HeadSmemTma<NP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sRaw = reinterpret_cast<float *>(smem_raw);                  // [ring][8][3][136]
    float *sB = sRaw + kHeadRing * S::kRaw;                             // [ring][9][NP x 8]
    float *sA = sB + kHeadRing * S::kB;                                 // [9][128 x 8]
    uint64_t *full = reinterpret_cast<uint64_t *>(sA + S::kA);          // [ring]
    uint64_t *mma_bar = full + kHeadRing;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int x0 = blockIdx.x * kHeadTM, y = blockIdx.y;
    const int b = blockIdx.z;
    const long P = (long)H * W;
    constexpr uint32_t kCols = NP <= 32 ? 32 : NP <= 64 ? 64 : NP <= 128 ? 128 : 256;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kHeadRing; ++i) tma::mbar_init(&full[i], 1);
        tma::mbar_init(mma_bar, 1);
        tma::fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(&tmem_base_s)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(kHeadTM >> 4) << 24);

    // stage -> (source tensor, first channel); slot = stage % ring
    auto issue = [&](int stage) {          // thread 0 only
        const int slot = stage % kHeadRing;
        const int s = stage / (kHeadCin / kHeadChunk), c0 = (stage % (kHeadCin / kHeadChunk)) * kHeadChunk;
        const CUtensorMap *mp = s == 0 ? &map_id : s == 1 ? &map_oa : s == 2 ? &map_cf : &map_fe;
        tma::mbar_arrive_expect_tx(&full[slot], S::kStageBytes);
        tma::load_4d(sRaw + slot * S::kRaw, mp, &full[slot], x0 - 4, y - 1, c0, b);   // start column: a multiple of 16 bytes
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         tma::smem_u32(sB + slot * S::kB)), "l"(packed + (long)stage * S::kB), "r"((uint32_t)(S::kB * 4)),
                     "r"(tma::smem_u32(&full[slot])) : "memory");
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kHeadRing; ++i) issue(i);
    }
    const int m = tid & (kHeadTM - 1), j = tid >> 7;

    for (int stage = 0; stage < kHeadStages; ++stage) {
        const int slot = stage % kHeadRing;
        if (stage >= 1) {
            // the MMAs of the previous stage have retired: the A tiles and that stage's weight slot are free again
            mbar_wait_bounded(mma_bar, (uint32_t)((stage - 1) & 1));
            if (tid == 0 && stage - 1 + kHeadRing < kHeadStages) issue(stage - 1 + kHeadRing);
        }
        mbar_wait_bounded(&full[slot], (uint32_t)((stage / kHeadRing) & 1));
        // ---- re-pack: thread (pixel m, k-column j) moves channels 4j .. 4j+3 of nine shifted pixels
        const float *raw = sRaw + slot * S::kRaw + (4 * j) * 3 * kHeadRawW + m;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float *p = raw + (tap / 3) * kHeadRawW + (tap % 3) + 3;     // box column 0 = image column x0 - 4
            const float4 v = make_float4(p[0], p[3 * kHeadRawW], p[6 * kHeadRawW], p[9 * kHeadRawW]);
            *reinterpret_cast<float4 *>(sA + tap * S::kATile + ((j * (kHeadTM / 8) + (m >> 3)) * 8 + (m & 7)) * 4) = v;
        }
        tma::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = tma::smem_u32(sA), b0 = tma::smem_u32(sB + slot * S::kB);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint64_t da = umma_desc_kmajor(a0 + tap * S::kATile * 4, (kHeadTM / 8) * 128, 128);
                const uint64_t db = umma_desc_kmajor(b0 + tap * kHeadChunk * NP * 4, (NP / 8) * 128, 128);
                const uint32_t acc = (stage | tap) != 0 ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db),
                             "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             tma::smem_u32(mma_bar)) : "memory");
        }
    }
    mbar_wait_bounded(mma_bar, (uint32_t)((kHeadStages - 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    head_epilogue<NP>(tmem, warp, tid, x0, y, b, P, N3, W, bias, pred_init, guidance, confidence);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols));
}

// host side: opt in to the dynamic shared memory once per instantiation, launch
template <int NP>
inline cudaError_t head_launch(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                               const float *packed, const float *bias, int B, int H, int W, int N3, float *pred_init,
                               float *guidance, float *confidence, cudaStream_t st)
{
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_fused_kernel<NP>), (int)HeadSmem<NP>::bytes)) return ae;
    const dim3 grid((unsigned)((W + kHeadTM - 1) / kHeadTM), (unsigned)H, (unsigned)B);
    head_fused_kernel<NP><<<grid, kHeadThreads, HeadSmem<NP>::bytes, st>>>(id_fd1, oa_fd1, cf_fd1, fe1, packed, bias, N3, H, W,
                                                                           pred_init, guidance, confidence);
    return cudaGetLastError();
}

template <int NP>
inline cudaError_t head_launch_tma(const CUtensorMap &m_id, const CUtensorMap &m_oa, const CUtensorMap &m_cf,
                                   const CUtensorMap &m_fe, const float *packed, const float *bias, int B, int H, int W,
                                   int N3, float *pred_init, float *guidance, float *confidence, cudaStream_t st)
{
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_fused_tma_kernel<NP>), (int)HeadSmemTma<NP>::bytes)) return ae;
    const dim3 grid((unsigned)((W + kHeadTM - 1) / kHeadTM), (unsigned)H, (unsigned)B);
    head_fused_tma_kernel<NP><<<grid, kHeadThreads, HeadSmemTma<NP>::bytes, st>>>(m_id, m_oa, m_cf, m_fe, packed, bias, N3, H,
                                                                                  W, pred_init, guidance, confidence);
    return cudaGetLastError();
}

} // namespace nlspn
