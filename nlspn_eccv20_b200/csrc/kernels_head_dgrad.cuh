// kernels_head_dgrad.cuh -- the two WIDE data gradients of the head convolutions as one tcgen05 implicit GEMM (training path
// of SURVEY 8f row f3; in the reference autograd's cudnn_convolution_backward_input of nlspnmodel.py:73-82 [off_aff_dec0] and of
// the fe1 halves of all three heads, :69-86):
//     d_oa_fd1[b, c, y, x] = sum over (n >= 2, dy, dx) of g[b, n, y - dy + 1, x - dx + 1] * w_oa[n - 2, c, dy, dx]
//     d_fe1   [b, c, y, x] = sum over (n, dy, dx)      of g[b, n, y - dy + 1, x - dx + 1] * w_n[64 + c, dy, dx]
// with g = the concatenated upstream gradient (channel 0 = init, 1 = confidence, 2.. = guidance; prop_kernel 3: 26 channels,
// padded to 32 by TMA's zero fill).  Per tile of 128 pixels of one row: D[px][128 columns = 64 oa + 64 fe1 channels] over
// K = 9 taps x 32 gradient channels = 36 MMAs (M128 N128 K8, kind::tf32).
//   * A = the gradient as it lies (pixels contiguous = MN-major), boxes {32 px, 8 channels, 3 rows} in the
//     SWIZZLE_128B_ATOM_32B layout exactly as kernels_head2.cuh; the vertical tap is the box row, and the horizontal tap needs
//     no trick here: head_grad_prep_kernel's three shifted copies ARE the three dx operands.
//   * B = the weights, K-major core matrices, 144 KB, loaded ONCE per persistent CTA and resident in shared memory next to
//     a ring of six 12 KB gradient stages (stage = 8 channels of one copy).
//   * two accumulator sets: the epilogue of a tile (128 coalesced channel-plane stores per thread) overlaps the next tile.
#pragma once
#include "kernels_head2.cuh"

namespace nlspn {

struct HeadDgrad {
    static constexpr int NCH = 32, GROUPS = NCH / 8;             // gradient channels (padded), 8-channel K-steps
    static constexpr int NCOL = 2 * kHeadCin;                    // output channels: oa branch | fe1
    static constexpr int R = 2, ROWS = R + 2;                    // output rows per tile, gradient rows they read
    static constexpr int A_BOX = ROWS * 1024, A_STAGE = 4 * A_BOX;   // box = [ROWS rows][8 ch][32 px]; four boxes = 128 pixels
    static constexpr int W_BLOCK = NCOL * 8 * 4;                 // weights of one (tap, channel group): [2][NCOL / 8][8][4] floats
    static constexpr int W_BYTES = 9 * GROUPS * W_BLOCK;         // 147456
    static constexpr int RING = 5, STAGES = GROUPS * 3;          // stages per tile: (channel group, copy)
    static constexpr int THREADS = 192;                          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
    static constexpr size_t smem = (size_t)W_BYTES + (size_t)RING * A_STAGE + 1024;
    static constexpr long packed_floats = W_BYTES / 4;
};

// packed[tap = dy * 3 + dx][group q][k-column j][column group ng][column row nr][kk]: weight of output column ng * 8 + nr
// (0..63 = input channel c of the guidance head's own branch, 64..127 = fe1 channel c) for gradient channel q * 8 + j * 4 + kk
__global__ void head_dgrad_pack_kernel(const float *__restrict__ w_id, const float *__restrict__ w_oa, const float *__restrict__ w_cf,
                                       int NT, float *__restrict__ packed)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < HeadDgrad::packed_floats; i += (long)gridDim.x * blockDim.x) {
        int t = (int)i;
        const int kk = t % 4; t /= 4;
        const int nr = t % 8; t /= 8;
        const int ng = t % (HeadDgrad::NCOL / 8); t /= (HeadDgrad::NCOL / 8);
        const int j = t % 2; t /= 2;
        const int q = t % HeadDgrad::GROUPS; t /= HeadDgrad::GROUPS;
        const int tap = t;
        const int col = ng * 8 + nr, n = q * 8 + j * 4 + kk;
        const int cin = col;                                     // channel inside the 128-channel concatenation: own branch | fe1
        float v = 0.f;
        if (n < NT) {
            if (n == 0) { if (col >= kHeadCin) v = w_id[(long)cin * 9 + tap]; }
            else if (n == 1) { if (col >= kHeadCin) v = w_cf[(long)cin * 9 + tap]; }
            else v = w_oa[((long)(n - 2) * 2 * kHeadCin + cin) * 9 + tap];
        }
        packed[i] = v;
    }
}

// map_g: g_shift [3 B images][NT channels][H][W] with dims (x, channel, row, image), box {32, 8, ROWS, 1}, SWIZZLE_128B_ATOM_32B.
// A tile = 128 pixels x R = 2 output rows: the ROWS = 4 gradient rows of a stage feed 2 R + 2 = 6 MMAs, and the two inner
// rows are read ONCE for their two MMAs (A-operand collector: fill / lastuse).
// grid = min(tiles, SMs); d_oa may be NULL (that branch needs no gradient).
__global__ void __launch_bounds__(HeadDgrad::THREADS, 1)
head_dgrad_wide_kernel(const __grid_constant__ CUtensorMap map_g, const float *__restrict__ packed, int B, int H, int W, int tiles_x,
                       int tiles_y, int ntiles, float *__restrict__ d_oa, float *__restrict__ d_fe)
{
    using C = HeadDgrad;
    extern __shared__ unsigned char dgrad_smem_raw[];
    __shared__ __align__(8) uint64_t full[C::RING], empty[C::RING], acc_full[2], acc_empty[2], wbar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (tma::smem_u32(dgrad_smem_raw) + 1023u) & ~1023u;       // [weights][RING stages]
    const uint32_t wsm = base, ring = base + C::W_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long P = (long)H * W;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < C::RING; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], 1);
        }
        tma::mbar_init(&acc_full[0], 1);
        tma::mbar_init(&acc_full[1], 1);
        tma::mbar_init(&acc_empty[0], 128);
        tma::mbar_init(&acc_empty[1], 128);
        tma::mbar_init(&wbar, 1);
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&map_g);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ---- TMA producer: the weights once, then stage g of the CTA's tile sequence -> ring slot g % RING
        if (elect_one_sync()) {
            tma::mbar_arrive_expect_tx(&wbar, C::W_BYTES);
            for (int i = 0; i < 9; ++i)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(wsm + i * (C::W_BYTES / 9)), "l"(packed + (long)i * (C::W_BYTES / 36)), "r"(C::W_BYTES / 9),
                             "r"(tma::smem_u32(&wbar)) : "memory");
        }
        __syncwarp();
        uint32_t g = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int tx = t % tiles_x, y = ((t / tiles_x) % tiles_y) * C::R, b = t / (tiles_x * tiles_y);
            for (int st = 0; st < C::STAGES; ++st, ++g) {
                const uint32_t slot = g % C::RING;
                mbar_wait_bounded(&empty[slot], ((g / C::RING) & 1u) ^ 1u);
                const int q = st / 3, d = st % 3;
                const uint32_t sa = ring + slot * C::A_STAGE;
                if (elect_one_sync()) {
                    tma::mbar_arrive_expect_tx(&full[slot], C::A_STAGE);
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        asm volatile(
                            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                            ::"r"(sa + w * C::A_BOX), "l"(reinterpret_cast<uint64_t>(&map_g)), "r"(tma::smem_u32(&full[slot])),
                            "r"(tx * 128 + 32 * w), "r"(q * 8), "r"(y - 1), "r"(d * B + b) : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: D fp32, A tf32 MN-major, B tf32 K-major, M = 128, N = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(C::NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mbar_wait_bounded(&wbar, 0u);
        uint32_t g = 0, i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
            const uint32_t buf = i & 1u;
            mbar_wait_bounded(&acc_empty[buf], ((i >> 1) & 1u) ^ 1u);                 // the epilogue has drained this set
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem + buf * (uint32_t)(C::R * C::NCOL);
            for (int st = 0; st < C::STAGES; ++st, ++g) {
                const uint32_t slot = g % C::RING;
                mbar_wait_bounded(&full[slot], (g / C::RING) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int q = st / 3, d = st % 3;
                const uint32_t sa = ring + slot * C::A_STAGE;
                if (elect_one_sync()) {
                    // box row i = gradient row y0 - 1 + i feeds output row r with the tap dy = r + 2 - i (0 <= dy <= 2); copy d = dx
                    const uint32_t wq = wsm + q * C::W_BLOCK;
                    const uint32_t first = st != 0 ? 1u : 0u;
#define DGRAD_MMA(MODE, i, r, accum)                                                                                         \
    umma_tf32<MODE>(tacc + (r) * C::NCOL, umma_desc_mn_tf32(sa + (i) * 1024, C::A_BOX, 512),                                  \
                    umma_desc_kmajor(wq + (((r) + 2 - (i)) * 3 + d) * C::GROUPS * C::W_BLOCK, (C::NCOL / 8) * 128, 128), idesc, accum)
                    DGRAD_MMA(0, 0, 0, first);       // row 0: dy = 2 of output row 0 (its first MMA of the tile when st == 0)
                    DGRAD_MMA(1, 1, 0, 1u);          // row 1: dy = 1 of output row 0 ...
                    DGRAD_MMA(3, 1, 1, first);       //        ... and dy = 2 of output row 1 (its first MMA), same A
                    DGRAD_MMA(1, 2, 0, 1u);          // row 2: dy = 0 of output row 0 ...
                    DGRAD_MMA(3, 2, 1, 1u);          //        ... and dy = 1 of output row 1
                    DGRAD_MMA(0, 3, 1, 1u);          // row 3: dy = 0 of output row 1
#undef DGRAD_MMA
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     tma::smem_u32(&empty[slot])) : "memory");
                    if (st == C::STAGES - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         tma::smem_u32(&acc_full[buf])) : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue: lane = pixel, column = output channel; one coalesced 128-byte store per warp and channel plane
        const int quarter = warp & 3;
        uint32_t i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
            const int tx = t % tiles_x, y0 = ((t / tiles_x) % tiles_y) * C::R, b = t / (tiles_x * tiles_y);
            const int x = tx * 128 + quarter * 32 + lane;
            const uint32_t buf = i & 1u;
            mbar_wait_bounded(&acc_full[buf], (i >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tq = tmem + buf * (uint32_t)(C::R * C::NCOL) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int cb = 0; cb < C::R * C::NCOL; cb += 32) {
                const int y = y0 + cb / C::NCOL, col = cb % C::NCOL;
                if (y >= H) break;                                   // warp-uniform: the second row of the last tile of an odd H
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 32; c += 16)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
                                   "=r"(v[c + 6]), "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]),
                                   "=r"(v[c + 12]), "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15])
                                 : "r"(tq + (uint32_t)(cb + c)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float *out = col < kHeadCin ? d_oa : d_fe;
                if (out && x < W) {
                    out += ((long)b * kHeadCin + (col & (kHeadCin - 1))) * P + (long)y * W + x;
#pragma unroll
                    for (int c = 0; c < 32; ++c) __stcs(out + (long)c * P, __uint_as_float(v[c]));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma::smem_u32(&acc_empty[buf])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

inline cudaError_t head_dgrad_wide_launch(const CUtensorMap &map_g, const float *packed, int B, int H, int W, float *d_oa, float *d_fe,
                                          int sm_count, cudaStream_t st)
{
    using C = HeadDgrad;
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_dgrad_wide_kernel), (int)C::smem)) return ae;
    const int tiles_x = (W + 127) / 128, tiles_y = (H + C::R - 1) / C::R;
    const long ntiles = (long)tiles_x * tiles_y * B;
    if (ntiles > 0x7fffffffL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(ntiles < sm_count ? ntiles : sm_count);
    head_dgrad_wide_kernel<<<grid, C::THREADS, C::smem, st>>>(map_g, packed, B, H, W, tiles_x, tiles_y, (int)ntiles, d_oa, d_fe);
    return cudaGetLastError();
}

} // namespace nlspn
