// kernels_head_wgrad.cuh -- weight gradient of the three final 3x3 head convolutions on tcgen05 (training path of SURVEY 8f
// row f3; the layers are /root/reference/src/model/nlspnmodel.py:69-86,297,301,313, their backward is autograd's
// cudnn_convolution_backward_weight in the reference).
//
//     dW[n, c, dy, dx] = sum over (b, y, x) of  g[b, n, y, x] * X[b, c, y + dy - 1, x + dx - 1]          (X zero-padded)
//
// as a split-K GEMM over PIXELS: a persistent CTA per SM walks over 32-pixel chunks (b, y, x0) of the input X and
// accumulates  D[(tap, n), c] += G_tap[(tap, n), px] * X[c, px]  in tensor memory: M = 128 rows = (tap slot, gradient
// channel), N = 64 input channels, K = 32 pixels = 4 x (K = 8 tf32).  Pixels are contiguous in NCHW, so BOTH operands are
// K-major as they lie and arrive by TMA in the 128-byte swizzle; the vertical tap shift is a row coordinate of the box.
// The horizontal shift is one pixel = 4 bytes along K, which neither a shared-memory descriptor nor a TMA box start
// (16-byte aligned) can express: head_grad_prep_kernel writes the upstream gradient three times, shifted by +1 / 0 / -1
// pixels (the centre copy is what the data gradient reads; the same kernel applies the ReLU / Sigmoid derivatives of the
// init / confidence heads, concatenates the three gradients and sums the bias gradients).
//   NP = 32: four tap slots of 32 gradient channels per MMA, three MMAs cover the nine taps (the 3N + 2 = 26 channels of
//            K = 3 in one launch; wider heads in blocks of 32 channels);
//   NP = 8 : sixteen tap slots of 8 channels per MMA, ONE MMA covers the nine taps (the one-channel init / confidence heads).
// Split-K partials of the CTAs are added to dW with fp32 atomics (order-dependent in the last bits, like cuDNN's default
// weight-gradient algorithms).  Developed in tools/wgrad_probe.cu.
#pragma once
#include "kernels_head2.cuh"

namespace nlspn {

template <int NP>
struct HeadWgrad {
    static constexpr int SLOTS = 128 / NP;                       // tap slots per MMA (M = 128)
    static constexpr int MMAS = (9 + SLOTS - 1) / SLOTS;         // MMAs per K-step: 3 (NP = 32) or 1 (NP = 8)
    static constexpr int XT = kHeadCin * 128;                    // bytes of an X tile [64 channels][32 px]
    static constexpr int GT = NP * 128;                          // bytes of one tap slot [NP channels][32 px]
    static constexpr int STAGE = XT + MMAS * SLOTS * GT;         // X + every slot an MMA reads (9 of them loaded)
    static constexpr int RING = NP == 32 ? 3 : 6;
    static constexpr int THREADS = 192;                          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
    static constexpr uint32_t TMEM_COLS = NP == 32 ? 256u : 64u; // MMAS x 64 accumulator columns, a power of two
    static constexpr size_t smem = (size_t)RING * STAGE + 1024;
    static_assert(NP == 8 || NP == 32, "tap slots are whole 8-row swizzle atoms and divide M = 128");
};

// K-major operand in the 128-byte swizzle: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO); LBO unused
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// Upstream gradients of the three heads -> g_shift [3][B][NT][H][W], NT = 3N + 2 (channel 0 = init, 1 = confidence,
// 2 .. 3N + 1 = guidance -- the two one-channel heads side by side, so that one 8-channel box serves both):
// copy d holds g[.., x - d + 1] (zero outside the row).  Init: g * (pred_init > 0) (ReLU, nlspnmodel.py
// :69-72); confidence: g * c (1 - c) (Sigmoid, :83-86).  A NULL gradient is a zero gradient.  g_bias[n] += sum of channel n.
// grid (plane chunks, NT, B), one thread per four pixels; W % 4 == 0.
constexpr int kGradPrepThreads = 256;
constexpr int kGradPrepQuads = 2048;              // quads of four pixels per block

__device__ __forceinline__ float head_grad_value(const float *g, const float *act, int mode, long i)
{
    const float v = __ldg(g + i);
    if (mode == 0) return v;
    const float a = __ldg(act + i);
    return mode == 1 ? (a > 0.f ? v : 0.f) : __fmul_rn(v, __fmul_rn(a, 1.f - a));     // the same roundings as the vector path below
}

__global__ void __launch_bounds__(kGradPrepThreads)
head_grad_prep_kernel(const float *__restrict__ g_init, const float *__restrict__ pred_init, const float *__restrict__ g_guid,
                      const float *__restrict__ g_conf, const float *__restrict__ confidence, int B, int NT, int H, int W,
                      float *__restrict__ g_shift, float *__restrict__ g_bias)
{
    const int n = blockIdx.y, b = blockIdx.z, W4 = W >> 2;
    const long P = (long)H * W, quads = (long)H * W4;
    const float *g, *act = nullptr;
    int mode = 0;
    if (n == 0) { g = g_init ? g_init + (long)b * P : nullptr; act = pred_init + (long)b * P; mode = 1; }
    else if (n == 1) { g = g_conf ? g_conf + (long)b * P : nullptr; act = confidence + (long)b * P; mode = 2; }
    else g = g_guid ? g_guid + ((long)b * (NT - 2) + (n - 2)) * P : nullptr;
    const long copy = (long)B * NT * P;
    float *out = g_shift + ((long)b * NT + n) * P;
    float sum = 0.f;
    const long q0 = (long)blockIdx.x * kGradPrepQuads;
    for (long q = q0 + threadIdx.x; q < q0 + kGradPrepQuads && q < quads; q += kGradPrepThreads) {
        const int x0 = (int)(q % W4) * 4;
        const long i = q * 4;                                   // = y * W + x0
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        float left = 0.f, right = 0.f;
        if (g) {
            c = __ldg(reinterpret_cast<const float4 *>(g + i));
            if (mode) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(act + i));
                if (mode == 1) {
                    c.x = a.x > 0.f ? c.x : 0.f; c.y = a.y > 0.f ? c.y : 0.f; c.z = a.z > 0.f ? c.z : 0.f; c.w = a.w > 0.f ? c.w : 0.f;
                } else {
                    c.x = __fmul_rn(c.x, __fmul_rn(a.x, 1.f - a.x)); c.y = __fmul_rn(c.y, __fmul_rn(a.y, 1.f - a.y));
                    c.z = __fmul_rn(c.z, __fmul_rn(a.z, 1.f - a.z)); c.w = __fmul_rn(c.w, __fmul_rn(a.w, 1.f - a.w));
                }
            }
            if (x0 > 0) left = head_grad_value(g, act, mode, i - 1);
            if (x0 + 4 < W) right = head_grad_value(g, act, mode, i + 4);
        }
        reinterpret_cast<float4 *>(out + i)[0] = make_float4(c.y, c.z, c.w, right);              // copy 0: g[x + 1]
        reinterpret_cast<float4 *>(out + copy + i)[0] = c;                                       // copy 1: g[x]
        reinterpret_cast<float4 *>(out + 2 * copy + i)[0] = make_float4(left, c.x, c.y, c.z);    // copy 2: g[x - 1]
        sum += (c.x + c.y) + (c.z + c.w);
    }
    if (!g_bias) return;
    __shared__ float part[kGradPrepThreads / 32];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < kGradPrepThreads / 32; ++i) s += part[i];
        atomicAdd(g_bias + n, s);
    }
}

// Data gradient of the two ONE-channel heads (init, confidence) with respect to their own 64-channel branches:
//     dX[b, c, y, x] = sum over (dy, dx) of  g[b, y - dy + 1, x - dx + 1] * w[c, dy, dx]       (g zero-padded)
// 9 multiply-adds per output from one gradient channel: a plain fp32 stencil that is bound by its 64-channel store (what
// cuDNN spends 0.7-0.8 ms on per head at KITTI B = 8 streams out at the HBM rate).  g_all = the centre copy of
// head_grad_prep_kernel (channel 0 = init, 1 = confidence); w = the head's [1,128,3,3] weight, input channels 0..63.
// grid (quads of four pixels / 256, 2 heads), one thread per four pixels; W % 4 == 0.
constexpr int kDgradOneThreads = 256;

__global__ void __launch_bounds__(kDgradOneThreads)
head_dgrad_one_kernel(const float *__restrict__ g_all, const float *__restrict__ w_id, const float *__restrict__ w_cf, int B, int NT,
                      int H, int W, float *__restrict__ d_id, float *__restrict__ d_cf)
{
    const int head = blockIdx.y;
    float *__restrict__ out = head == 0 ? d_id : d_cf;
    if (!out) return;
    __shared__ float ws[kHeadCin * 9];
    const float *w = head == 0 ? w_id : w_cf;
    for (int i = threadIdx.x; i < kHeadCin * 9; i += kDgradOneThreads) ws[i] = __ldg(w + i);     // [c][dy][dx], c < 64
    __syncthreads();
    const int W4 = W >> 2;
    const long P = (long)H * W, quads = (long)B * H * W4;
    const long q = (long)blockIdx.x * kDgradOneThreads + threadIdx.x;
    if (q >= quads) return;
    const int x0 = (int)(q % W4) * 4, y = (int)((q / W4) % H), b = (int)(q / ((long)W4 * H));
    const float *g = g_all + ((long)b * NT + head) * P;
    // v[r][i] = g[y - 1 + r][x0 - 1 + i], r = 0..2, i = 0..5
    float v[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int yy = y - 1 + r;
        if (yy >= 0 && yy < H) {
            const float *row = g + (long)yy * W + x0;
            const float4 c = __ldg(reinterpret_cast<const float4 *>(row));
            v[r][0] = x0 > 0 ? __ldg(row - 1) : 0.f;
            v[r][1] = c.x; v[r][2] = c.y; v[r][3] = c.z; v[r][4] = c.w;
            v[r][5] = x0 + 4 < W ? __ldg(row + 4) : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) v[r][i] = 0.f;
        }
    }
    float *o = out + (long)b * kHeadCin * P + (long)y * W + x0;
#pragma unroll 4
    for (int c = 0; c < kHeadCin; ++c) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float wv = ws[c * 9 + dy * 3 + dx];
                // output pixel x0 + i reads g[y - dy + 1][x0 + i - dx + 1] = v[2 - dy][i - dx + 2]
                acc.x = fmaf(v[2 - dy][2 - dx], wv, acc.x);
                acc.y = fmaf(v[2 - dy][3 - dx], wv, acc.y);
                acc.z = fmaf(v[2 - dy][4 - dx], wv, acc.z);
                acc.w = fmaf(v[2 - dy][5 - dx], wv, acc.w);
            }
        __stcs(reinterpret_cast<float4 *>(o + (long)c * P), acc);
    }
}

// One input tensor X [B, 64, H, W] against gradient channels n0 .. n0 + n_cnt - 1 (n_cnt <= NP) of g_shift:
//     dW[(n0 + n) * ldw + c * 9 + dy * 3 + dx] += ...      (dW already offset to X's channel block; ldw = 128 * 9)
// map_x: dims (x, channel, row, image), box {32, 64, 1, 1}; map_g: dims (x, channel, row, copy * B + image), box {32, NP, 3, 1};
// both SWIZZLE_128B, zero fill outside the tensor = the convolution's padding.  grid = min(chunks, SMs), 192 threads.
template <int NP>
__global__ void __launch_bounds__(HeadWgrad<NP>::THREADS, 1)
head_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, int B, int H, int W,
                  int n0, int n_cnt, int ldw, float *__restrict__ dW)
{
    using C = HeadWgrad<NP>;
    extern __shared__ unsigned char wgrad_smem_raw[];
    __shared__ __align__(8) uint64_t full[C::RING], empty[C::RING], done_bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t ring = (tma::smem_u32(wgrad_smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = (W + 31) / 32;
    const long items = (long)B * H * chunks;
    if (tid == 0) {
        for (int i = 0; i < C::RING; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], 1);
        }
        tma::mbar_init(&done_bar, 1);
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&map_x);
        tma::prefetch_descriptor(&map_g);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(&tmem_base_s)), "r"(C::TMEM_COLS));
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
            const uint32_t slot = g % C::RING;
            mbar_wait_bounded(&empty[slot], ((g / C::RING) & 1u) ^ 1u);
            const uint32_t sx = ring + slot * C::STAGE, sg = sx + C::XT;
            if (elect_one_sync()) {
                tma::mbar_arrive_expect_tx(&full[slot], C::XT + 9 * C::GT);
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             ::"r"(sx), "l"(reinterpret_cast<uint64_t>(&map_x)), "r"(tma::smem_u32(&full[slot])), "r"(xc * 32), "r"(0), "r"(y), "r"(b) : "memory");
                // copy d (g[x - d + 1]) of rows y - 1 .. y + 1 -> slots d * 3 + r; row r holds tap dy = 2 - r, dx = d
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(sg + d * 3 * C::GT), "l"(reinterpret_cast<uint64_t>(&map_g)), "r"(tma::smem_u32(&full[slot])), "r"(xc * 32), "r"(n0),
                                 "r"(y - 1), "r"(d * B + b) : "memory");
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // D fp32, A = B = tf32, both K-major, N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kHeadCin >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t g = 0;
        for (long it = blockIdx.x; it < items; it += gridDim.x, ++g) {
            const uint32_t slot = g % C::RING;
            mbar_wait_bounded(&full[slot], (g / C::RING) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sx = ring + slot * C::STAGE, sg = sx + C::XT;
            if (elect_one_sync()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t db = umma_desc_k_sw128(sx + kk * 32);
#pragma unroll
                    for (int m = 0; m < C::MMAS; ++m)
                        umma_tf32<0>(tmem + m * kHeadCin, umma_desc_k_sw128(sg + m * C::SLOTS * C::GT + kk * 32), db, idesc,
                                     (g | (uint32_t)kk) != 0 ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(&empty[slot])) : "memory");
            }
            __syncwarp();
        }
        if (elect_one_sync())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(&done_bar)) : "memory");
        __syncwarp();
    } else {
        // accumulator m: lane = (tap slot within the MMA, gradient channel), 64 columns = input channels
        mbar_wait_bounded(&done_bar, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;                           // the TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane, tl = row / NP, n = row % NP;
        for (int m = 0; m < C::MMAS; ++m) {
            const int s = m * C::SLOTS + tl;                    // tap slot = d * 3 + r
            uint32_t v[64];
#pragma unroll
            for (int cb = 0; cb < 64; cb += 16) {
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * kHeadCin + cb);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[cb + 0]), "=r"(v[cb + 1]), "=r"(v[cb + 2]), "=r"(v[cb + 3]), "=r"(v[cb + 4]), "=r"(v[cb + 5]),
                               "=r"(v[cb + 6]), "=r"(v[cb + 7]), "=r"(v[cb + 8]), "=r"(v[cb + 9]), "=r"(v[cb + 10]), "=r"(v[cb + 11]),
                               "=r"(v[cb + 12]), "=r"(v[cb + 13]), "=r"(v[cb + 14]), "=r"(v[cb + 15])
                             : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (s < 9 && n < n_cnt) {
                const int d = s / 3, r = s % 3;
                float *out = dW + (long)(n0 + n) * ldw + (2 - r) * 3 + d;
#pragma unroll
                for (int c = 0; c < 64; ++c) atomicAdd(out + c * 9, __uint_as_float(v[c]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS));
}

template <int NP>
inline cudaError_t head_wgrad_launch(const CUtensorMap &map_x, const CUtensorMap &map_g, int B, int H, int W, int n0, int n_cnt,
                                     int ldw, float *dW, int sm_count, cudaStream_t st)
{
    using C = HeadWgrad<NP>;
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_wgrad_kernel<NP>), (int)C::smem)) return ae;
    const long items = (long)B * H * ((W + 31) / 32);
    const unsigned grid = (unsigned)(items < sm_count ? items : sm_count);
    head_wgrad_kernel<NP><<<grid, C::THREADS, C::smem, st>>>(map_x, map_g, B, H, W, n0, n_cnt, ldw, dW);
    return cudaGetLastError();
}

// ---- rolling form: a CTA walks DOWN a 32-pixel column strip, so that a gradient row is loaded once for the three input rows
// it meets (instead of three times), and NX input tensors share it (N = 64 NX columns per MMA).  Stage j of the ring holds
// {X tiles of input row y, the three shifted copies of gradient row y + 1}; the MMAs of row y read the gradient rows of stages
// j (dy = 0), j - 1 (dy = 1), j - 2 (dy = 2), one accumulator per dy.  Roles swapped against the per-chunk form: M = 128 rows =
// the input channels (A = X, shared by the three dy MMAs of a K-step: read once into the collector), N = 3 NP columns =
// (copy d = dx, gradient channel) -- tf32 MMAs with K = 8 read 32 bytes per operand row and are bound by shared-memory
// reads, not by the tensor core (measured: 110 cycles per M128 N128 K8 without the sharing).  Per step: 8 NX KB of X +
// 3 NP x 128 B of gradient instead of 8 KB + 9 NP x 128 B, and the ring is 7-11 stages deep instead of 3.
// Items = (image, strip, segment of SEG rows); an item starts with two gradient-only steps (rows y0 - 1 and y0).
struct HeadWgradX { int n_lo[2], n_hi[2], c_off[2]; };       // per input tensor: valid gradient channels, first dW input channel

template <int NP, int NX>
struct HeadWgradRoll {
    static constexpr int XT = kHeadCin * 128 * NX;               // bytes of the X tiles of a stage: [64 NX channels][32 px]
    static constexpr int GT = NP * 128;                          // one shifted copy of a gradient row [NP channels][32 px]
    static constexpr int STAGE = XT + 3 * GT;
    static constexpr int NCOL = NP == 32 ? 96 : 32;              // MMA N: the 3 NP (copy, gradient channel) rows, rounded up to 16
    // the last stage's MMAs read 128 A rows from its X base (NX = 1: 64 of them are whatever follows) and NCOL B rows from its
    // gradient base: what of that lies beyond the stage must still be shared memory
    static constexpr int PAD = (STAGE < 128 * 128 ? 128 * 128 - STAGE : 0) + (NCOL * 128 - 3 * GT);
    static constexpr int RING = (232448 - 3072 - PAD) / STAGE < 12 ? (232448 - 3072 - PAD) / STAGE : 12;   // 227 KB - alignment, static
    static constexpr int THREADS = 192;
    static constexpr uint32_t TMEM_COLS = NP == 32 ? 512u : 128u; // 3 accumulators x NCOL columns, a power of two
    static constexpr size_t smem = (size_t)RING * STAGE + PAD + 1024;
    static_assert(RING >= 4, "three stages are in use by the MMAs of one step");
};

__host__ __device__ constexpr int head_wgrad_seg_rows(int H) { return H <= 32 ? H : (H + (H + 31) / 32 - 1) / ((H + 31) / 32); }

// map_x*: dims (x, channel, row, image), box {32, 64, 1, 1}; map_g: dims (x, channel, copy, row, image), box {32, NP, 3, 1, 1}
template <int NP, int NX>
__global__ void __launch_bounds__(HeadWgradRoll<NP, NX>::THREADS, 1)
head_wgrad_roll_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                       const __grid_constant__ CUtensorMap map_g, int B, int H, int W, int seg, int n0, HeadWgradX xs, int ldw,
                       float *__restrict__ dW)
{
    using C = HeadWgradRoll<NP, NX>;
    extern __shared__ unsigned char wgrad_smem_raw[];
    __shared__ __align__(8) uint64_t full[C::RING], empty[C::RING], done_bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t ring = (tma::smem_u32(wgrad_smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunks = (W + 31) / 32, nseg = (H + seg - 1) / seg;
    const long items = (long)B * nseg * chunks;
    if (tid == 0) {
        for (int i = 0; i < C::RING; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], 1);
        }
        tma::mbar_init(&done_bar, 1);
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&map_x0);
        if (NX == 2) tma::prefetch_descriptor(&map_x1);
        tma::prefetch_descriptor(&map_g);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(&tmem_base_s)), "r"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        uint32_t j = 0;
        for (long it = blockIdx.x; it < items; it += gridDim.x) {
            const int xc = (int)(it % chunks), sg = (int)((it / chunks) % nseg), b = (int)(it / ((long)chunks * nseg));
            const int y0 = sg * seg, rows = min(seg, H - y0);
            for (int k = 0; k < rows + 2; ++k, ++j) {
                const uint32_t slot = j % C::RING;
                mbar_wait_bounded(&empty[slot], ((j / C::RING) & 1u) ^ 1u);
                const uint32_t sx = ring + slot * C::STAGE;
                if (elect_one_sync()) {
                    tma::mbar_arrive_expect_tx(&full[slot], (k >= 2 ? C::XT : 0) + 3 * C::GT);
                    if (k >= 2) {
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                     ::"r"(sx), "l"(reinterpret_cast<uint64_t>(&map_x0)), "r"(tma::smem_u32(&full[slot])), "r"(xc * 32), "r"(0),
                                     "r"(y0 + k - 2), "r"(b) : "memory");
                        if (NX == 2)
                            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                         ::"r"(sx + kHeadCin * 128), "l"(reinterpret_cast<uint64_t>(&map_x1)), "r"(tma::smem_u32(&full[slot])),
                                         "r"(xc * 32), "r"(0), "r"(y0 + k - 2), "r"(b) : "memory");
                    }
                    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                                 ::"r"(sx + C::XT), "l"(reinterpret_cast<uint64_t>(&map_g)), "r"(tma::smem_u32(&full[slot])), "r"(xc * 32), "r"(n0),
                                 "r"(0), "r"(y0 - 1 + k), "r"(b) : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // D fp32, A = X tiles [128 input channels][32 px], B = gradient tiles [NCOL (copy, channel) rows][32 px], both tf32 K-major.
        // The three MMAs of a K-step share A: it stays in the tensor core's collector (fill / use / lastuse), so a K-step reads
        // 4 KB + 3 x NCOL x 32 B of shared memory instead of three times both operands.
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C::NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t j = 0, started = 0;
        for (long it = blockIdx.x; it < items; it += gridDim.x) {
            const int sg = (int)((it / chunks) % nseg);
            const int rows = min(seg, H - sg * seg);
            for (int k = 0; k < rows + 2; ++k, ++j) {
                const uint32_t slot = j % C::RING;
                mbar_wait_bounded(&full[slot], (j / C::RING) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one_sync()) {
                    if (k >= 2) {
                        const uint32_t sx = ring + slot * C::STAGE;
                        const uint32_t sb0 = sx + C::XT, sb1 = ring + ((j + C::RING - 1) % C::RING) * C::STAGE + C::XT,
                                       sb2 = ring + ((j + C::RING - 2) % C::RING) * C::STAGE + C::XT;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t da = umma_desc_k_sw128(sx + kk * 32);
                            const uint32_t acc = (started | (uint32_t)kk) != 0 ? 1u : 0u;
                            umma_tf32<1>(tmem + 0 * C::NCOL, da, umma_desc_k_sw128(sb0 + kk * 32), idesc, acc);
                            umma_tf32<2>(tmem + 1 * C::NCOL, da, umma_desc_k_sw128(sb1 + kk * 32), idesc, acc);
                            umma_tf32<3>(tmem + 2 * C::NCOL, da, umma_desc_k_sw128(sb2 + kk * 32), idesc, acc);
                        }
                    }
                    // stage j - 2 (its gradient row was this step's dy = 2 operand) is free once these MMAs have read it
                    if (j >= 2)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     ::"r"(tma::smem_u32(&empty[(j - 2) % C::RING])) : "memory");
                }
                __syncwarp();
                if (k >= 2) started = 1;
            }
        }
        if (elect_one_sync())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(&done_bar)) : "memory");
        __syncwarp();
    } else {
        // accumulator dy: lane = input channel (tensor t = lane / 64), column = copy d (= dx) x NP + gradient channel
        mbar_wait_bounded(&done_bar, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;                           // the TMEM lane quarter this warp may read
        const int t = quarter >> 1, c = (quarter & 1) * 32 + lane;
        if (t < NX) {
            float *out = dW + (xs.c_off[t] + c) * 9;
            const int lo = xs.n_lo[t], hi = xs.n_hi[t];
            for (int dy = 0; dy < 3; ++dy)
                for (int cb = 0; cb < 3 * NP; cb += (NP == 32 ? 16 : 8)) {
                    uint32_t v[16];
                    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(dy * C::NCOL + cb);
                    if (NP == 32)
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                     : "r"(taddr));
                    else
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                     : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int d = cb / NP, nb = n0 + cb % NP;
#pragma unroll
                    for (int i = 0; i < (NP == 32 ? 16 : 8); ++i)
                        if (nb + i >= lo && nb + i < hi) atomicAdd(out + (long)(nb + i) * ldw + dy * 3 + d, __uint_as_float(v[i]));
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS));
}

template <int NP, int NX>
inline cudaError_t head_wgrad_roll_launch(const CUtensorMap &map_x0, const CUtensorMap &map_x1, const CUtensorMap &map_g, int B, int H,
                                          int W, int n0, const HeadWgradX &xs, int ldw, float *dW, int sm_count, cudaStream_t st)
{
    using C = HeadWgradRoll<NP, NX>;
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_wgrad_roll_kernel<NP, NX>), (int)C::smem)) return ae;
    const int seg = head_wgrad_seg_rows(H);
    const long items = (long)B * ((H + seg - 1) / seg) * ((W + 31) / 32);
    const unsigned grid = (unsigned)(items < sm_count ? items : sm_count);
    head_wgrad_roll_kernel<NP, NX><<<grid, C::THREADS, C::smem, st>>>(map_x0, map_x1, map_g, B, H, W, seg, n0, xs, ldw, dW);
    return cudaGetLastError();
}

} // namespace nlspn
