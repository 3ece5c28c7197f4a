// kernels_head2.cuh -- the three head convolutions as a tcgen05 GEMM WITHOUT any re-packing of the activations, and
// with the propagation's prologue as its epilogue (first half; sm_100a only).
//
// kernels_head.cuh stages nine tap-shifted copies of every activation tile (K-major): ncu showed it bound by
// shared-memory bandwidth (72 KB of re-pack traffic + 45 KB of operand reads per 128 pixels x 8 channels), 4.6 ms at
// KITTI B = 8.  Here the activations are used as they lie in NCHW -- pixels contiguous = an MN-major A operand:
//
//   * one TMA box {32 px, 8 channels, R + 2 rows} (dims ordered x, channel, row, image; CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
//     IS a column of MMA atoms: tf32 MN-major operands take exactly one shared-memory layout, SWIZZLE_128B_BASE32B
//     (32 px x 4 channels per atom; LBO = bytes between the four 32-pixel boxes of a tile, SBO = 512 = the second
//     group of four channels).  Four boxes = 128 pixels = M.  tools/umma_mn_probe.cu pins all of this on the GPU.
//   * a row shift (dy) is a descriptor start address (+1024 bytes per row).  A column shift (dx) would be a 4-byte
//     shift, which no descriptor can express -- so it is moved to the OUTPUT side: the weights of the three dx taps
//     sit side by side in N (accumulator column 3 n + dx), the MMA runs on the unshifted tile, and the epilogue forms
//     D[px][n] = E[px - 1][3n] + E[px][3n + 1] + E[px + 1][3n + 2] with two warp shuffles.  A tile therefore loads
//     128 pixels and produces 124 of them (lanes 1 .. 124; x tiles step by 124, a multiple of four pixels because TMA
//     box starts must be 16-byte aligned); zero fill outside the image is the convolution's padding.
//   * per output pixel the tensor core reads 3 KB of activations (three dy) instead of 9 KB, nothing is re-packed,
//     and the block sparsity of the weights is used: the 64-channel branches of the init / confidence heads only feed
//     one output each, so their MMAs are N = 16 instead of N = NW.
//   * R output rows per tile (R accumulators side by side in TMEM, R * NW <= 256 columns) share R + 2 input rows.
//   * two kernels run this GEMM: head_rows_kernel (one CTA per tile, two per SM: warp 0 feeds a ring of TMA stages,
//     warp 1 issues the MMAs -- one elected lane each, the warps stay converged -- and all four warps run the epilogue)
//     and, the default, head_persist_kernel (one persistent CTA per SM: TMA warp, MMA warp, two groups of four
//     epilogue warps, two accumulator sets so that the epilogue of a tile overlaps the main loop of the next).
//   * epilogue: bias, relu / sigmoid, and -- when the caller passes the prologue's outputs -- _off_insert,
//     _affinity_normalization, _aff_insert, the confidence fix-up and the first blend + pre-multiply
//     (nlspnmodel.py:252-269,179-201,328-351; same expressions as prologue_fwd_kernel), so that `guidance` never
//     reaches HBM (it is still written when the caller asks for it: training needs it for the backward).
#pragma once
#include "kernels_head.cuh"
#include "kernels_v1.cuh"      // flag / affinity constants, blend_fix

namespace nlspn {

template <int K, int RING_ = 0>
struct HeadRows {
    static constexpr int KK = K * K, N = KK - 1, N3 = 3 * N, NOUT = N3 + 2, REF = N / 2;
    static constexpr int NW = (3 * NOUT + 15) / 16 * 16;          // accumulator columns per output row
    static constexpr int R = 256 / NW >= 3 ? 3 : 256 / NW;        // output rows per CTA (K = 3: 3, K = 5: 1)
    static constexpr int ROWS = R + 2;
    static constexpr int CF_BASE = (3 * (NOUT - 1)) / 16 * 16;    // first column of the confidence head's narrow MMA
    static constexpr int A_BOX = ROWS * 1024, A_BYTES = 4 * A_BOX;
    static constexpr int B_WIDE = 3 * NW * 32, B_NARROW = 3 * 16 * 32;     // bytes per stage: [dy][N x 8 tf32]
    static constexpr int B_SLOT = B_WIDE;                         // multiple of 128 bytes
    // ring depth of head_rows_kernel: the deepest that still lets two CTAs share an SM (228 KB minus 1 KB reserved per CTA)
    static constexpr int RING = RING_ > 0 ? RING_ : ((4 * (A_BYTES + B_SLOT) + 1024 + 1024 + 256) * 2 <= 228 * 1024 ? 4 : 3);
    static constexpr int TILE_OUT = 124;                          // lanes 1 .. 124 of a 128-pixel tile are outputs (see head_tiles_x)
    static constexpr int STAGES = 32, WIDE_STAGES = 16;           // (fe1, oa) x 8 chunks wide, (id, cf) x 8 chunks narrow
    static constexpr size_t smem = (size_t)RING * (A_BYTES + B_SLOT) + 1024;      // [RING A slots][RING B slots] + alignment slack
    static constexpr long packed_floats = (long)WIDE_STAGES * (B_WIDE + B_NARROW) / 4;
    static_assert(R >= 1 && R * NW <= 256 && NW <= 256, "accumulators must fit 256 TMEM columns");
    static_assert(3 * (NOUT - 1) + 2 < CF_BASE + 16, "the confidence columns must fit one N = 16 MMA");
    static_assert(2 * (smem + 1024 + 256) <= 228 * 1024, "two CTAs per SM");
    static_assert(B_SLOT % 128 == 0 && (size_t)2 * 4 * R * NOUT * 4 <= (size_t)RING * A_BYTES, "exchange buffer aliases the ring");
};

// Tile tx loads pixels 124 tx - 4 .. 124 tx + 123 (a 16-byte aligned start) and produces lanes 1 .. 124 = pixels
// 124 tx - 3 .. 124 tx + 120: every output lane has both neighbours in the tile, consecutive tiles abut, and the first
// tile's lanes 1 .. 3 fall left of the image.  (With outputs on lanes 4 .. 123 KITTI's 1216 columns took 11 tiles; now 10.)
__host__ __device__ constexpr int head_tiles_x(int W, int tile_out) { return (W + 3 + tile_out - 1) / tile_out; }

// stage -> source tensor: 0..7 fe1 (shared, feeds every output), 8..15 off_aff branch, 16..23 init branch, 24..31
// confidence branch.  The first stage must be a wide one (it initialises all accumulator columns).
__host__ __device__ constexpr int head_rows_source(int stage) { return stage < 8 ? 3 : stage < 16 ? 1 : stage < 24 ? 0 : 2; }

// Packs w_id [1,128,3,3], w_oa [3N,128,3,3], w_cf [1,128,3,3] into the per-stage B blocks: [dy][k-column 2][Ns / 8][8][4]
// (K-major un-swizzled core matrices, as kernels_head.cuh), column 3 n + dx of output n (0 = init, 1..3N = off_aff,
// 3N + 1 = confidence); channels 0..63 of a head = its own branch, 64..127 = fe1 (torch.cat((x_fd1, fe1), 1)).
template <int K>
__global__ void head_rows_pack_kernel(const float *__restrict__ w_id, const float *__restrict__ w_oa,
                                      const float *__restrict__ w_cf, float *__restrict__ packed)
{
    using C = HeadRows<K>;
    constexpr long kWideFloats = (long)C::WIDE_STAGES * C::B_WIDE / 4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < C::packed_floats; i += (long)gridDim.x * blockDim.x) {
        const bool wide = i < kWideFloats;
        const long li = wide ? i : i - kWideFloats;
        const int ns = wide ? C::NW : 16;
        const int per_stage = 3 * ns * 8;
        const int stage = (int)(li / per_stage) + (wide ? 0 : C::WIDE_STAGES);
        int t = (int)(li % per_stage);
        const int kk = t % 4; t /= 4;
        const int nr = t % 8; t /= 8;
        const int ng = t % (ns / 8); t /= (ns / 8);
        const int j = t % 2; t /= 2;
        const int dy = t;
        const int s = head_rows_source(stage);
        const int col = ng * 8 + nr + (wide ? 0 : (s == 2 ? C::CF_BASE : 0));
        const int n = col / 3, dx = col % 3;
        const int c = (stage % 8) * 8 + j * 4 + kk;
        const int cin = s == 3 ? kHeadCin + c : c;
        const int tap = dy * 3 + dx;
        float v = 0.f;
        if (n == 0) {
            if (s == 0 || s == 3) v = w_id[(long)cin * 9 + tap];
        } else if (n <= C::N3) {
            if (s == 1 || s == 3) v = w_oa[((long)(n - 1) * 2 * kHeadCin + cin) * 9 + tap];
        } else if (n == C::N3 + 1) {
            if (s == 2 || s == 3) v = w_cf[(long)cin * 9 + tap];
        }
        packed[i] = v;
    }
}

// MN-major tf32 operand: layout type 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t umma_desc_mn_tf32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return umma_desc_kmajor(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}

// one lane of a converged warp (cute::elect_one_sync)
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// tcgen05.mma kind::tf32 with the A-operand collector hint: consecutive MMAs on the SAME A tile keep it in the tensor
// core's collector buffer instead of re-reading 4 KB from shared memory (SASS: UTCHMMA ... .A_KEEP / .A_REUSE).
// MODE 0 = discard (default), 1 = fill (first of a run), 2 = use (middle), 3 = lastuse (last).
template <int MODE>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    if (MODE == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db),
                     "r"(idesc), "r"(accumulate) : "memory");
    else if (MODE == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32.collector::a::use [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db),
                     "r"(idesc), "r"(accumulate) : "memory");
    else if (MODE == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db),
                     "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc),
                     "r"(accumulate) : "memory");
}

// All MMAs of one stage, grouped by input row so that a row's A tile is read once: input row i feeds output rows
// r = i - dy (0 <= r < R, 0 <= dy < 3).  Per accumulator the order stays dy = 0, 1, 2.
template <class C, bool REUSE>
__device__ __forceinline__ void head_stage_mmas(uint32_t tacc, uint32_t sa, uint32_t sb, uint32_t ns, uint32_t idesc, uint32_t dcol,
                                                int st, uint32_t row_bytes = 1024, uint32_t box_bytes = C::A_BOX)
{
#pragma unroll
    for (int i = 0; i < C::ROWS; ++i) {
        const uint64_t da = umma_desc_mn_tf32(sa + i * row_bytes, box_bytes, 512);
        constexpr int R = C::R;
        const int dy_lo = i - (R - 1) > 0 ? i - (R - 1) : 0, dy_hi = i < 2 ? i : 2;      // dy range with 0 <= i - dy < R
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            if (dy < dy_lo || dy > dy_hi) continue;
            const int r = i - dy;
            const uint64_t db = umma_desc_kmajor(sb + dy * ns * 32, (ns / 8) * 128, 128);
            const uint32_t acc = (st | dy) != 0 ? 1u : 0u;
            const uint32_t d = tacc + r * C::NW + dcol;
            if (!REUSE || dy_lo == dy_hi) umma_tf32<0>(d, da, db, idesc, acc);
            else if (dy == dy_lo) umma_tf32<1>(d, da, db, idesc, acc);
            else if (dy == dy_hi) umma_tf32<3>(d, da, db, idesc, acc);
            else umma_tf32<2>(d, da, db, idesc, acc);
        }
    }
}

struct HeadRowsOut {
    float *pred_init, *confidence, *guidance;        // head outputs; guidance may be NULL when the prologue is fused
    // fused prologue (all NULL = heads only)
    const float *feat_fix, *gamma;
    float *offset, *aff, *conf_fixed, *src0;
    int affinity;
    unsigned flags;
};

template <int NCOLS>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, uint32_t (&v)[48])
{
#pragma unroll
    for (int c = 0; c < NCOLS; c += 16)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
                       "=r"(v[c + 6]), "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]),
                       "=r"(v[c + 12]), "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15])
                     : "r"(taddr + (uint32_t)c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct HeadRowsPixel { int x0, y0, b, H, W; long P; };

// Epilogue, part 1: what the neighbouring warps need from this warp's edge lanes.  `tq` = TMEM address of the
// accumulator set with this warp's lane quarter `quarter`; xch[0][quarter][r][n] = column 3n of lane 31,
// xch[1][quarter][r][n] = column 3n + 2 of lane 0.
template <class C>
__device__ __forceinline__ void head_rows_edge_columns(uint32_t tq, int quarter, int lane, int r,
                                                       float (*xch)[4][C::R][C::NOUT])
{
    constexpr int kChunks = (C::NW + 47) / 48;
    constexpr int kTail = C::NW - 48 * (kChunks - 1);
#pragma unroll
    for (int q = 0; q < kChunks; ++q) {
        uint32_t e[48];
        if (q + 1 < kChunks) tmem_load_cols<48>(tq + r * C::NW + q * 48, e);
        else tmem_load_cols<kTail>(tq + r * C::NW + q * 48, e);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = q * 16 + i;
            if (n < C::NOUT) {
                if (lane == 31) xch[0][quarter][r][n] = __uint_as_float(e[3 * i]);
                if (lane == 0) xch[1][quarter][r][n] = __uint_as_float(e[3 * i + 2]);
            }
        }
    }
}

template <class C>
__device__ __forceinline__ void head_rows_emit(const float (&d)[C::NOUT], int tl, int y, const HeadRowsPixel &t,
                                               const HeadRowsOut &o, float gamma);

// Epilogue, part 2: output row r of the tile.  D[px][n] = E[px - 1][3n] + E[px][3n + 1] + E[px + 1][3n + 2] + bias, the
// activations, and (fused) the propagation's prologue for that pixel.
template <class C>
__device__ __forceinline__ void head_rows_finish_row(uint32_t tq, int quarter, int lane, int r,
                                                     float (*xch)[4][C::R][C::NOUT], const HeadRowsPixel &t,
                                                     const float *__restrict__ bias, const HeadRowsOut &o, float gamma)
{
    constexpr int kChunks = (C::NW + 47) / 48;
    constexpr int kTail = C::NW - 48 * (kChunks - 1);
    const int tl = quarter * 32 + lane;                 // lane of the 128-pixel tile
    const int y = t.y0 + r;
    float d[C::NOUT];
#pragma unroll
    for (int q = 0; q < kChunks; ++q) {
        uint32_t e[48];
        if (q + 1 < kChunks) tmem_load_cols<48>(tq + r * C::NW + q * 48, e);
        else tmem_load_cols<kTail>(tq + r * C::NW + q * 48, e);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = q * 16 + i;
            if (n < C::NOUT) {
                float left = __shfl_up_sync(0xffffffffu, __uint_as_float(e[3 * i]), 1);
                float right = __shfl_down_sync(0xffffffffu, __uint_as_float(e[3 * i + 2]), 1);
                if (lane == 0) left = quarter > 0 ? xch[0][quarter - 1][r][n] : 0.f;
                if (lane == 31) right = quarter < 3 ? xch[1][quarter + 1][r][n] : 0.f;
                d[n] = ((__uint_as_float(e[3 * i + 1]) + left) + right) + __ldg(bias + n);
            }
        }
    }
    head_rows_emit<C>(d, tl, y, t, o, gamma);
}

// one output pixel: activations, the head outputs and (fused) the propagation's prologue from the NOUT pre-activations d[]
template <class C>
__device__ __forceinline__ void head_rows_emit(const float (&d)[C::NOUT], int tl, int y, const HeadRowsPixel &t,
                                               const HeadRowsOut &o, float gamma)
{
    const int px = t.x0 + tl;
    const int W = t.W;
    const long P = t.P;
    const bool fused = o.aff != nullptr;
    const bool preserve = fused && (o.flags & kPreserve) != 0;
    if (!(tl >= 1 && tl <= C::TILE_OUT && px >= 0 && px < W) || y >= t.H) return;
    const long q = (long)t.b * P + (long)y * W + px;
    const float init = fmaxf(d[0], 0.f);                                   // nlspnmodel.py:68 (relu)
    const float conf = 1.f / (1.f + expf(-d[C::N3 + 1]));                  // :83-86 (sigmoid)
    o.pred_init[q] = init;
    o.confidence[q] = conf;
    if (o.guidance) {
        float *gp = o.guidance + (long)t.b * C::N3 * P + (long)y * W + px;   // :81 (no activation)
#pragma unroll
        for (int n = 0; n < C::N3; ++n) gp[(long)n * P] = d[1 + n];
    }
    if (!fused) return;
    // ---- the propagation's prologue (same expressions as prologue_fwd_kernel, kernels_v1.cuh)
    const float dep = preserve ? __ldg(o.feat_fix + q) : 0.f;
    float *ob = o.offset + (long)t.b * 2 * C::KK * P + (long)y * W + px;
    float *ab = o.aff + (long)t.b * C::KK * P + (long)y * W + px;
#pragma unroll
    for (int tp = 0; tp < C::KK; ++tp) {
        if (tp == C::REF) {
            ob[(long)(2 * tp) * P] = 0.f;
            ob[(long)(2 * tp + 1) * P] = 0.f;
        } else {
            const int n = tp < C::REF ? tp : tp - 1;
            ob[(long)(2 * tp) * P] = d[1 + 2 * n];
            ob[(long)(2 * tp + 1) * P] = d[1 + 2 * n + 1];
        }
    }
    float a[C::N];
    float abs_sum = 0.f;
    const bool use_tanh = o.affinity == kTC || o.affinity == kTGASS;
    const float g = o.affinity == kTGASS ? gamma + 1e-8f : gamma;
#pragma unroll
    for (int n = 0; n < C::N; ++n) {
        float v = d[1 + 2 * C::N + n];
        if (use_tanh) v = tanhf(v) / g;
        a[n] = v;
        abs_sum += fabsf(v);
    }
    abs_sum += 1e-4f;
    if ((o.affinity == kASS || o.affinity == kTGASS) && abs_sum < 1.0f) abs_sum = 1.0f;
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < C::N; ++n) {
        if (o.affinity != kTC) a[n] = a[n] / abs_sum;
        sum += a[n];
    }
#pragma unroll
    for (int tp = 0; tp < C::KK; ++tp) ab[(long)tp * P] = tp == C::REF ? 1.0f - sum : a[tp < C::REF ? tp : tp - 1];
    float x = init, c = conf;
    if (preserve) x = blend_fix(x, dep);
    if (o.flags & kAlwaysClip) x = fmaxf(x, 0.f);
    if (o.conf_fixed) {
        if (preserve) {
            const float m = dep > 0.f ? 1.f : 0.f;
            c = (1.0f - m) * c + m;
        }
        o.conf_fixed[q] = c;
        x = x * c;
    }
    o.src0[q] = x;
}

// grid = (head_tiles_x(W), ceil(H / R), B), block = 128, dynamic shared memory HeadRows<K>::smem
template <int K, int RING_>
__global__ void __launch_bounds__(128, 2)
head_rows_kernel(const __grid_constant__ CUtensorMap map_id, const __grid_constant__ CUtensorMap map_oa,
                 const __grid_constant__ CUtensorMap map_cf, const __grid_constant__ CUtensorMap map_fe,
                 const float *__restrict__ packed, const float *__restrict__ bias, int H, int W, HeadRowsOut o)
{
    using C = HeadRows<K, RING_>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[C::RING], empty[C::RING];
    __shared__ uint32_t tmem_base_s;
    const uint32_t ring = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;       // [RING][A_BYTES] then [RING][B_SLOT]
    // the epilogue's exchange buffer aliases the ring (free once the last MMA has retired):
    // [0]: column 3n of lane 31, [1]: column 3n + 2 of lane 0, per warp
    float (*xch)[4][C::R][C::NOUT] = reinterpret_cast<float (*)[4][C::R][C::NOUT]>(smem_raw + (ring - tma::smem_u32(smem_raw)));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * C::TILE_OUT - 4, y0 = blockIdx.y * C::R;
    const int b = blockIdx.z;
    const long P = (long)H * W;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < C::RING; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], 1);
        }
        tma::fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // Warp-specialised main loop.  Warp 0 = TMA producer, warp 1 = MMA issuer; both run CONVERGED (all 32 lanes carry the
    // same loop state, one elected lane issues): a loop inside `if (tid == 0)` is a divergent region in which every
    // tcgen05 / TMA instruction gets a uniform-register waterfall (ELECT / PLOP3 / BRA.U.ANY per instruction; ncu: the
    // issuing thread spent 90 % of the main loop in those sequences, 1.43 ms per call at KITTI B = 8).
    if (warp == 0) {
        for (int st = 0; st < C::STAGES; ++st) {
            const int slot = st % C::RING;
            if (st >= C::RING) mbar_wait_bounded(&empty[slot], (uint32_t)((st / C::RING - 1) & 1));   // its MMAs have retired
            const int s = head_rows_source(st), c0 = (st & 7) * 8;
            const CUtensorMap *mp = s == 0 ? &map_id : s == 1 ? &map_oa : s == 2 ? &map_cf : &map_fe;
            const bool wide = st < C::WIDE_STAGES;
            const uint32_t sa = ring + slot * C::A_BYTES, sb = ring + C::RING * C::A_BYTES + slot * C::B_SLOT;
            const uint32_t bbytes = wide ? C::B_WIDE : C::B_NARROW;
            const float *wsrc = packed + (wide ? (long)st * (C::B_WIDE / 4)
                                               : (long)C::WIDE_STAGES * (C::B_WIDE / 4) + (long)(st - C::WIDE_STAGES) * (C::B_NARROW / 4));
            if (elect_one_sync()) {
                tma::mbar_arrive_expect_tx(&full[slot], C::A_BYTES + bbytes);
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                        ::"r"(sa + w * C::A_BOX), "l"(reinterpret_cast<uint64_t>(mp)), "r"(tma::smem_u32(&full[slot])),
                        "r"(x0 + 32 * w), "r"(c0), "r"(y0 - 1), "r"(b) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb),
                             "l"(wsrc), "r"(bbytes), "r"(tma::smem_u32(&full[slot])) : "memory");
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // D fp32, A = B = TF32, A MN-major (bit 15), B K-major, N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t kIdescWide = kIdescBase | ((uint32_t)(C::NW >> 3) << 17);
        constexpr uint32_t kIdescNarrow = kIdescBase | ((uint32_t)(16 >> 3) << 17);
        for (int st = 0; st < C::STAGES; ++st) {
            const int slot = st % C::RING;
            mbar_wait_bounded(&full[slot], (uint32_t)((st / C::RING) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool wide = st < C::WIDE_STAGES;
            const uint32_t ns = wide ? C::NW : 16;
            const uint32_t idesc = wide ? kIdescWide : kIdescNarrow;
            const uint32_t dcol = wide ? 0u : (head_rows_source(st) == 2 ? (uint32_t)C::CF_BASE : 0u);
            const uint32_t sa = ring + slot * C::A_BYTES, sb = ring + C::RING * C::A_BYTES + slot * C::B_SLOT;
            if (elect_one_sync()) {
                head_stage_mmas<C, false>(tmem, sa, sb, ns, idesc, dcol, st);     // (with two CTAs per SM the collector hints cost 10 %)
                // the barrier completes when every MMA issued so far has retired (also with its shared-memory reads)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 tma::smem_u32(&empty[slot])) : "memory");
            }
            __syncwarp();
        }
        mbar_wait_bounded(&empty[(C::STAGES - 1) % C::RING], (uint32_t)(((C::STAGES - 1) / C::RING) & 1));
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---------------- epilogue: lane = pixel x0 + tid; TMEM lane = tid
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int r = 0; r < C::R; ++r) head_rows_edge_columns<C>(tlane, warp, lane, r, xch);
    __syncthreads();
    const HeadRowsPixel px{x0, y0, b, H, W, P};
    const float gamma = o.aff != nullptr ? __ldg(o.gamma) : 1.f;
#pragma unroll 1
    for (int r = 0; r < C::R; ++r) head_rows_finish_row<C>(tlane, warp, lane, r, xch, px, bias, o, gamma);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

template <int K, int RING_ = 0>
inline cudaError_t head_rows_launch(const CUtensorMap &m_id, const CUtensorMap &m_oa, const CUtensorMap &m_cf,
                                    const CUtensorMap &m_fe, const float *packed, const float *bias, int B, int H, int W,
                                    const HeadRowsOut &o, cudaStream_t st)
{
    using C = HeadRows<K, RING_>;
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_rows_kernel<K, RING_>), (int)C::smem)) return ae;
    const dim3 grid((unsigned)head_tiles_x(W, C::TILE_OUT), (unsigned)((H + C::R - 1) / C::R), (unsigned)B);
    head_rows_kernel<K, RING_><<<grid, 128, C::smem, st>>>(m_id, m_oa, m_cf, m_fe, packed, bias, H, W, o);
    return cudaGetLastError();
}

// ======================================================================================
// The same GEMM as ONE persistent CTA per SM.  ncu on head_rows_kernel (two CTAs per SM): a CTA's life is ~60 % main
// loop and ~40 % epilogue, and the two CTAs of an SM run in lockstep, so the tensor core and the TMA ring idle during
// every epilogue.  Here the roles never stop: warp 0 streams TMA stages of tile after tile through one deep ring, warp 1
// issues the MMAs into one of TWO accumulator sets in tensor memory (2 x 256 columns), and NG groups of four epilogue
// warps drain the set the MMA warp finished last while it fills the other one (mbarriers acc_full / acc_empty).  With
// NG = 2 the rows of a tile are dealt to the two groups alternately, so that every scheduler holds two epilogue warps.
// grid = min(tiles, SMs), block = 64 + 128 NG, dynamic shared memory HeadPersist<K>::smem
// ======================================================================================
template <int K, int KS = 1>                                       // KS 8-channel K-steps per TMA stage
struct HeadPersist : HeadRows<K, 3> {
    using Base = HeadRows<K, 3>;
    static constexpr int NG = 2;    // epilogue groups of four warps: the rows of a tile are dealt alternately (K = 3); with R = 1 (K = 5)
                                    // they take alternate tiles (one group: 2.74 ms at KITTI B = 8, two: 2.18; three for K = 3: no gain)
    static constexpr int THREADS = 64 + 128 * NG;
    static constexpr int XCH_BYTES = 2 * 4 * Base::R * Base::NOUT * 4;              // per group
    static constexpr int PA_ROW = KS * 1024;                                        // one input row of one 32-pixel box: [8 KS ch][32 px]
    static constexpr int PA_BOX = Base::ROWS * PA_ROW, PA_BYTES = 4 * PA_BOX;
    static constexpr int PB_BYTES = KS * Base::B_SLOT;
    static constexpr int SLOT_BYTES = PA_BYTES + PB_BYTES;
    static constexpr int PSTAGES = Base::STAGES / KS, PWIDE = Base::WIDE_STAGES / KS;
    static constexpr int RING_MAX = (227 * 1024 - 1024 - NG * XCH_BYTES - 1024) / SLOT_BYTES;
    static constexpr int PRING = RING_MAX > 7 ? 7 : RING_MAX;
    static constexpr size_t smem = (size_t)PRING * SLOT_BYTES + NG * XCH_BYTES + 1024;
    static_assert(PRING >= 3 && 8 % KS == 0, "ring too shallow");
};

template <int K, int KS>
__global__ void __launch_bounds__(HeadPersist<K, KS>::THREADS, 1)
head_persist_kernel(const __grid_constant__ CUtensorMap map_id, const __grid_constant__ CUtensorMap map_oa,
                    const __grid_constant__ CUtensorMap map_cf, const __grid_constant__ CUtensorMap map_fe,
                    const float *__restrict__ packed, const float *__restrict__ bias, int H, int W, int tiles_x,
                    int tiles_y, int ntiles, int reuse, HeadRowsOut o)
{
    using C = HeadPersist<K, KS>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[C::PRING], empty[C::PRING], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t ring = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;       // [PRING][A_BYTES], [PRING][B_SLOT], [NG] exchange
    unsigned char *ring_p = smem_raw + (ring - tma::smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long P = (long)H * W;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < C::PRING; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], 1);
        }
        tma::mbar_init(&acc_full[0], 1);
        tma::mbar_init(&acc_full[1], 1);
        tma::mbar_init(&acc_empty[0], 128 * C::NG);
        tma::mbar_init(&acc_empty[1], 128 * C::NG);
        tma::fence_barrier_init();
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
        // ---- TMA producer: stage g of the CTA's whole tile sequence goes to ring slot g % PRING
        uint32_t g = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
            const int x0 = tx * C::TILE_OUT - 4, y0 = ty * C::R;
            for (int st = 0; st < C::PSTAGES; ++st, ++g) {
                const uint32_t slot = g % C::PRING;
                mbar_wait_bounded(&empty[slot], ((g / C::PRING) & 1u) ^ 1u);          // passes at once on the first lap
                const int st8 = st * KS;                                              // first 8-channel block of the stage
                const int s = head_rows_source(st8), c0 = (st8 & 7) * 8;
                const CUtensorMap *mp = s == 0 ? &map_id : s == 1 ? &map_oa : s == 2 ? &map_cf : &map_fe;
                const bool wide = st < C::PWIDE;
                const uint32_t sa = ring + slot * C::PA_BYTES, sb = ring + C::PRING * C::PA_BYTES + slot * C::PB_BYTES;
                const uint32_t bbytes = KS * (wide ? C::B_WIDE : C::B_NARROW);
                const float *wsrc = packed + (wide ? (long)st8 * (C::B_WIDE / 4)
                                                   : (long)C::WIDE_STAGES * (C::B_WIDE / 4) + (long)(st8 - C::WIDE_STAGES) * (C::B_NARROW / 4));
                if (elect_one_sync()) {
                    tma::mbar_arrive_expect_tx(&full[slot], C::PA_BYTES + bbytes);
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        asm volatile(
                            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                            ::"r"(sa + w * C::PA_BOX), "l"(reinterpret_cast<uint64_t>(mp)), "r"(tma::smem_u32(&full[slot])),
                            "r"(x0 + 32 * w), "r"(c0), "r"(y0 - 1), "r"(b) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb),
                                 "l"(wsrc), "r"(bbytes), "r"(tma::smem_u32(&full[slot])) : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t kIdescWide = kIdescBase | ((uint32_t)(C::NW >> 3) << 17);
        constexpr uint32_t kIdescNarrow = kIdescBase | ((uint32_t)(16 >> 3) << 17);
        uint32_t g = 0, i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
            const uint32_t buf = i & 1u;
            mbar_wait_bounded(&acc_empty[buf], ((i >> 1) & 1u) ^ 1u);                 // the epilogue has drained this set
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem + buf * 256u;
            for (int st = 0; st < C::PSTAGES; ++st, ++g) {
                const uint32_t slot = g % C::PRING;
                mbar_wait_bounded(&full[slot], (g / C::PRING) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool wide = st < C::PWIDE;
                const uint32_t ns = wide ? C::NW : 16;
                const uint32_t idesc = wide ? kIdescWide : kIdescNarrow;
                const uint32_t dcol = wide ? 0u : (head_rows_source(st * KS) == 2 ? (uint32_t)C::CF_BASE : 0u);
                const uint32_t sa = ring + slot * C::PA_BYTES, sb = ring + C::PRING * C::PA_BYTES + slot * C::PB_BYTES;
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) {
                        // K-step kk: channels 8 kk .. 8 kk + 7 of every row of the box, its own block of weights
                        if (reuse) head_stage_mmas<C, true>(tacc, sa + kk * 1024, sb + kk * (3 * ns * 32), ns, idesc, dcol, st * KS + kk, C::PA_ROW, C::PA_BOX);
                        else head_stage_mmas<C, false>(tacc, sa + kk * 1024, sb + kk * (3 * ns * 32), ns, idesc, dcol, st * KS + kk, C::PA_ROW, C::PA_BOX);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     tma::smem_u32(&empty[slot])) : "memory");
                    if (st == C::PSTAGES - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         tma::smem_u32(&acc_full[buf])) : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue groups
        const int grp = (warp - 2) >> 2, quarter = warp & 3;          // TMEM lane quarter of a warp = warp id % 4
        float (*xch)[4][C::R][C::NOUT] =
            reinterpret_cast<float (*)[4][C::R][C::NOUT]>(ring_p + (size_t)C::PRING * C::SLOT_BYTES + (size_t)grp * C::XCH_BYTES);
        const float gamma = o.aff != nullptr ? __ldg(o.gamma) : 1.f;
        uint32_t i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
            const HeadRowsPixel px{tx * C::TILE_OUT - 4, ty * C::R, b, H, W, P};
            const uint32_t buf = i & 1u;
            mbar_wait_bounded(&acc_full[buf], (i >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tq = tmem + buf * 256u + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
            for (int r = 0; r < C::R; ++r)
                if ((int)((r + i) % C::NG) == grp) head_rows_edge_columns<C>(tq, quarter, lane, r, xch);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll 1
            for (int r = 0; r < C::R; ++r)
                if ((int)((r + i) % C::NG) == grp) head_rows_finish_row<C>(tq, quarter, lane, r, xch, px, bias, o, gamma);
            // this thread's reads of the accumulator set are complete (tcgen05.wait::ld inside the loads)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma::smem_u32(&acc_empty[buf])) : "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");           // the exchange buffer is free again
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int K, int KS = 1>
inline cudaError_t head_persist_launch(const CUtensorMap &m_id, const CUtensorMap &m_oa, const CUtensorMap &m_cf,
                                       const CUtensorMap &m_fe, const float *packed, const float *bias, int B, int H, int W,
                                       int sm_count, bool reuse, const HeadRowsOut &o, cudaStream_t st)
{
    using C = HeadPersist<K, KS>;
    if (const cudaError_t ae = ensure_dynamic_smem(reinterpret_cast<const void *>(&head_persist_kernel<K, KS>), (int)C::smem)) return ae;
    const int tiles_x = head_tiles_x(W, C::TILE_OUT), tiles_y = (H + C::R - 1) / C::R;
    const long ntiles = (long)tiles_x * tiles_y * B;
    if (ntiles > 0x7fffffffL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(ntiles < sm_count ? ntiles : sm_count);
    head_persist_kernel<K, KS><<<grid, C::THREADS, C::smem, st>>>(m_id, m_oa, m_cf, m_fe, packed, bias, H, W, tiles_x, tiles_y,
                                                               (int)ntiles, reuse ? 1 : 0, o);
    return cudaGetLastError();
}

} // namespace nlspn
