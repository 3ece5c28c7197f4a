// kernels_local.cuh -- pass A of the backward as a CTA-LOCAL TRANSPOSE through shared memory (sm_100a, K = 3).
//
// Where the RED form stands (kernels_v2.cuh, ncu in profiles/): every tap of every pixel ships one 16-byte
// vector RED = one 32-byte sector slot on the SM->L2 write port; 8 taps x 3.4 M pixels = 27 M sectors per
// launch is 1.29 cycles per lane (B300_MICROARCH "REDG spread") = ~330 cycles per warp and iteration, i.e.
// 125-141 us per launch whatever the rest of the kernel does (tools/red_bench.cu), with DRAM at 42 %.
// Shared-memory float atomics are CAS loops (2 cycles per lane), so accumulating the scatter in a shared tile
// loses as well (tools/smem_scatter_bench.cu).
//
// This form needs NO atomics inside the SM.  The sampling geometry is the same for all T iterations, so WHERE
// every (pixel, tap) footprint lands is known before the first iteration:
//
//   sched_build_kernel (once per backward call; grid = 32x8-pixel tiles)
//       Every footprint that lies inside the tile's halo'd region (tile + 8 px on each side, the same box
//       the forward gathers from) is one 2x2 block of one of four block phases of that region.  The kernel
//       counts the footprints per block (shared-memory integer atomics -- slow, but once per step), sorts the
//       blocks by that count and lays the tile's <= 2048 footprints out as JAGGED DIAGONALS: the j-th footprint
//       of the block at sorted position p lives in slot jd[j] + p.  Output: 8 x uint16 slots per pixel, the
//       sorted block list and the diagonal starts (25 B per pixel).  Invalid taps (cuh:180) get kSlotSkip,
//       footprints outside the region kSlotGlobal.
//
//   bwd_state_local_kernel (one launch per iteration; same grid)
//       stage 1  exactly bwd_state_kernel's: gs from the scatter plane (now ONE padded plane: one coalesced
//                load, and the same thread clears the cell), G, gy, confidence gradient.
//       stage 2  each thread writes its taps' weighted corner quads (cuh:71-79) to shared memory at their slots:
//                plain 16-byte stores (a slot has exactly one writer).
//       stage 3a thread p sums the quads of sorted block p: consecutive threads read consecutive slots of one
//                diagonal (conflict-free) and a warp's blocks have (nearly) the same length (no divergence).
//       stage 3b one thread per aligned 2x2 cell group adds the four block phases that overlap it and the
//                centre tap: the tile's complete contribution to its region.
//       stage 4  ONE TMA tensor reduction (cp.reduce.async.bulk.tensor .add, UTMAREDG) adds the 48x24 region to
//                the global plane: 144 full sectors per CTA instead of 2048 half-used ones.
//       Footprints outside the region (|offset| > 7 px) keep four guarded scalar REDs.
//
// Per pixel and iteration the write port sees ~0.6 sectors instead of 8.  The scatter planes shrink from three
// sets of four phase planes (3 x 55 MB at KITTI B=8) to two padded planes (2 x 14.5 MB) that stay in L2.
// ncu of the first version (one thread per block over block-sorted slots, profiles/r02_local_v1_*): 180 us --
// issue 62 %, LSU 74 %: the per-block loops diverge (a warp runs to its longest row; 364 instructions per warp
// for 8 useful loads per thread) and rows 128 B apart collide in the banks (917 wavefronts for 2048 quads).
// The jagged-diagonal layout removes both.
#pragma once
#include "kernels_tiled.cuh"

namespace nlspn {

template <int K> struct LocalGeo;
template <> struct LocalGeo<3> {
    static constexpr int TW = 32, TH = 8, R = 8;
    static constexpr int RW = TW + 2 * R, RH = TH + 2 * R;   // region: 48 x 24 cells
    static constexpr int BW = RW / 2, BH = RH / 2;           // 24 x 12 blocks per phase
    static constexpr int NBLK = 4 * BH * BW;                 // 1152 blocks
    static constexpr int NCELL = RW * RH;                    // 1152 cells
    static constexpr int NT = TW * TH;                       // 256 threads = pixels per tile
    static constexpr int NTAP = 8;
    static constexpr int NQUAD = NT * NTAP;                  // 2048 slots
    static constexpr int MAXLEN = 31;                        // footprints per block kept in the tile (more -> global REDs)
    // per-tile table (uint16): sorted block list [NBLK] (block | length << 11), diagonal starts [32], nnz + padding [8]
    static constexpr int TAB_JD = NBLK, TAB_NNZ = NBLK + 32;
    static constexpr int ROWS = NBLK + 32 + 8;               // 1192 uint16 = 149 x 16 bytes
};

template <int K> struct LocalSmem {
    using L = LocalGeo<K>;
    static constexpr size_t bytes = sizeof(float4) * (L::NQUAD + L::NBLK) + sizeof(float) * L::NT +
                                    sizeof(unsigned short) * L::ROWS;
};

constexpr unsigned kSlotSkip = 0xFFFEu;      // invalid tap (cuh:180) or pixel outside the image
constexpr unsigned kSlotGlobal = 0xFFFFu;    // footprint outside the tile's region: guarded global REDs

// padded scatter plane of one image: R cells on every side, row pitch a multiple of 16 bytes
struct PadGeo {
    int PH, PW;
    long plane;
};
__host__ __device__ inline PadGeo pad_geo(int H, int W, int R)
{
    PadGeo g;
    g.PH = H + 2 * R;
    g.PW = ((W + 2 * R + 3) / 4) * 4;
    g.plane = (long)g.PH * g.PW;
    return g;
}

// ======================================================================================
// Schedule build.  grid = (ceil(W/32), ceil(H/8), nb), block = (32, 8).
//   slots     [nb][tiles][256] uint4 = 8 x uint16 per pixel (tap n in bits 16*(n&1) of word n>>1)
//   table     [nb][tiles][ROWS] uint16: sorted blocks (block | length << 11), diagonal starts, nnz
// ======================================================================================
template <int K>
__global__ void __launch_bounds__(LocalGeo<K>::NT)
sched_build_kernel(const float *__restrict__ offset, int H, int W, uint4 *__restrict__ slots,
                   unsigned short *__restrict__ table_g)
{
    using G = Geo<K>;
    using L = LocalGeo<K>;
    __shared__ unsigned int cnt[L::NBLK];            // footprints per block
    __shared__ unsigned short pos_s[L::NBLK];        // sorted position of a block
    __shared__ __align__(16) unsigned short tab[L::ROWS];
    __shared__ unsigned int hist[32], fill[32], start_s[32], jd_s[32];
    const int tid = threadIdx.y * L::TW + threadIdx.x;
    for (int i = tid; i < L::NBLK; i += L::NT) cnt[i] = 0u;
    for (int i = tid; i < L::ROWS; i += L::NT) tab[i] = 0;
    if (tid < 32) hist[tid] = fill[tid] = 0u;
    __syncthreads();
    const int P = H * W;
    const int x0 = blockIdx.x * L::TW, y0 = blockIdx.y * L::TH;
    const long b = blockIdx.z;
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    const float *ob = offset + b * 2 * G::KK * P + r;
    int blk[L::NTAP];
    unsigned rank[L::NTAP];
#pragma unroll
    for (int n = 0; n < L::NTAP; ++n) {
        const int t = n < G::REF ? n : n + 1;
        blk[n] = -1;
        rank[n] = 0u;
        if (!inside) continue;
        const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
        const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
        if (!tap_valid(h_im, w_im, H, W)) continue;
        float hf, wf;
        int hl, wl;
        floor_small(h_im, hf, hl);
        floor_small(w_im, wf, wl);
        const int Y = hl - (y0 - L::R), X = wl - (x0 - L::R);   // region coordinates of the top-left corner
        if ((unsigned)Y <= (unsigned)(L::RH - 2) && (unsigned)X <= (unsigned)(L::RW - 2)) {
            blk[n] = (((Y & 1) * 2 + (X & 1)) * L::BH + (Y >> 1)) * L::BW + (X >> 1);
            rank[n] = atomicAdd(&cnt[blk[n]], 1u);
        } else {
            blk[n] = -2;
        }
    }
    __syncthreads();
    // histogram of block lengths (capped)
    for (int i = tid; i < L::NBLK; i += L::NT) {
        const unsigned len = cnt[i] < (unsigned)L::MAXLEN ? cnt[i] : (unsigned)L::MAXLEN;
        if (len) atomicAdd(&hist[len], 1u);
    }
    __syncthreads();
    if (tid < 32) {
        // start[l] = blocks longer than l = blocks on diagonal l;  jd[l] = first slot of diagonal l
        const unsigned hl_ = hist[tid];
        unsigned suf = hl_;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_down_sync(0xffffffffu, suf, o);
            if (tid + o < 32) suf += u;
        }
        const unsigned st = suf - hl_;
        unsigned inc = st;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += u;
        }
        start_s[tid] = st;
        jd_s[tid] = inc - st;
        tab[L::TAB_JD + tid] = (unsigned short)(inc - st);
        if (tid == 0) tab[L::TAB_NNZ] = (unsigned short)st;
    }
    __syncthreads();
    // sorted positions: blocks of length l occupy [start[l], start[l] + hist[l])
    for (int i = tid; i < L::NBLK; i += L::NT) {
        const unsigned len = cnt[i] < (unsigned)L::MAXLEN ? cnt[i] : (unsigned)L::MAXLEN;
        if (len) {
            const unsigned p = start_s[len] + atomicAdd(&fill[len], 1u);
            pos_s[i] = (unsigned short)p;
            tab[p] = (unsigned short)(i | (len << 11));
        }
    }
    __syncthreads();
    unsigned sl[L::NTAP];
#pragma unroll
    for (int n = 0; n < L::NTAP; ++n) {
        if (blk[n] >= 0)
            sl[n] = rank[n] < (unsigned)L::MAXLEN ? jd_s[rank[n]] + pos_s[blk[n]] : kSlotGlobal;
        else
            sl[n] = blk[n] == -2 ? kSlotGlobal : kSlotSkip;
    }
    const long tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    slots[tile * L::NT + tid] = make_uint4(sl[0] | (sl[1] << 16), sl[2] | (sl[3] << 16), sl[4] | (sl[5] << 16),
                                           sl[6] | (sl[7] << 16));
    uint4 *tg = reinterpret_cast<uint4 *>(table_g + tile * L::ROWS);
    if (tid < L::ROWS / 8) tg[tid] = reinterpret_cast<const uint4 *>(tab)[tid];
}

// TMA tensor reduction: global[box at (x, y, z)] += smem box (fp32 add), bulk-group completion
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, const void *smem_src, int x, int y, int z)
{
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(tma::smem_u32(smem_src)), "r"(x), "r"(y), "r"(z)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ======================================================================================
// One backward iteration of the state gradient, local-transpose form.
//   s_in   padded planes [nb][PH][PW] written by iteration t+1 (nullptr at t = T): read, then cleared
//   s_out  padded planes of this iteration (plane_map describes [planes][PH][PW]; z_out = plane of image 0)
// grid = (ceil(W/32), ceil(H/8), nb), block = (32, 8).
// ======================================================================================
template <int K, bool STREAM>
__global__ void __launch_bounds__(LocalGeo<K>::NT, 3)
bwd_state_local_kernel(const __grid_constant__ CUtensorMap plane_map, int z_out,
                       const float *__restrict__ offset, const float *__restrict__ aff,
                       const float *__restrict__ conf, const float *__restrict__ dep,
                       const float *__restrict__ x_t, const float *__restrict__ g_ext,
                       float *__restrict__ s_in, float *__restrict__ s_out,
                       const uint4 *__restrict__ slots, const unsigned short *__restrict__ table_g,
                       float *__restrict__ gy_out, float *__restrict__ g_conf_acc, unsigned flags, int H, int W)
{
    using G = Geo<K>;
    using L = LocalGeo<K>;
    // dynamic shared memory (LocalSmem<K>::bytes > the 48 KB static limit)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *quads = reinterpret_cast<float4 *>(smem_raw);                     // stage 2/3a; reused as the flush tile
    float4 *blocksum = quads + L::NQUAD;
    float *centre = reinterpret_cast<float *>(blocksum + L::NBLK);
    unsigned short *tab = reinterpret_cast<unsigned short *>(centre + L::NT);   // sorted blocks, diagonal starts, nnz
    const int P = H * W;
    const int x0 = blockIdx.x * L::TW, y0 = blockIdx.y * L::TH;
    const long b = blockIdx.z;
    const int tid = threadIdx.y * L::TW + threadIdx.x;
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    const long q = b * P + r;
    const PadGeo pg = pad_geo(H, W, L::R);
    const long tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

    // ---- stage 0: everything that does not depend on the previous backward iteration
    tma::grid_launch_dependents();
    if (tid < L::ROWS / 8)
        reinterpret_cast<uint4 *>(tab)[tid] = __ldg(reinterpret_cast<const uint4 *>(table_g + tile * L::ROWS) + tid);
    for (int k = tid; k < L::NBLK; k += L::NT) blocksum[k] = make_float4(0.f, 0.f, 0.f, 0.f);   // empty blocks stay zero
    const uint4 sl4 = __ldg(slots + tile * L::NT + tid);
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
    float oh[G::KK], ow[G::KK], av[G::KK];
    float gext = 0.f, cf = 1.f, xt = 1.f, dp = 0.f;
    const bool need_x = (s_in && conf) || (flags & kAlwaysClip);
    if (inside) {
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            av[t] = ld_geo<STREAM>(ab + (long)t * P);
            oh[t] = ow[t] = 0.f;
            if (t != G::REF) {
                oh[t] = ld_geo<STREAM>(ob + (long)(2 * t) * P);
                ow[t] = ld_geo<STREAM>(ob + (long)(2 * t + 1) * P);
            }
        }
        gext = g_ext ? __ldg(g_ext + q) : 0.f;
        cf = conf ? __ldg(conf + q) : 1.f;
        xt = need_x ? __ldg(x_t + q) : 1.f;
        dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    } else {
#pragma unroll
        for (int t = 0; t < G::KK; ++t) av[t] = oh[t] = ow[t] = 0.f;
    }
    tma::grid_dependency_wait();   // s_in / s_out / g_conf_acc belong to the previous launch

    // ---- stage 1: gs, G, gy (nlspnmodel.py:351,357,361 backwards; identical to bwd_state_kernel)
    float gy = 0.f;
    const long pc = b * pg.plane + (long)(h + L::R) * pg.PW + (w + L::R);
    float gs = 0.f, gca = 0.f;
    if (inside && s_in) {
        gs = __ldcg(s_in + pc);
        if (conf) gca = g_conf_acc[q];
    }
    // keep every use of the stage-0 loads BELOW the two loads above: otherwise the scheduler parks a use of `dp`
    // in front of griddepcontrol.wait and the CTA pays two serialised memory round trips (ncu: 1167 + 806 stall
    // samples on those two instructions)
    asm volatile("" : "+f"(dp), "+f"(cf), "+f"(xt), "+f"(gext));
    if (inside) {
        float Gx = gext;
        if (flags & kBlendPre) {   // upstream order: the blend sits on the gather's INPUT
            if (s_in) Gx += (flags & kPreserve) ? (1.0f - (dp > 0.f ? 1.f : 0.f)) * gs : gs;
        } else {
            if (s_in) Gx += conf ? cf * gs : gs;
            if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
            if (flags & kPreserve) Gx = (1.0f - (dp > 0.f ? 1.f : 0.f)) * Gx;
        }
        gy = Gx;
        if (s_in) {
            s_in[pc] = 0.f;                      // this plane is the target of iteration t-1
            if (conf) g_conf_acc[q] = gca + xt * gs;
        }
        gy_out[q] = gy;
    }

    // ---- stage 2: weighted corner quads to their slots
    const unsigned slw[4] = {sl4.x, sl4.y, sl4.z, sl4.w};
#pragma unroll
    for (int n = 0; n < L::NTAP; ++n) {
        const int t = n < G::REF ? n : n + 1;
        const unsigned slot = (slw[n >> 1] >> (16 * (n & 1))) & 0xFFFFu;
        if (slot == kSlotSkip) continue;
        const float top = gy * av[t];
        const float h_im = (float)(h - G::PAD + t / K) + oh[t];
        const float w_im = (float)(w - G::PAD + t % K) + ow[t];
        const float hf = floor_small_f(h_im), wf = floor_small_f(w_im);
        // mdmcn_get_gradient_weight, cuh:71-79 (expressions kept literal; (float)(hl+1) == hf + 1 exactly)
        const float h1 = hf + 1.f, w1 = wf + 1.f;
        const float th = h1 - h_im, bh = (h_im + 1.f) - h1;
        const float lw_ = w1 - w_im, rw = (w_im + 1.f) - w1;
        const float4 qd = make_float4(th * lw_ * top, th * rw * top, bh * lw_ * top, bh * rw * top);
        if (slot != kSlotGlobal) {
            quads[slot] = qd;
        } else if (gy != 0.f) {
            // far footprint: guarded scalar REDs (cuh:229-252 with the guards of :37-48) into the padded plane
            const int hl = (int)hf, wl = (int)wf;
            float *sp = s_out + b * pg.plane + (long)(hl + L::R) * pg.PW + (wl + L::R);
            const bool topv = hl >= 0, botv = hl + 1 <= H - 1, lefv = wl >= 0, rigv = wl + 1 <= W - 1;
            if (topv && lefv) atomicAdd(sp, qd.x);
            if (topv && rigv) atomicAdd(sp + 1, qd.y);
            if (botv && lefv) atomicAdd(sp + pg.PW, qd.z);
            if (botv && rigv) atomicAdd(sp + pg.PW + 1, qd.w);
        }
    }
    centre[tid] = gy * av[G::REF];     // the centre tap has a structurally zero offset: lands on the pixel itself
    __syncthreads();

    // ---- stage 3a: block sums over the jagged diagonals: thread p owns sorted block p, its j-th quad is slot jd[j] + p
    {
        const int nnz = tab[L::TAB_NNZ];
        for (int p = tid; p < nnz; p += L::NT) {
            const unsigned e = tab[p];
            const int len = (int)(e >> 11);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < len; ++j) {
                const float4 v = quads[tab[L::TAB_JD + j] + p];
                acc.x += v.x;
                acc.y += v.y;
                acc.z += v.z;
                acc.w += v.w;
            }
            blocksum[e & 0x7FFu] = acc;
        }
    }
    __syncthreads();

    // ---- stage 3b: every aligned 2x2 cell group adds the four block phases that overlap it
    //   phase (sy, sx) block (by, bx) covers region rows 2*by + sy, +1 and columns 2*bx + sx, +1;
    //   components x y z w = (row0,col0) (row0,col1) (row1,col0) (row1,col1).
    float *flush = reinterpret_cast<float *>(quads);           // all slots were consumed in stage 3a
    constexpr int PB = L::BH * L::BW;
    for (int k = tid; k < PB; k += L::NT) {
        const int cy = k / L::BW, cx = k - cy * L::BW;
        const bool up = cy > 0, lf = cx > 0;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 p00 = blocksum[k];
        const float4 p01l = lf ? blocksum[PB + k - 1] : z4, p01r = blocksum[PB + k];
        const float4 p10u = up ? blocksum[2 * PB + k - L::BW] : z4, p10d = blocksum[2 * PB + k];
        const float4 p11ul = (up && lf) ? blocksum[3 * PB + k - L::BW - 1] : z4;
        const float4 p11ur = up ? blocksum[3 * PB + k - L::BW] : z4;
        const float4 p11dl = lf ? blocksum[3 * PB + k - 1] : z4;
        const float4 p11dr = blocksum[3 * PB + k];
        float a00 = ((p00.x + p01l.y) + p10u.z) + p11ul.w;
        float a01 = ((p00.y + p01r.x) + p10u.w) + p11ur.z;
        float a10 = ((p00.z + p01l.w) + p10d.x) + p11dl.y;
        float a11 = ((p00.w + p01r.z) + p10d.y) + p11dr.x;
        const int ty = 2 * cy - L::R, tx = 2 * cx - L::R;       // tile coordinates of the group's first cell
        if ((unsigned)ty < (unsigned)L::TH && (unsigned)tx < (unsigned)L::TW) {
            const float2 c0 = *reinterpret_cast<const float2 *>(centre + ty * L::TW + tx);
            const float2 c1 = *reinterpret_cast<const float2 *>(centre + (ty + 1) * L::TW + tx);
            a00 += c0.x;
            a01 += c0.y;
            a10 += c1.x;
            a11 += c1.y;
        }
        *reinterpret_cast<float2 *>(flush + (2 * cy) * L::RW + 2 * cx) = make_float2(a00, a01);
        *reinterpret_cast<float2 *>(flush + (2 * cy + 1) * L::RW + 2 * cx) = make_float2(a10, a11);
    }
    tma::fence_proxy_async_smem();
    __syncthreads();

    // ---- stage 4: one tensor reduction adds the region to the global plane.  Region origin = pixel
    // (y0 - R, x0 - R) = padded coordinates (y0, x0): never negative, x0 a multiple of 32 (16-byte aligned);
    // the part of the box beyond the plane's high edges is clipped by the TMA unit.
    if (tid == 0) {
        tma_reduce_add_3d(&plane_map, flush, x0, y0, z_out + (int)b);
        tma_bulk_wait_all();      // the reduction is performed before this CTA (and with it the grid) completes
    }
}

} // namespace nlspn
