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
//       footprints outside the region kSlotGlobal.  The same kernel re-packs everything the T iterations read
//       per pixel and that does not change between them -- the slots, the 8 sampling coordinates, the 9 affinities, the
//       fixed confidence and the input-preserving multiplier -- into ONE record of 8 x 16 bytes per pixel, laid
//       out [tile][vector][thread]: an iteration then issues 8 coalesced 128-bit loads from one base address
//       instead of 30 scalar loads with a 64-bit address chain each (ncu, first version: 248 of 982 instructions
//       per warp and iteration were stage-0 address arithmetic).
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

template <int K, int TH_> struct LocalGeo;
template <int TH_> struct LocalGeo<3, TH_> {
    static_assert(TH_ == 8 || TH_ == 16, "tile height 8 or 16");
    static constexpr int TW = 32, TH = TH_, R = 8;
    static constexpr int RW = TW + 2 * R, RH = TH + 2 * R;   // region: 48 x 24 (48 x 32) cells
    static constexpr int BW = RW / 2, BH = RH / 2;           // 24 x 12 (24 x 16) blocks per phase
    static constexpr int NBLK = 4 * BH * BW;                 // 1152 (1536) blocks: 4.5 (3) per pixel
    static constexpr int NCELL = RW * RH;                    // 1152 (1536) cells
    static constexpr int NT = TW * TH;                       // 256 (512) threads = pixels per tile
    static constexpr int NTAP = 8;
    static constexpr int NQUAD = NT * NTAP;                  // 2048 (4096) slots
    static constexpr int MAXLEN = 31;                        // footprints per block kept in the tile (more -> global REDs)
    // per-tile table (uint16): sorted block list [NBLK] (block | length << 11), diagonal starts [32], nnz + padding [8]
    static constexpr int TAB_JD = NBLK, TAB_NNZ = NBLK + 32;
    static constexpr int ROWS = NBLK + 32 + 8;               // 1192 (1576) uint16 = 149 (197) x 16 bytes
    static_assert(NBLK <= 2048 && ROWS % 8 == 0 && ROWS / 8 <= NT, "table entry = block (11 bits) | length (5 bits)");
    // packed per-pixel record, [tile][GEOV][NT] x 16 bytes: 0 slots | 1-4 sampling coordinates of taps (2i, 2i+1)
    // as (h, w, h, w) | 5-6 affinities of the 8 neighbours | 7 (centre affinity, confidence, preserve factor, 0)
    static constexpr int GEOV = 8;
};

template <int K, int TH> struct LocalSmem {
    using L = LocalGeo<K, TH>;
    // quads (32 KB; after stage 3a the same bytes hold the block sums and the flush tile) + centre taps + table
    static constexpr size_t bytes = sizeof(float4) * L::NQUAD + sizeof(float) * L::NT + sizeof(unsigned short) * L::ROWS;
    static_assert(sizeof(float4) * L::NBLK + sizeof(float) * L::NCELL <= sizeof(float4) * L::NQUAD, "block sums + flush tile alias the quads");
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

// ---- host side: tiles per image and workspace bytes of the local form
template <int TH>
inline long local_tiles(int H, int W)
{
    using L = LocalGeo<3, TH>;
    return (long)((W + L::TW - 1) / L::TW) * ((H + L::TH - 1) / L::TH);
}

template <int TH>
inline size_t ws_bytes_local_th(int B, int H, int W, int K, int T)
{
    using L = LocalGeo<3, TH>;
    const size_t BP = (size_t)B * H * W;
    const PadGeo pg = pad_geo(H, W, L::R);
    const size_t tiles = (size_t)B * (size_t)local_tiles<TH>(H, W);
    // two padded planes + confidence-gradient accumulator + gy of every iteration + raw affinity-gradient
    // accumulator + the schedule (packed per-pixel records: GEOV x 16 B per pixel slot of every tile; block lists:
    // ROWS uint16 per tile)
    return sizeof(float) * (2 * (size_t)B * pg.plane + BP + (size_t)T * BP + (size_t)K * K * BP) + 64 +
           tiles * (L::GEOV * L::NT * sizeof(uint4) + L::ROWS * sizeof(unsigned short)) + 64;
}

// ======================================================================================
// Schedule build + geometry re-pack.  grid = (ceil(W/32), ceil(H/8), nb), block = (32, 8).
//   geo       [nb][tiles][GEOV][256] x 16 B: the packed per-pixel record (LocalGeo::GEOV)
//   table     [nb][tiles][ROWS] uint16: sorted blocks (block | length << 11), diagonal starts, nnz
// ======================================================================================
template <int K, int TH, int MINB>
__global__ void __launch_bounds__(LocalGeo<K, TH>::NT, MINB)
sched_build_kernel(const float *__restrict__ offset, const float *__restrict__ aff, const float *__restrict__ conf,
                   const float *__restrict__ dep, unsigned flags, int H, int W, uint4 *__restrict__ geo,
                   unsigned short *__restrict__ table_g)
{
    using G = Geo<K>;
    using L = LocalGeo<K, TH>;
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
    float oh[L::NTAP], ow[L::NTAP];
    // every global load of the kernel is issued here, ahead of the shared-memory atomics and barriers
    float av[G::KK];
    float cf = 1.f, pm = 1.f;
    const float *ab = aff + b * G::KK * P + r;
#pragma unroll
    for (int n = 0; n < L::NTAP; ++n) {
        const int t = n < G::REF ? n : n + 1;
        oh[n] = inside ? __ldg(ob + (long)(2 * t) * P) : 0.f;
        ow[n] = inside ? __ldg(ob + (long)(2 * t + 1) * P) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < G::KK; ++t) av[t] = inside ? __ldg(ab + (long)t * P) : 0.f;
    if (inside) {
        if (conf) cf = __ldg(conf + b * P + r);
        if (flags & kPreserve) pm = 1.0f - (__ldg(dep + b * P + r) > 0.f ? 1.f : 0.f);   // nlspnmodel.py:357 backwards
    }
#pragma unroll
    for (int n = 0; n < L::NTAP; ++n) {
        const int t = n < G::REF ? n : n + 1;
        blk[n] = -1;
        rank[n] = 0u;
        if (!inside) continue;
        // the record keeps the sampling COORDINATE (integer part added first, one floating add: cuh:157-158), which
        // is what every iteration would rebuild from the offset
        const float h_im = (float)(h - G::PAD + t / K) + oh[n];
        const float w_im = (float)(w - G::PAD + t % K) + ow[n];
        oh[n] = h_im;
        ow[n] = w_im;
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
        // empty blocks (len 0) follow the nnz non-empty ones: start_s[0] = nnz, so the list covers all NBLK blocks
        const unsigned p = start_s[len] + atomicAdd(&fill[len], 1u);
        pos_s[i] = (unsigned short)p;
        tab[p] = (unsigned short)(i | (len << 11));
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
    // the packed record (pixels of the tile that lie outside the image: all taps skipped, zeros)
    uint4 *gp = geo + tile * (L::GEOV * L::NT) + tid;
    gp[0] = make_uint4(sl[0] | (sl[1] << 16), sl[2] | (sl[3] << 16), sl[4] | (sl[5] << 16), sl[6] | (sl[7] << 16));
    float4 *gf = reinterpret_cast<float4 *>(gp);
#pragma unroll
    for (int i = 0; i < 4; ++i) gf[(1 + i) * L::NT] = make_float4(oh[2 * i], ow[2 * i], oh[2 * i + 1], ow[2 * i + 1]);
    // neighbour n = tap n (n < REF) or n + 1
    gf[5 * L::NT] = make_float4(av[0], av[1], av[2], av[3]);
    gf[6 * L::NT] = make_float4(av[5], av[6], av[7], av[8]);
    gf[7 * L::NT] = make_float4(av[G::REF], cf, pm, 0.f);
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
// 16-byte store to a shared-memory address computed once per thread (the generic-pointer form makes the compiler
// rebuild the CTA's shared window base in front of every store inside a divergent region: 3 instructions per tap)
__device__ __forceinline__ void sts128(unsigned addr, const float4 &v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// the CTA may exit once the bulk reduction has READ its shared-memory source; the global side completes before the
// grid does (same rule as every TMA-store epilogue)
__device__ __forceinline__ void tma_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// pull `bytes` (multiple of 16, 16-byte aligned) into L2 without a destination
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ======================================================================================
// One backward iteration of the state gradient, local-transpose form.
//   geo    packed per-pixel records of sched_build_kernel
//   s_in   padded planes [nb][PH][PW] written by iteration t+1 (nullptr at t = T): read, then cleared
//   s_out  padded planes of this iteration (plane_map describes [planes][PH][PW]; z_out = plane of image 0)
// grid = (ceil(W/32), ceil(H/8), nb), block = (32, 8).
// ======================================================================================
template <int K, int TH, bool STREAM, int MINB>
__global__ void __launch_bounds__(LocalGeo<K, TH>::NT, MINB)
bwd_state_local_kernel(const __grid_constant__ CUtensorMap plane_map, int z_out, const uint4 *__restrict__ geo,
                       int has_conf, const float *__restrict__ x_t, const float *__restrict__ g_ext,
                       float *__restrict__ s_in, float *__restrict__ s_out,
                       const unsigned short *__restrict__ table_g,
                       float *__restrict__ gy_out, float *__restrict__ g_conf_acc, unsigned flags, int H, int W,
                       int pf_dist)
{
    static_assert(K == 3, "the packed record holds 8 neighbours");
    using G = Geo<K>;
    using L = LocalGeo<K, TH>;
    // dynamic shared memory (LocalSmem<K>::bytes > the 48 KB static limit)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *quads = reinterpret_cast<float4 *>(smem_raw);                     // stage 2/3a
    float4 *blocksum = quads;                                                 // stage 3a' / 3b: aliases the consumed quads
    float *flush = reinterpret_cast<float *>(quads + L::NBLK);                // stage 3b / 4: behind the block sums
    float *centre = reinterpret_cast<float *>(quads + L::NQUAD);
    unsigned short *tab = reinterpret_cast<unsigned short *>(centre + L::NT);   // sorted blocks, diagonal starts, nnz
    const int P = H * W;
    const int x0 = blockIdx.x * L::TW, y0 = blockIdx.y * L::TH;
    const int b = blockIdx.z;
    const int tid = threadIdx.y * L::TW + threadIdx.x;
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    const PadGeo pg = pad_geo(H, W, L::R);
    const long tile = ((long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    // per-image bases are CTA-uniform 64-bit values; everything per thread is a 32-bit index on top of them
    const float *xb = x_t + (long)b * P;
    const float *geb = g_ext ? g_ext + (long)b * P : nullptr;
    float *gyb = gy_out + (long)b * P;
    float *gcb = g_conf_acc + (long)b * P;
    float *sib = s_in ? s_in + (long)b * pg.plane : nullptr;
    float *sob = s_out + (long)b * pg.plane;

    // ---- stage 0: everything that does not depend on the previous backward iteration
    tma::grid_launch_dependents();
    if (tid == 0 && pf_dist > 0) {
        // the CTA that will run `pf_dist` tiles from now (about one wave of resident CTAs) finds its record in L2:
        // CTA lifetime here is ~8x its issue time, most of it the DRAM round trip of stage 0
        const long nt = tile + pf_dist;
        if (nt < (long)gridDim.x * gridDim.y * gridDim.z) {
            bulk_prefetch_l2(geo + nt * (L::GEOV * L::NT), L::GEOV * L::NT * 16);
            bulk_prefetch_l2(table_g + nt * L::ROWS, L::ROWS * 2);
        }
    }
    // the tile's block list goes global -> shared without a register round trip (first needed in stage 3a)
    if (tid < L::ROWS / 8)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tma::smem_u32(tab) + tid * 16u),
                     "l"(table_g + tile * L::ROWS + tid * 8) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    const uint4 *gp = geo + tile * (L::GEOV * L::NT) + tid;
    const float4 *gf = reinterpret_cast<const float4 *>(gp);
    const uint4 sl4 = STREAM ? __ldcs(gp) : __ldg(gp);
    // SPLIT (6 CTAs per SM, 40 registers): only the first four neighbours' vectors are loaded here; the other three
    // vectors are pulled into L2 and loaded when the first four have been consumed (32 -> 20 registers of record)
    constexpr bool SPLIT = MINB >= 6;
    float4 o4[SPLIT ? 2 : 4], a4[SPLIT ? 1 : 2], m4;
    o4[0] = STREAM ? __ldcs(gf + 1 * L::NT) : __ldg(gf + 1 * L::NT);
    o4[1] = STREAM ? __ldcs(gf + 2 * L::NT) : __ldg(gf + 2 * L::NT);
    a4[0] = STREAM ? __ldcs(gf + 5 * L::NT) : __ldg(gf + 5 * L::NT);
    if constexpr (!SPLIT) {
        o4[2] = STREAM ? __ldcs(gf + 3 * L::NT) : __ldg(gf + 3 * L::NT);
        o4[3] = STREAM ? __ldcs(gf + 4 * L::NT) : __ldg(gf + 4 * L::NT);
        a4[1] = STREAM ? __ldcs(gf + 6 * L::NT) : __ldg(gf + 6 * L::NT);
    } else if ((tid & 7) == 0) {          // one lane per 128-byte line
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gf + 3 * L::NT));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gf + 4 * L::NT));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gf + 6 * L::NT));
    }
    m4 = STREAM ? __ldcs(gf + 7 * L::NT) : __ldg(gf + 7 * L::NT);   // (centre affinity, confidence, preserve factor, -)
    const bool need_x = (s_in && has_conf) || (flags & kAlwaysClip);
    float gext = 0.f, xt = 1.f;
    if (inside) {
        gext = geb ? __ldg(geb + r) : 0.f;
        xt = need_x ? __ldg(xb + r) : 1.f;
    }
    tma::grid_dependency_wait();   // s_in / s_out / g_conf_acc belong to the previous launch

    // ---- stage 1: gs, G, gy (nlspnmodel.py:351,357,361 backwards; identical to bwd_state_kernel)
    float gy = 0.f;
    const int pc = (h + L::R) * pg.PW + (w + L::R);
    float gs = 0.f, gca = 0.f;
    if (inside && sib) {
        gs = __ldcg(sib + pc);
        if (has_conf) gca = gcb[r];
    }
    // keep every use of the stage-0 loads BELOW the two loads above: otherwise the scheduler parks a use of them
    // in front of griddepcontrol.wait and the CTA pays two serialised memory round trips (ncu: 1167 + 806 stall
    // samples on those two instructions)
    asm volatile("" : "+f"(m4.y), "+f"(m4.z), "+f"(xt), "+f"(gext));
    const float cf = m4.y, pm = m4.z;
    if (inside) {
        float Gx = gext;
        if (flags & kBlendPre) {   // upstream order: the blend sits on the gather's INPUT
            if (sib) Gx += (flags & kPreserve) ? pm * gs : gs;
        } else {
            if (sib) Gx += has_conf ? cf * gs : gs;
            if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
            if (flags & kPreserve) Gx = pm * Gx;
        }
        gy = Gx;
        if (sib) {
            sib[pc] = 0.f;                      // this plane is the target of iteration t-1
            if (has_conf) gcb[r] = gca + xt * gs;
        }
        gyb[r] = gy;
    }

    // ---- stage 2: weighted corner quads to their slots
    const unsigned slw[4] = {sl4.x, sl4.y, sl4.z, sl4.w};
    unsigned quads_s = tma::smem_u32(quads);
    asm volatile("" : "+r"(quads_s));            // opaque: computed once, not rebuilt in front of every store
    // four neighbours n0 .. n0+3: coordinates (h, w, h, w) of two taps per vector, one vector of affinities
    auto four_taps = [&](int n0, const float4 &oa, const float4 &ob_, const float4 &aa) {
        const float ohv[4] = {oa.x, oa.z, ob_.x, ob_.z};
        const float owv[4] = {oa.y, oa.w, ob_.y, ob_.w};
        const float avv[4] = {aa.x, aa.y, aa.z, aa.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + i;
            const unsigned slot = (slw[n >> 1] >> (16 * (n & 1))) & 0xFFFFu;
            if (slot == kSlotSkip) continue;
            const float top = gy * avv[i];
            const float h_im = ohv[i], w_im = owv[i];
            const float hf = floor_small_f(h_im), wf = floor_small_f(w_im);
            // mdmcn_get_gradient_weight, cuh:71-79 (expressions kept literal; (float)(hl+1) == hf + 1 exactly)
            const float h1 = hf + 1.f, w1 = wf + 1.f;
            const float th = h1 - h_im, bh = (h_im + 1.f) - h1;
            const float lw_ = w1 - w_im, rw = (w_im + 1.f) - w1;
            const float4 qd = make_float4(th * lw_ * top, th * rw * top, bh * lw_ * top, bh * rw * top);
            if (slot != kSlotGlobal) {
                sts128(quads_s + slot * 16u, qd);
            } else if (gy != 0.f) {
                // far footprint: guarded scalar REDs (cuh:229-252 with the guards of :37-48) into the padded plane
                const int hl = (int)hf, wl = (int)wf;
                float *sp = sob + ((hl + L::R) * pg.PW + (wl + L::R));
                const bool topv = hl >= 0, botv = hl + 1 <= H - 1, lefv = wl >= 0, rigv = wl + 1 <= W - 1;
                if (topv && lefv) atomicAdd(sp, qd.x);
                if (topv && rigv) atomicAdd(sp + 1, qd.y);
                if (botv && lefv) atomicAdd(sp + pg.PW, qd.z);
                if (botv && rigv) atomicAdd(sp + pg.PW + 1, qd.w);
            }
        }
    };
    four_taps(0, o4[0], o4[1], a4[0]);
    if constexpr (SPLIT) {
        const float4 oc = __ldg(gf + 3 * L::NT), od = __ldg(gf + 4 * L::NT), ab_ = __ldg(gf + 6 * L::NT);
        four_taps(4, oc, od, ab_);
    } else {
        four_taps(4, o4[2], o4[3], a4[1]);
    }
    centre[tid] = gy * m4.x;     // the centre tap has a structurally zero offset: lands on the pixel itself
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- stage 3a: block sums over the jagged diagonals.  Thread `tid` owns the sorted blocks p_k = tid + 256 k;
    // the j-th quad of block p is slot jd[j] + p.  Blocks are sorted by descending length, so len(p_0) >= len(p_1)
    // >= ...: at diagonal j the thread's active blocks are a prefix of k, and lanes of a warp agree on it except at
    // one boundary.  One read of jd[j] serves all of them.  The sums stay in registers until every quad has been
    // read; then they overwrite the quads (block sums and flush tile alias them: 36 KB per CTA instead of 55).
    {
        constexpr int NK = (L::NBLK + L::NT - 1) / L::NT;     // 5 (the last one half full)
        unsigned e[NK];
        int len[NK];
        float4 acc[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int p = tid + k * L::NT;
            e[k] = p < L::NBLK ? tab[p] : 0u;               // the list covers all NBLK blocks, empty ones last
            len[k] = (int)(e[k] >> 11);
            acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int j = 0; j < len[0]; ++j) {
            const float4 *qj = quads + tab[L::TAB_JD + j] + tid;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                if (j >= len[k]) break;
                const float4 v = qj[k * L::NT];
                acc[k].x += v.x;
                acc[k].y += v.y;
                acc[k].z += v.z;
                acc[k].w += v.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NK; ++k)
            if (tid + k * L::NT < L::NBLK) blocksum[e[k] & 0x7FFu] = acc[k];
    }
    __syncthreads();

    // ---- stage 3b: every aligned 2x2 cell group adds the four block phases that overlap it
    //   phase (sy, sx) block (by, bx) covers region rows 2*by + sy, +1 and columns 2*bx + sx, +1;
    //   components x y z w = (row0,col0) (row0,col1) (row1,col0) (row1,col1).
    //   Every lane reads the six blocks at its own column with full 16-byte loads (consecutive lanes, consecutive
    //   blocks: no bank conflicts) and takes the three left-hand blocks' components from lane - 1 by shuffle; only
    //   lane 0 of a warp fetches them itself.  (The first version read single components at a 16-byte lane stride:
    //   4-way conflicts, 56 wavefronts per warp for 20 ideal.)
    constexpr int PB = L::BH * L::BW;
    static_assert(PB % 32 == 0, "whole warps of cell groups");
    for (int k = tid; k < PB; k += L::NT) {          // warp-uniform: PB is a multiple of 32
        const int cy = k / L::BW, cx = k - cy * L::BW;
        const bool up = cy > 0, lf = cx > 0;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 p00 = blocksum[k];
        const float4 p01r = blocksum[PB + k];
        const float4 p10d = blocksum[2 * PB + k];
        const float4 p11dr = blocksum[3 * PB + k];
        const float4 p10u = up ? blocksum[2 * PB + k - L::BW] : z4;
        const float4 p11ur = up ? blocksum[3 * PB + k - L::BW] : z4;
        // left neighbours = lane - 1's right-hand blocks (same row whenever cx > 0)
        float p01l_y = __shfl_up_sync(0xffffffffu, p01r.y, 1), p01l_w = __shfl_up_sync(0xffffffffu, p01r.w, 1);
        float p11dl_y = __shfl_up_sync(0xffffffffu, p11dr.y, 1), p11ul_w = __shfl_up_sync(0xffffffffu, p11ur.w, 1);
        if ((tid & 31) == 0 && lf) {
            const float4 a = blocksum[PB + k - 1];
            p01l_y = a.y;
            p01l_w = a.w;
            p11dl_y = blocksum[3 * PB + k - 1].y;
            p11ul_w = up ? blocksum[3 * PB + k - L::BW - 1].w : 0.f;
        }
        if (!lf) p01l_y = p01l_w = p11dl_y = p11ul_w = 0.f;
        float a00 = ((p00.x + p01l_y) + p10u.z) + p11ul_w;
        float a01 = ((p00.y + p01r.x) + p10u.w) + p11ur.z;
        float a10 = ((p00.z + p01l_w) + p10d.x) + p11dl_y;
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
        tma_reduce_add_3d(&plane_map, flush, x0, y0, z_out + b);
        tma_bulk_wait_read();     // shared memory must outlive the read; the grid completes after the global side
    }
}

} // namespace nlspn
