// kernels_gather.cuh -- pass A of the backward in GATHER form (sm_100a).
//
// The RED scatter of kernels_v2.cuh is bound by the SM->L2 write port: every scattered 16-byte
// vector RED occupies a 32-byte sector slot, 8 per pixel and iteration at K = 3 (ncu: DRAM 40 %,
// l1tex2xbar the busiest unit).  But the sampling geometry is the same for all T iterations, so
// the transposed operator can be TABULATED ONCE per backward call:
//
//   table_build_kernel   one pass over (pixel, tap): the footprint's 2x2 block (same four
//                        phase-shifted blocked planes as the RED path, so consumers are unchanged)
//                        gets an entry {source pixel (row << 16 | col), aff, th, lw} (16 B; th, lw = the top-row /
//                        left-column bilinear weights) in the next free
//                        slot of that block (one ATOMG per tap, once per step).  Blocks are grouped
//                        by 32; inside a group the layout is slot-major, so the readers are coalesced.
//                        Taps that overflow a block's CAP slots are flagged per pixel and keep
//                        the scatter form (scalar REDs into a plain "overflow" plane).
//   table_compact_kernel (long runs) sorts each group's 32 columns by entry count, so that the live entries
//                        of every row form a prefix: the readers then stream exactly the table's bytes.
//   bwd_gy_kernel        per iteration, per pixel: gs = 4 phase cells + overflow plane, then
//                        G, gy exactly as bwd_state_kernel's stages 1-3 (nlspnmodel.py:351,357,361);
//                        stores gy, accumulates the confidence gradient.
//   bwd_gather_kernel    per iteration, per block: sums its entries' gy[src] * aff * corner
//                        weights (cuh:71-79, factored as (th, 1-th) x (lw, 1-lw))
//                        and STORES the block -- no atomics, no clearing, one 16-byte coalesced
//                        store per thread; reads 16 B per entry, coalesced.
//
// Per pixel and iteration this moves ~16 B x N entries + ~60 B through HBM on the READ side instead of N
// RED sectors through the write port.  Measured on B200 (KITTI, B = 8), uncompacted / compacted table:
//   K = 3, T = 18: gy 0.72 + gather 2.50 / 2.25 + table 0.61 / 0.85 = 3.83 ms  vs  3.17 ms RED scatter -> RED stays
//   K = 5, T = 36: gy 1.45 + gather 10.3 / 9.2  + table 1.97 / 2.60 = 13.7 / 13.3 ms  vs  16.1 ms     -> gather form
// ncu, K = 5, compacted: 1.48 GB of DRAM reads in 260 us = 5.7 TB/s, 87 % of the measured HBM peak.
// (RED cost grows with 32-byte sectors per tap, the table with 16 bytes per tap).  Also measured: a
// TMA-delivered gy box for the gy[src] gathers (no gain: the kernel is bound by table bytes), 8 instead
// of 4 entries in flight per thread (slower), fetching the first slots in parallel with the counter
// (slower: wasted bytes), slot-major planes instead of groups of 32 blocks (K = 5: 5 % slower).
#pragma once
#include "kernels_v2.cuh"

namespace nlspn {

// entries per block: mean occupancy is N (N taps per pixel, one block per pixel and phase);
// CAP covers mean + ~3 sigma of a Poisson(N) load, the rest takes the overflow path
__host__ __device__ constexpr int gather_cap(int K) { return K == 3 ? 16 : (K == 5 ? 40 : 72); }

#ifndef NLSPN_GATHER_U
#define NLSPN_GATHER_U 4
#endif

// Table layout: blocks are grouped by 32 (one warp of readers); a group's CAP x 32 entries are contiguous,
// slot-major inside the group: entry(slot, blk) = ((blk / 32) * CAP + slot) * 32 + blk % 32.  A warp's loads are
// 512-byte rows of one 16*CAP*32-byte region (sequential in DRAM, one TLB entry) instead of CAP rows that lie
// one whole slot plane (megabytes) apart.  The per-image table holds ceil(NB / 32) groups.
__host__ __device__ inline long table_groups(long NB) { return (NB + 31) / 32; }
template <int CAP>
__device__ __forceinline__ long table_index(long blk, int slot)
{
    return ((blk >> 5) * CAP + slot) * 32 + (blk & 31);
}

template <int K> struct OvfMask { using type = unsigned long long; };
template <> struct OvfMask<3> { using type = unsigned char; };
template <> struct OvfMask<5> { using type = unsigned int; };

// ---------------------------------------------------------------------------------------------
// Table build: grid = (ceil(P/256), nb).  count [nb][NB] must be zero on entry.
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kBlock)
table_build_kernel(const float *__restrict__ offset, const float *__restrict__ aff, int H, int W,
                   int *__restrict__ count, float4 *__restrict__ entries,
                   typename OvfMask<K>::type *__restrict__ ovf)
{
    using G = Geo<K>;
    using M = typename OvfMask<K>::type;
    constexpr int CAP = gather_cap(K);
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const ScatterGeo sg = scatter_geo(H, W);
    const long NB = sg.plane;                 // blocks per image = 4 phases * Hb * Wb
    const int blocks_per_phase = sg.Hb * sg.Wb;
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
    int *cnt = count + b * table_groups(NB) * 32;          // counters are padded to whole groups
    float4 *ent = entries + b * table_groups(NB) * CAP * 32;
    M mask = 0;
    // eight neighbours at a time: their 24 loads are issued before the first counter atomic (an atomic with a
    // return value otherwise serialises load -> atomic -> store per tap)
#pragma unroll
    for (int n0 = 0; n0 < G::N; n0 += 8) {
        float oh[8], ow[8], av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = n0 + i, t = n < G::REF ? n : n + 1;
            oh[i] = __ldg(ob + (long)(2 * t) * P);
            ow[i] = __ldg(ob + (long)(2 * t + 1) * P);
            av[i] = __ldg(ab + (long)t * P);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = n0 + i, t = n < G::REF ? n : n + 1;
            const float h_im = (float)(h - G::PAD + t / K) + oh[i];
            const float w_im = (float)(w - G::PAD + t % K) + ow[i];
            if (!tap_valid(h_im, w_im, H, W)) continue;
            float hf, wf;
            int hl, wl;
            floor_small(h_im, hf, hl);
            floor_small(w_im, wf, wl);
            const int Y = hl + 1, X = wl + 1;         // padded coordinates of the top-left corner
            const int sy = Y & 1, sx = X & 1;
            const int blk = (sy * 2 + sx) * blocks_per_phase + ((Y + sy) >> 1) * sg.Wb + ((X + sx) >> 1);
            const int slot = atomicAdd(cnt + blk, 1);
            // top-row and left-column weights of cuh:71-79, (hl+1) - h and (wl+1) - w, formed literally; the
            // readers take the bottom / right weights as 1 - th and 1 - lw (equal up to one rounding of 2^-25)
            if (slot < CAP)
                ent[table_index<CAP>(blk, slot)] = make_float4(__int_as_float((h << 16) | w), av[i],
                                                               (hf + 1.f) - h_im, (wf + 1.f) - w_im);
            else
                mask |= (M)1 << n;
        }
    }
    ovf[b * P + r] = mask;
}

// ---------------------------------------------------------------------------------------------
// Table compaction: one warp per group of 32 blocks, once per call, after the build.
// Past the mean occupancy a table row is only partly used, and which of its 32 columns are used is
// random, so the readers touch most 64-byte DRAM chunks of a row for half the bytes.  Here the group's
// columns are permuted by descending entry count: every row's live entries become a PREFIX of the row.
// count[] is rewritten in column order and owner[] maps a column back to its block (-1: padding).
//   grid = (ceil(groups / 8), nb), block = 256 (8 warps).
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kBlock)
table_compact_kernel(int H, int W, int *__restrict__ count, float4 *__restrict__ entries, int *__restrict__ owner)
{
    constexpr int CAP = gather_cap(K);
    const ScatterGeo sg = scatter_geo(H, W);
    const long NB = sg.plane;
    const long groups = table_groups(NB);
    const long g = (long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    if (g >= groups) return;
    const int lane = threadIdx.x & 31;
    const long b = blockIdx.y;
    const long blk = g * 32 + lane;
    int *cnt = count + b * groups * 32 + g * 32;
    int n = cnt[lane];
    n = n < CAP ? n : CAP;
    if (blk >= NB) n = 0;
    // rank of this column in descending count order (stable)
    int r = 0, nmax = 0;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const int v = __shfl_sync(0xffffffffu, n, o);
        r += (v > n || (v == n && o < lane)) ? 1 : 0;
        nmax = v > nmax ? v : nmax;
    }
    float4 *ent = entries + (b * groups + g) * CAP * 32;
    for (int j = 0; j < nmax; ++j) {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < n) e = ent[j * 32 + lane];
        __syncwarp();
        if (j < n) ent[j * 32 + r] = e;
        __syncwarp();
    }
    __syncwarp();
    cnt[r] = n;
    owner[b * groups * 32 + g * 32 + r] = blk < NB ? (int)blk : -1;
}

// ---------------------------------------------------------------------------------------------
// gy of one backward iteration.  grid = (ceil(P/256), nb).
//   s_prev   blocked planes written by bwd_gather_kernel of iteration t+1 (nullptr at t = T)
//   f_in     overflow plane of iteration t+1 (read, then cleared); f_out: this iteration's
// The centre tap (zero offset, nlspnmodel.py:256) lands on the pixel itself: iteration t+1 left it
// out of the table and its contribution gy_{t+1}[p] * aff_ref[p] is added here (gy_next).
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kBlock)
bwd_gy_kernel(const float *__restrict__ offset, const float *__restrict__ aff,
              const float *__restrict__ conf, const float *__restrict__ dep,
              const float *__restrict__ x_t, const float *__restrict__ g_ext,
              const float *__restrict__ s_prev, float *__restrict__ f_in, float *__restrict__ f_out,
              const float *__restrict__ gy_next, const typename OvfMask<K>::type *__restrict__ ovf,
              float *__restrict__ gy_out, float *__restrict__ g_conf_acc, unsigned flags, int H, int W)
{
    using G = Geo<K>;
    using M = typename OvfMask<K>::type;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;
    const ScatterGeo sg = scatter_geo(H, W);
    tma::grid_launch_dependents();
    // loads that do not depend on the previous launch
    const float gext = g_ext ? __ldg(g_ext + q) : 0.f;
    const float cf = conf ? __ldg(conf + q) : 1.f;
    const bool need_x = (s_prev && conf) || (flags & kAlwaysClip);
    const float xt = need_x ? __ldg(x_t + q) : 1.f;
    const float dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    const M mask = ovf[q];
    const float a_ref = s_prev ? __ldg(aff + (b * G::KK + G::REF) * P + r) : 0.f;
    tma::grid_dependency_wait();
    float gs = 0.f;
    float gca = 0.f;
    if (s_prev) {
        const float *sl = s_prev + b * sg.image;
        float cv[4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) cv[ph] = __ldcg(sl + scatter_cell(sg, ph >> 1, ph & 1, h + 1, w + 1));
        const float fo = __ldcg(f_in + q);
        const float gyn = __ldcg(gy_next + q);
        if (conf) gca = g_conf_acc[q];
        gs = ((((cv[0] + cv[1]) + cv[2]) + cv[3]) + fo) + gyn * a_ref;
        if (fo != 0.f) f_in[q] = 0.f;      // the overflow plane is sparse: clear only what was written
    }
    float Gx = gext;
    if (flags & kBlendPre) {   // upstream order: the blend sits on the gather's INPUT
        if (s_prev) Gx += (flags & kPreserve) ? (1.0f - (dp > 0.f ? 1.f : 0.f)) * gs : gs;
    } else {
        if (s_prev) Gx += conf ? cf * gs : gs;
        if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
        if (flags & kPreserve) Gx = (1.0f - (dp > 0.f ? 1.f : 0.f)) * Gx;
    }
    const float gy = Gx;
    if (s_prev && conf) g_conf_acc[q] = gca + xt * gs;
    gy_out[q] = gy;
    if (mask == 0 || gy == 0.f) return;
    // taps that found their block full keep the scatter form (cuh:229-252 with the guards of :37-48)
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
    float *fo = f_out + b * P;
#pragma unroll 1
    for (int n = 0; n < G::N; ++n) {
        if (!((mask >> n) & 1)) continue;
        const int t = n < G::REF ? n : n + 1;
        const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
        const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
        const float top = gy * __ldg(ab + (long)t * P);
        float hf, wf;
        int hl, wl;
        floor_small(h_im, hf, hl);
        floor_small(w_im, wf, wl);
        const float h1 = hf + 1.f, w1 = wf + 1.f;
        const float th = h1 - h_im, bh = (h_im + 1.f) - h1;
        const float lw_ = w1 - w_im, rw = (w_im + 1.f) - w1;
        const bool topv = hl >= 0, botv = hl + 1 <= H - 1, lefv = wl >= 0, rigv = wl + 1 <= W - 1;
        float *sp = fo + (long)hl * W + wl;
        if (topv && lefv) atomicAdd(sp, th * lw_ * top);
        if (topv && rigv) atomicAdd(sp + 1, th * rw * top);
        if (botv && lefv) atomicAdd(sp + W, bh * lw_ * top);
        if (botv && rigv) atomicAdd(sp + W + 1, bh * rw * top);
    }
}

// ---------------------------------------------------------------------------------------------
// Gather-reduce of one backward iteration.  One thread per block; grid = (ceil(NB/256), nb).
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kBlock)
bwd_gather_kernel(const int *__restrict__ count, const int *__restrict__ owner,
                  const float4 *__restrict__ entries, const float *__restrict__ gy, int H, int W,
                  float4 *__restrict__ s_out)
{
    constexpr int CAP = gather_cap(K);
    constexpr int U = NLSPN_GATHER_U;         // entries in flight per thread and trip
    const int P = H * W;
    const ScatterGeo sg = scatter_geo(H, W);
    const long NB = sg.plane;
    const long NBpad = table_groups(NB) * 32;      // thread i serves COLUMN i of the compacted table
    const long i = (long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= NBpad) return;
    const long b = blockIdx.y;
    tma::grid_launch_dependents();
    // owner == nullptr: the table was not compacted (short runs), column i is block i
    const int blk = owner ? __ldg(owner + b * NBpad + i) : (i < NB ? (int)i : -1);
    if (blk < 0) return;
    // The table does not depend on the previous launch: its first U slots are fetched before waiting for gy.
    // (Fetching them unconditionally, in parallel with the counter, was measured and is slower: the kernel is
    // bound by table bytes, and slots past the counter are wasted traffic.)
    const float4 *ent = entries + b * table_groups(NB) * CAP * 32 + table_index<CAP>(i, 0);
    int n = __ldg(count + b * NBpad + i);
    n = n < CAP ? n : CAP;
    float4 e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = u < n ? __ldg(ent + u * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
    tma::grid_dependency_wait();
    const float *gb = gy + b * P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < n; j0 += U) {
        float4 nx[U];
        const bool more = j0 + U < n;
#pragma unroll
        for (int u = 0; u < U; ++u)
            nx[u] = (more && j0 + U + u < n) ? __ldg(ent + (j0 + U + u) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        float g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pk = __float_as_int(e[u].x);
            g[u] = j0 + u < n ? __ldg(gb + (pk >> 16) * W + (pk & 0xffff)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (j0 + u >= n) continue;        // stale slot
            // mdmcn_get_gradient_weight, cuh:71-79, factored: (th, 1-th) x (lw, 1-lw) x gy*aff
            const float top = g[u] * e[u].y;
            const float th = e[u].z, lw_ = e[u].w;
            const float tt = th * top, bt = (1.f - th) * top, rw = 1.f - lw_;
            acc.x += tt * lw_;
            acc.y += tt * rw;
            acc.z += bt * lw_;
            acc.w += bt * rw;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = nx[u];
    }
    s_out[b * NB + blk] = acc;
}

} // namespace nlspn
