// kernels_tiled.cuh -- TMA-tiled gather kernels (sm_100a).
//
// ncu on the direct-gather kernels (profiles/r01_v1_*): the 4*K^2 bilinear corner loads are
// uncoalesced LDGs averaging ~13 sectors per request and L1 is the busiest unit (67 %), while
// DRAM idles at 48 %.  Here each CTA owns a TW x TH pixel tile; the plane it gathers from
// arrives in shared memory as ONE halo'd (TW+2R) x (TH+2R) box moved by the Tensor Memory
// Accelerator (cp.async.bulk.tensor.3d, completion on an mbarrier).  TMA zero-fills whatever
// lies outside the image, which is precisely the sampler's zero padding, so the gather loop has
// no image-border guards; the only test left is "does the 2x2 footprint lie inside my box".
// Taps that leave the box (|offset| > R-1 px) fall back to guarded global loads, so results are
// identical for arbitrary offsets -- the box only makes the common case cheap.
//
//   iter_fwd_tiled_kernel   one propagation iteration (same epilogue as iter_fwd_kernel)
//   bwd_param_tiled_kernel  pass B of the backward: loops t = T..1 with a 2-deep TMA pipeline
//                           (plane t-1 is in flight while plane t is consumed)
#pragma once
#include "kernels_v2.cuh"
#include "tma.cuh"

namespace nlspn {

constexpr int kTileW = 32;        // one warp per tile row: coalesced per-pixel planes

// box halo R in pixels: covers the tap displacement (K-1)/2 plus |offset| <= 7 (3.5 sigma of the
// sigma = 2 px offsets of SURVEY 8d).  R is a multiple of 4 so that a box row is a multiple of
// 32 bytes: a 208-byte row (R = 10) made cp.async.bulk.tensor trap with error 715 on B200,
// 192- and 224-byte rows are fine.
__host__ __device__ constexpr int halo_for(int K) { return K == 3 ? 8 : 12; }

template <int K, int TH>
struct TileGeo {
    static constexpr int R = halo_for(K);
    static constexpr int BoxW = kTileW + 2 * R;
    static constexpr int BoxH = TH + 2 * R;
    static constexpr int BoxFloats = BoxH * BoxW;
    static constexpr int BoxBytes = BoxFloats * 4;
};

// four corner values of a sample from the shared-memory box, falling back to global memory
// when the footprint is not inside the box.  box(0,0) is pixel (y0 - R, x0 - R).
template <int K, int TH>
__device__ __forceinline__ Quad box_quad(const float *__restrict__ box, int y0, int x0,
                                         const float *__restrict__ im, int H, int W, float h_im, float w_im)
{
    constexpr int kHalo = TileGeo<K, TH>::R;
    constexpr int kBoxW = TileGeo<K, TH>::BoxW;
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int hl = (int)hf, wl = (int)wf;
    const int ty = hl - (y0 - kHalo), tx = wl - (x0 - kHalo);
    if ((unsigned)ty < (unsigned)(TileGeo<K, TH>::BoxH - 1) && (unsigned)tx < (unsigned)(kBoxW - 1)) {
        Quad q;
        q.hl = hl;
        q.wl = wl;
        q.lh = h_im - hf;
        q.lw = w_im - wf;
        const float *p = box + ty * kBoxW + tx;
        q.v1 = p[0];
        q.v2 = p[1];
        q.v3 = p[kBoxW];
        q.v4 = p[kBoxW + 1];
        return q;
    }
    return load_quad(im, H, W, h_im, w_im);
}

// ======================================================================================
// Forward iteration, tiled.  grid = (tiles_x, tiles_y, nb), block = (32, TH).
// plane_z = index of the source plane of image 0 of this launch inside the tensor map.
// ======================================================================================
template <int K, int TH, bool STREAM>
__global__ void __launch_bounds__(kTileW * TH)
iter_fwd_tiled_kernel(const __grid_constant__ CUtensorMap src_map, int plane_z,
                      const float *__restrict__ src_prev, const float *__restrict__ offset,
                      const float *__restrict__ aff, const float *__restrict__ conf,
                      const float *__restrict__ dep, unsigned flags, int H, int W,
                      float *__restrict__ out, float *__restrict__ src_next)
{
    using G = Geo<K>;
    using TG = TileGeo<K, TH>;
    constexpr int kHalo = TG::R;
    constexpr int kBoxW = TG::BoxW;
    __shared__ __align__(128) float box[TG::BoxFloats];
    __shared__ __align__(8) uint64_t bar;
    const int P = H * W;
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * TH;
    const long b = blockIdx.z;
    const int tid = threadIdx.y * kTileW + threadIdx.x;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    // PDL: let the next iteration's CTAs become resident as soon as ours have started; they
    // only prefetch their (iteration-independent) geometry before their own grid_dependency_wait
    tma::grid_launch_dependents();
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = h * W + w;
    // streamed geometry: independent of the previous iteration, issued before waiting for it
    float oh[G::KK], ow[G::KK], av[G::KK];
    float dp = 0.f, cf = 1.f;
    if (inside) {
        const float *ob = offset + b * 2 * G::KK * P + r;
        const float *ab = aff + b * G::KK * P + r;
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            av[t] = ld_geo<STREAM>(ab + (long)t * P);
            oh[t] = ow[t] = 0.f;
            if (t != G::REF) {
                oh[t] = ld_geo<STREAM>(ob + (long)(2 * t) * P);
                ow[t] = ld_geo<STREAM>(ob + (long)(2 * t + 1) * P);
            }
        }
        if (flags & kPreserve) dp = __ldg(dep + b * P + r);
        if (conf && src_next) cf = __ldg(conf + b * P + r);
    }
    // everything below depends on the previous iteration's output
    tma::grid_dependency_wait();
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&bar, TG::BoxBytes);
        tma::load_3d(box, &src_map, &bar, x0 - kHalo, y0 - kHalo, plane_z + (int)b);
    }
    tma::mbar_wait(&bar, 0);
    if (!inside) return;

    // Footprint (rows hl, hl+1; cols wl, wl+1) lies inside the box  <=>  ylo <= h_im < yhi and
    // xlo <= w_im < xhi (NaN -> false).  Inside the box the validity test of cuh:180 is implied by
    // TMA's zero fill: an out-of-image footprint reads zeros, and a coordinate of exactly -1 puts
    // all its weight (lh = 0) on the zero row.
    const float ylo = (float)(y0 - kHalo), yhi = (float)(y0 - kHalo + TG::BoxH - 1);
    const float xlo = (float)(x0 - kHalo), xhi = (float)(x0 - kHalo + kBoxW - 1);
    float hi[G::KK], wi[G::KK];
    bool all_in = true;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        if (t == G::REF) continue;
        hi[t] = (float)(h - G::PAD + t / K) + oh[t];
        wi[t] = (float)(w - G::PAD + t % K) + ow[t];
        all_in = all_in && hi[t] >= ylo && hi[t] < yhi && wi[t] >= xlo && wi[t] < xhi;
    }
    const float *im = src_prev + b * P;
    float acc = 0.f;
    if (__all_sync(__activemask(), all_in)) {
        // branch-free: every tap of every lane of this warp reads the box
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            float v;
            if (t == G::REF) {
                v = box[(threadIdx.y + kHalo) * kBoxW + threadIdx.x + kHalo];
            } else {
                const float hf = floorf(hi[t]), wf = floorf(wi[t]);
                const float lh = hi[t] - hf, lw = wi[t] - wf;
                const float hh = 1.f - lh, hw = 1.f - lw;
                const float *p = box + ((int)hf - (y0 - kHalo)) * kBoxW + ((int)wf - (x0 - kHalo));
                v = (hh * hw) * p[0] + (hh * lw) * p[1] + (lh * hw) * p[kBoxW] + (lh * lw) * p[kBoxW + 1];
            }
            acc += v * av[t];
        }
    } else {
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            float v;
            if (t == G::REF) {
                v = box[(threadIdx.y + kHalo) * kBoxW + threadIdx.x + kHalo];
            } else {
                v = 0.f;
                if (tap_valid(hi[t], wi[t], H, W))
                    v = quad_value(box_quad<K, TH>(box, y0, x0, im, H, W, hi[t], wi[t]));
            }
            acc += v * av[t];
        }
    }
    const long q = b * P + r;
    if (flags & kBlendPre) {   // upstream order: raw gather to the list, blended state to the next gather
        out[q] = acc;
        if (src_next) src_next[q] = (flags & kPreserve) ? blend_fix(acc, dp) : acc;
        return;
    }
    if (flags & kPreserve) acc = blend_fix(acc, dp);
    if (flags & kAlwaysClip) acc = clip_keep_sign(acc);
    out[q] = acc;
    if (src_next) src_next[q] = conf ? acc * cf : acc;
}

// ======================================================================================
// Pass B of the backward, tiled.  grid = (tiles_x, tiles_y, nb * NCH), block = (32, TH).
// For iteration t the gather source is plane  z_t  of one of two tensor maps:
//   with confidence:     src_map,  z = (t-1)*Bsrc + b0 + b
//   without confidence:  t == 1 -> src_map, z = b0 + b ; t >= 2 -> list_map, z = (t-2)*Bsrc + b0 + b
//
// The sampling geometry of a tap does not depend on t, so everything derived from the offsets
// (box index of the footprint, fractional weights, validity) is computed ONCE and kept in
// registers; the t loop is 4 LDS + ~20 FP instructions per tap.  Each tap is classified once:
//   FAST  footprint inside the box and the tap passes the validity test of cuh:180
//   SLOW  valid but outside the box  -> guarded global loads every iteration
//   SKIP  invalid (cuh:308-311: contributes nothing)
// A warp whose lanes are all-FAST runs a branch-free loop.
// ======================================================================================
#ifndef NLSPN_PARAM_WARPS_PER_SM
#define NLSPN_PARAM_WARPS_PER_SM 24   // 24 warps/SM -> <= 85 registers per thread (was 16 warps / 128 registers)
#endif
// FACTORED = true (default): the bilinear value and its two coordinate derivatives are evaluated in factored form
// (11 FP instructions per tap and iteration instead of ~23; the round-1 kernel issued at 65 % next to an LSU at 85 %):
//     d1 = v2 - v1, d2 = v4 - v3, top = v1 + lw d1, bot = v3 + lw d2,
//     d/dh = bot - top, value = top + lh (bot - top), d/dw = d1 + lh (d2 - d1)
// (algebraically the expressions of cuh:50-52,101-122).  Measured and dropped: a second copy of every box shifted
// by one column, so that the row pair (p[0], p[1]) of any footprint is one aligned LDS.64 -- cp.async.bulk.tensor
// traps (error 715) on a start coordinate that is not 16-byte aligned, for loads as for stores.
template <int K, int C, int TH, int NS = 2, bool FACTORED = true>
__global__ void __launch_bounds__(kTileW * TH, (TH >= NLSPN_PARAM_WARPS_PER_SM ? 1 : NLSPN_PARAM_WARPS_PER_SM / TH))
bwd_param_tiled_kernel(const __grid_constant__ CUtensorMap src_map,
                       const __grid_constant__ CUtensorMap list_map, int Bsrc, int b0,
                       const float *__restrict__ offset, const float *__restrict__ aff,
                       const float *__restrict__ src, const float *__restrict__ list_feat,
                       const float *__restrict__ gy_all, int has_conf, int H, int W, int T, long BP,
                       long GP, float *__restrict__ g_guidance, float *__restrict__ g_aff_acc)
{
    using G = Geo<K>;
    using TG = TileGeo<K, TH>;
    constexpr int kHalo = TG::R;
    constexpr int kBoxW = TG::BoxW;
    constexpr int NCH = (G::KK + C - 1) / C;
    // NS-deep TMA pipeline: the boxes of iterations t-1 .. t-(NS-1) are in flight while plane t is consumed
    __shared__ __align__(128) float box[NS][TG::BoxFloats];
    __shared__ __align__(8) uint64_t bar[NS];
    const int P = H * W;
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * TH;
    const int b = NCH == 1 ? (int)blockIdx.z : (int)blockIdx.z / NCH;
    const int k0 = NCH == 1 ? 0 : ((int)blockIdx.z % NCH) * C;
    const int tid = threadIdx.y * kTileW + threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) tma::mbar_init(&bar[i], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int t, int buf) {   // thread 0 only
        tma::mbar_arrive_expect_tx(&bar[buf], TG::BoxBytes);
        const bool use_src = has_conf || t == 1;
        const int z = use_src ? (has_conf ? (t - 1) * Bsrc : 0) + b0 + b : (t - 2) * Bsrc + b0 + b;
        tma::load_3d(box[buf], use_src ? &src_map : &list_map, &bar[buf], x0 - kHalo, y0 - kHalo, z);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS - 1; ++i)
            if (T - i >= 1) issue(T - i, i);
    }

    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    const long q = (long)b * P + r;

    // ---- per-tap geometry, computed once
    // Register diet (24 warps per SM instead of 16): the coordinate weights (hl+1)-h and (wl+1)-w of
    // cuh:101-122 are taken as 1-lh and 1-lw (identical up to one rounding of 2^-25), and the affinity
    // factor of the offset gradients is applied once at the end (acc_h, acc_w sum gy*dh, gy*dw).
    int idx[C];                         // box index of the footprint's top-left corner (FAST taps)
    float lh[C], lw[C];
    float acc_h[C], acc_w[C], acc_a[C];
    unsigned slow_bits = 0u, skip_bits = 0u;
    constexpr int kCenterIdx = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int k = k0 + c;
        idx[c] = kCenterIdx;
        lh[c] = lw[c] = 0.f;
        acc_h[c] = acc_w[c] = acc_a[c] = 0.f;
        if (!inside || k >= G::KK) {
            skip_bits |= 1u << c;
            continue;
        }
        if (k == G::REF) {
            idx[c] = (threadIdx.y + kHalo) * kBoxW + threadIdx.x + kHalo;
            continue;
        }
        const float o_h = __ldg(offset + ((long)b * 2 * G::KK + 2 * k) * P + r);
        const float o_w = __ldg(offset + ((long)b * 2 * G::KK + 2 * k + 1) * P + r);
        const float h_im = (float)(h - G::PAD + k / K) + o_h;
        const float w_im = (float)(w - G::PAD + k % K) + o_w;
        if (!tap_valid(h_im, w_im, H, W)) {
            skip_bits |= 1u << c;
            continue;
        }
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int hl = (int)hf, wl = (int)wf;
        lh[c] = h_im - hf;                       // == h_im - (float)hl
        lw[c] = w_im - wf;
        const int ty = hl - (y0 - kHalo), tx = wl - (x0 - kHalo);
        if ((unsigned)ty < (unsigned)(TG::BoxH - 1) && (unsigned)tx < (unsigned)(kBoxW - 1)) {
            idx[c] = ty * kBoxW + tx;
        } else {
            slow_bits |= 1u << c;
            idx[c] = hl * W + wl;                // image index of the corner (may be negative)
        }
    }
    const bool warp_fast = __all_sync(0xffffffffu, (slow_bits | skip_bits) == 0u);

    float gy_n1 = inside ? __ldg(gy_all + (long)(T - 1) * GP + q) : 0.f;
    float gy_n2 = (inside && T > 1) ? __ldg(gy_all + (long)(T - 2) * GP + q) : 0.f;
    uint32_t phase_bits = 0u;   // bit i = parity the next wait on bar[i] expects
    int cur = NS - 1;
    for (int t = T; t >= 1; --t) {
        const int freed = cur;                    // the box consumed in the previous trip
        cur = cur + 1 == NS ? 0 : cur + 1;
        // everyone is done reading box[freed]: refill it with the plane NS-1 iterations ahead
        __syncthreads();
        if (tid == 0 && t - (NS - 1) >= 1) issue(t - (NS - 1), freed);
        const float gy = gy_n1;
        gy_n1 = gy_n2;
        if (t > 2) gy_n2 = inside ? __ldg(gy_all + (long)(t - 3) * GP + q) : 0.f;
        tma::mbar_wait(&bar[cur], (phase_bits >> cur) & 1u);
        phase_bits ^= 1u << cur;
        const float *bx = box[cur];
        if (warp_fast) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int k = k0 + c;
                const float *p = bx + idx[c];
                if (k == G::REF) {
                    acc_a[c] += gy * p[0];
                    continue;
                }
                if (FACTORED) {
                    const float v1 = p[0], v2 = p[1], v3 = p[kBoxW], v4 = p[kBoxW + 1];
                    const float d1 = v2 - v1, d2 = v4 - v3;
                    const float top = fmaf(lw[c], d1, v1), bot = fmaf(lw[c], d2, v3);
                    const float dh = bot - top;
                    acc_a[c] = fmaf(gy, fmaf(lh[c], dh, top), acc_a[c]);             // cuh:314-315
                    acc_h[c] = fmaf(gy, dh, acc_h[c]);
                    acc_w[c] = fmaf(gy, fmaf(lh[c], d2 - d1, d1), acc_w[c]);
                    continue;
                }
                const float v1 = p[0], v2 = p[1], v3 = p[kBoxW], v4 = p[kBoxW + 1];
                const float hh = 1.f - lh[c], hw = 1.f - lw[c];
                const float bil = (hh * hw) * v1 + (hh * lw[c]) * v2 + (lh[c] * hw) * v3 + (lh[c] * lw[c]) * v4;
                acc_a[c] += gy * bil;                                            // cuh:314-315
                const float dh = -1.f * hw * v1 + -1.f * lw[c] * v2 + hw * v3 + lw[c] * v4;
                const float dw = -1.f * hh * v1 + hh * v2 + -1.f * lh[c] * v3 + lh[c] * v4;
                acc_h[c] += dh * gy;
                acc_w[c] += dw * gy;
            }
        } else if (gy != 0.f) {
            const float *im;
            if (has_conf) im = src + (long)(t - 1) * BP + (long)b * P;
            else im = (t == 1 ? src : list_feat + (long)(t - 2) * BP) + (long)b * P;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int k = k0 + c;
                if (skip_bits & (1u << c)) continue;
                if (k == G::REF) {
                    acc_a[c] += gy * bx[idx[c]];
                    continue;
                }
                float v1, v2, v3, v4;
                if (slow_bits & (1u << c)) {
                    // guarded global loads (cuh:37-48); idx = hl*W + wl
                    const int hl = (int)floorf((float)(h - G::PAD + k / K) +
                                               __ldg(offset + ((long)b * 2 * G::KK + 2 * k) * P + r));
                    const int wl = idx[c] - hl * W;
                    const bool tp = hl >= 0, bt = hl + 1 <= H - 1, lf = wl >= 0, rg = wl + 1 <= W - 1;
                    const float *pg = im + idx[c];
                    v1 = (tp && lf) ? __ldg(pg) : 0.f;
                    v2 = (tp && rg) ? __ldg(pg + 1) : 0.f;
                    v3 = (bt && lf) ? __ldg(pg + W) : 0.f;
                    v4 = (bt && rg) ? __ldg(pg + W + 1) : 0.f;
                } else {
                    const float *p = bx + idx[c];
                    v1 = p[0]; v2 = p[1]; v3 = p[kBoxW]; v4 = p[kBoxW + 1];
                }
                if (FACTORED) {
                    const float d1 = v2 - v1, d2 = v4 - v3;
                    const float top = fmaf(lw[c], d1, v1), bot = fmaf(lw[c], d2, v3);
                    const float dh = bot - top;
                    acc_a[c] = fmaf(gy, fmaf(lh[c], dh, top), acc_a[c]);
                    acc_h[c] = fmaf(gy, dh, acc_h[c]);
                    acc_w[c] = fmaf(gy, fmaf(lh[c], d2 - d1, d1), acc_w[c]);
                    continue;
                }
                const float hh = 1.f - lh[c], hw = 1.f - lw[c];
                const float bil = (hh * hw) * v1 + (hh * lw[c]) * v2 + (lh[c] * hw) * v3 + (lh[c] * lw[c]) * v4;
                acc_a[c] += gy * bil;
                const float dh = -1.f * hw * v1 + -1.f * lw[c] * v2 + hw * v3 + lw[c] * v4;
                const float dw = -1.f * hh * v1 + hh * v2 + -1.f * lh[c] * v3 + lh[c] * v4;
                acc_h[c] += dh * gy;
                acc_w[c] += dw * gy;
            }
        }
    }
    if (!inside) return;
    float *ggb = g_guidance + (long)b * 3 * G::N * P + r;
    float *gab = g_aff_acc + (long)b * G::KK * P + r;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int k = k0 + c;
        if (k >= G::KK) continue;
        gab[(long)k * P] = acc_a[c];
        if (k != G::REF) {
            const int n = k < G::REF ? k : k - 1;
            const float a = __ldg(aff + ((long)b * G::KK + k) * P + r);   // d/d offset carries the affinity
            ggb[(long)(2 * n) * P] = acc_h[c] * a;
            ggb[(long)(2 * n + 1) * P] = acc_w[c] * a;
        }
    }
}

// ======================================================================================
// Pass A with the streamed geometry delivered by TMA.
// ncu on bwd_state_kernel: the SM->L2 request port is the busiest unit (l1tex2xbar request cycles
// 70-80 %): besides the 9 REDs, every thread issues 25 coalesced geometry loads = 100 sector
// requests per warp.  Here ONE thread per CTA asks the Tensor Memory Accelerator for the CTA's
// whole geometry -- a {32 px, TH rows, 2K^2 channels} box of `offset` and a {32, TH, K^2} box of
// `aff` -- which travels as bulk requests and lands in shared memory without touching the
// register file; the threads then read it conflict-free (lane = pixel).  Everything else is
// bwd_state_kernel.  grid = (ceil(W/32), ceil(H/TH), nb), block = (32, TH).
// ======================================================================================
template <int K, int TH>
__global__ void __launch_bounds__(kTileW * TH)
bwd_state_tma_kernel(const __grid_constant__ CUtensorMap off_map, const __grid_constant__ CUtensorMap aff_map,
                     int b0, const float *__restrict__ conf, const float *__restrict__ dep,
                     const float *__restrict__ x_t, const float *__restrict__ g_ext, float *__restrict__ s_in,
                     float *__restrict__ s_out, float *__restrict__ gy_out, float *__restrict__ g_conf_acc,
                     unsigned flags, int H, int W, float *__restrict__ s_zero)
{
    using G = Geo<K>;
    __shared__ __align__(128) float s_off[2 * G::KK][TH][kTileW];
    __shared__ __align__(128) float s_aff[G::KK][TH][kTileW];
    __shared__ __align__(8) uint64_t bar;
    const int P = H * W;
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * TH;
    const int b = blockIdx.z;
    const int tid = threadIdx.y * kTileW + threadIdx.x;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {   // the geometry does not depend on the previous launch: fetch it right away
        tma::mbar_arrive_expect_tx(&bar, (uint32_t)(sizeof(s_off) + sizeof(s_aff)));
        tma::load_4d(s_off, &off_map, &bar, x0, y0, 0, b0 + b);
        tma::load_4d(s_aff, &aff_map, &bar, x0, y0, 0, b0 + b);
    }
    tma::grid_launch_dependents();
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    const long q = (long)b * P + r;
    const ScatterGeo sg = scatter_geo(H, W);

    float gext = 0.f, cf = 1.f, xt = 1.f, dp = 0.f;
    const bool need_x = (s_in && conf) || (flags & kAlwaysClip);
    if (inside) {
        gext = g_ext ? __ldg(g_ext + q) : 0.f;
        cf = conf ? __ldg(conf + q) : 1.f;
        xt = need_x ? __ldg(x_t + q) : 1.f;
        dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    }
    tma::grid_dependency_wait();   // s_in / s_out / g_conf_acc belong to the previous launch
    float *si = s_in ? s_in + (long)b * sg.image : nullptr;
    long cell[4];
    float cv[4] = {0.f, 0.f, 0.f, 0.f};
    float gca = 0.f;
    if (inside) {
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            cell[ph] = scatter_cell(sg, ph >> 1, ph & 1, h + 1, w + 1);
            if (si) cv[ph] = __ldcg(si + cell[ph]);
        }
        if (si && conf) gca = g_conf_acc[q];
    }
    if (s_zero)   // three-set rotation: clear the set iteration t+1 consumed, linearly (see zero_set_linear)
        zero_set_linear(s_zero + (long)b * sg.image, sg.image / 4,
                        ((long)blockIdx.y * gridDim.x + blockIdx.x) * (kTileW * TH) + tid,
                        (long)gridDim.x * gridDim.y * (kTileW * TH));
    tma::mbar_wait(&bar, 0);
    if (!inside) return;

    const float gs = ((cv[0] + cv[1]) + cv[2]) + cv[3];
    float Gx = gext;
    if (flags & kBlendPre) {   // upstream order: the blend sits on the gather's INPUT
        if (si) Gx += (flags & kPreserve) ? (1.0f - (dp > 0.f ? 1.f : 0.f)) * gs : gs;
    } else {
        if (si) Gx += conf ? cf * gs : gs;
        if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
        if (flags & kPreserve) Gx = (1.0f - (dp > 0.f ? 1.f : 0.f)) * Gx;
    }
    const float gy = Gx;
    if (si && !s_zero) {
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) si[cell[ph]] = 0.f;
    }
    if (si && conf) g_conf_acc[q] = gca + xt * gs;
    gy_out[q] = gy;
    if (gy == 0.f) return;

    float *so = s_out + (long)b * sg.image;
    const int ty = threadIdx.y, tx = threadIdx.x;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const float top = gy * s_aff[t][ty][tx];
        if (t == G::REF) {
            atomicAdd(so + cell[0], top);
            continue;
        }
        const float h_im = (float)(h - G::PAD + t / K) + s_off[2 * t][ty][tx];
        const float w_im = (float)(w - G::PAD + t % K) + s_off[2 * t + 1][ty][tx];
        if (!tap_valid(h_im, w_im, H, W)) continue;
        float hf, wf;
        int hl, wl;
        floor_small(h_im, hf, hl);
        floor_small(w_im, wf, wl);
        // mdmcn_get_gradient_weight, cuh:71-79 (expressions kept literal; (float)(hl+1) == hf + 1 exactly)
        const float h1 = hf + 1.f, w1 = wf + 1.f;
        const float th = h1 - h_im, bh = (h_im + 1.f) - h1;
        const float lw_ = w1 - w_im, rw = (w_im + 1.f) - w1;
        const int Y = hl + 1, X = wl + 1;
        const int sy = Y & 1, sx = X & 1;
        float4 *blk = reinterpret_cast<float4 *>(
            so + (long)(sy * 2 + sx) * sg.plane + ((long)((Y + sy) >> 1) * sg.Wb + ((X + sx) >> 1)) * 4);
        atomicAdd(blk, make_float4(th * lw_ * top, th * rw * top, bh * lw_ * top, bh * rw * top));
    }
}

} // namespace nlspn
