// kernels_persist.cuh -- persistent forward kernel for batches small enough to live on chip.
//
// BASELINE.json's north-star kernel (2): every thread owns ONE pixel, keeps that pixel's K^2-1
// offset pairs and K^2 affinities in registers for all T iterations, and only the state plane
// ping-pongs through L2, with a grid-wide barrier per iteration (cooperative launch).
// It applies for K = 3 when B*H*W <= co-resident threads (122 registers per thread -> 2 CTAs of
// 256 threads per SM -> 148 x 512 = 75,776 pixels on B200: one NYU 228x304 frame); larger shapes do not fit the register file (one KITTI frame's
// geometry is 43 MB against 37 MB of registers) and take the per-iteration TMA-tiled kernels.
// What it buys is LATENCY: T+1 launches of a sub-wave grid (13 us each, launch-bound) become one
// launch whose iterations cost one L2 round trip plus a grid barrier.
#pragma once
#include <cooperative_groups.h>

#include "kernels_v2.cuh"

namespace nlspn {

constexpr int kPersistBlock = 256;

// corner values with the guards of cuh:37-48, read through L2 (ld.global.cg): the plane was
// written by other SMs earlier in THIS launch, so the (incoherent) L1 must be bypassed
__device__ __forceinline__ Quad load_quad_cg(const float *im, int H, int W, float h_im, float w_im)
{
    Quad q;
    const float hf = floorf(h_im), wf = floorf(w_im);
    q.hl = (int)hf;
    q.wl = (int)wf;
    q.lh = h_im - hf;
    q.lw = w_im - wf;
    const bool top = q.hl >= 0, bot = q.hl + 1 <= H - 1;
    const bool lef = q.wl >= 0, rig = q.wl + 1 <= W - 1;
    const float *p = im + (long)q.hl * W + q.wl;
    q.v1 = (top && lef) ? __ldcg(p) : 0.f;
    q.v2 = (top && rig) ? __ldcg(p + 1) : 0.f;
    q.v3 = (bot && lef) ? __ldcg(p + W) : 0.f;
    q.v4 = (bot && rig) ? __ldcg(p + W + 1) : 0.f;
    return q;
}

template <int K>
__global__ void __launch_bounds__(kPersistBlock)
persist_fwd_kernel(const float *__restrict__ offset, const float *__restrict__ aff,
                   const float *__restrict__ conf, const float *__restrict__ dep, unsigned flags, int H,
                   int W, int B, int T, float *src, int S, float *list_feat)
{
    namespace cg = cooperative_groups;
    using G = Geo<K>;
    cg::grid_group grid = cg::this_grid();
    const int P = H * W;
    const long BP = (long)B * P;
    const long q = (long)blockIdx.x * kPersistBlock + threadIdx.x;   // one pixel per thread
    const bool active = q < BP;
    const int b = active ? (int)(q / P) : 0;
    const int r = active ? (int)(q - (long)b * P) : 0;
    const int h = r / W, w = r - h * W;

    // ---- the pixel's geometry: loaded once, register-resident for all T iterations.
    // Per tap: image index of the footprint's top-left corner, the two fractional weights and four
    // corner-valid bits (validity test of cuh:180 AND the per-corner guards of cuh:37-48), so the
    // iteration body is branch-free: 4 predicated L2 loads + 8 FP ops per tap, all 32 loads of a
    // pixel in flight at once.
    int idx[G::KK];
    float lh[G::KK], lw[G::KK], av[G::KK];
    unsigned cmask = 0u;   // bit 4k+c: corner c of tap k is read
    float dp = 0.f, cf = 1.f;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        idx[t] = 0;
        lh[t] = lw[t] = av[t] = 0.f;
    }
    if (active) {
        const float *ob = offset + (long)b * 2 * G::KK * P + r;
        const float *ab = aff + (long)b * G::KK * P + r;
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            av[t] = __ldg(ab + (long)t * P);
            if (t == G::REF) continue;
            const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
            const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
            if (!tap_valid(h_im, w_im, H, W)) continue;
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int hl = (int)hf, wl = (int)wf;
            lh[t] = h_im - hf;
            lw[t] = w_im - wf;
            idx[t] = hl * W + wl;
            const bool top = hl >= 0, bot = hl + 1 <= H - 1, lef = wl >= 0, rig = wl + 1 <= W - 1;
            cmask |= ((top && lef) ? 1u : 0u) << (4 * (t < G::REF ? t : t - 1) + 0);
            cmask |= ((top && rig) ? 1u : 0u) << (4 * (t < G::REF ? t : t - 1) + 1);
            cmask |= ((bot && lef) ? 1u : 0u) << (4 * (t < G::REF ? t : t - 1) + 2);
            cmask |= ((bot && rig) ? 1u : 0u) << (4 * (t < G::REF ? t : t - 1) + 3);
        }
        if (flags & kPreserve) dp = __ldg(dep + q);
        if (conf) cf = __ldg(conf + q);
    }
    static_assert(K == 3, "corner mask packs 8 deformable taps into 32 bits");

    for (int t = 1; t <= T; ++t) {
        const float *src_prev;
        float *src_next = nullptr;
        if (conf) {
            src_prev = src + (long)((t - 1) % S) * BP;
            if (t < T) src_next = src + (long)(t % S) * BP;
        } else {
            src_prev = t == 1 ? src : list_feat + (long)(t - 2) * BP;
        }
        if (active) {
            const float *im = src_prev + (long)b * P;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < G::KK; ++k) {
                float v;
                if (k == G::REF) {
                    v = __ldcg(im + r);
                } else {
                    const int n = k < G::REF ? k : k - 1;       // slot of this tap in cmask
                    const float *p = im + idx[k];
                    const float v1 = (cmask >> (4 * n + 0)) & 1u ? __ldcg(p) : 0.f;
                    const float v2 = (cmask >> (4 * n + 1)) & 1u ? __ldcg(p + 1) : 0.f;
                    const float v3 = (cmask >> (4 * n + 2)) & 1u ? __ldcg(p + W) : 0.f;
                    const float v4 = (cmask >> (4 * n + 3)) & 1u ? __ldcg(p + W + 1) : 0.f;
                    const float hh = 1.f - lh[k], hw = 1.f - lw[k];
                    v = (hh * hw) * v1 + (hh * lw[k]) * v2 + (lh[k] * hw) * v3 + (lh[k] * lw[k]) * v4;
                }
                acc += v * av[k];
            }
            if (flags & kPreserve) acc = blend_fix(acc, dp);
            if (flags & kAlwaysClip) acc = fmaxf(acc, 0.f);
            list_feat[(long)(t - 1) * BP + q] = acc;
            if (src_next) src_next[q] = acc * cf;
        }
        if (t < T) grid.sync();   // every pixel of iteration t is written (and fenced) before t+1 gathers
    }
}

} // namespace nlspn
