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

constexpr int kPersistTH = 8;                       // tile: 32 x 8 pixels, one pixel per thread
constexpr int kPersistBlock = 32 * kPersistTH;
constexpr int kPersistHalo = 8;
constexpr int kPersistBoxW = 32 + 2 * kPersistHalo;
constexpr int kPersistBoxH = kPersistTH + 2 * kPersistHalo;

// grid = (ceil(W/32), ceil(H/8), B), block = (32, 8); cooperative launch (all CTAs co-resident)
template <int K>
__global__ void __launch_bounds__(kPersistBlock, 2)
persist_fwd_kernel(const float *__restrict__ offset, const float *__restrict__ aff,
                   const float *__restrict__ conf, const float *__restrict__ dep, unsigned flags, int H,
                   int W, int B, int T, float *src, int S, float *list_feat)
{
    namespace cg = cooperative_groups;
    using G = Geo<K>;
    static_assert(K == 3, "slow/skip masks pack 8 deformable taps");
    cg::grid_group grid = cg::this_grid();
    __shared__ float box[kPersistBoxH * kPersistBoxW];
    const int P = H * W;
    const long BP = (long)B * P;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * kPersistTH;
    const int b = blockIdx.z;
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool active = w < W && h < H;
    const int r = active ? h * W + w : 0;
    const long q = (long)b * P + r;

    // ---- the pixel's geometry: loaded once, register-resident for all T iterations.
    // Per tap: index of the footprint's top-left corner (in the shared-memory box when the
    // footprint lies inside it, else in the image), the two fractional weights, and for the
    // out-of-box taps four corner-valid bits (guards of cuh:37-48).  Invalid taps (cuh:180) get
    // zero weights and the box origin as index, so the common path is branch-free.
    int idx[G::KK];
    float lh[G::KK], lw[G::KK], av[G::KK];
    unsigned slow = 0u, cmask = 0u;
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
            if (t == G::REF) {
                av[t] = __ldg(ab + (long)t * P);
                idx[t] = (threadIdx.y + kPersistHalo) * kPersistBoxW + threadIdx.x + kPersistHalo;
                continue;
            }
            const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
            const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
            if (!tap_valid(h_im, w_im, H, W)) continue;      // av stays 0: the tap contributes nothing
            av[t] = __ldg(ab + (long)t * P);
            const float hf = floorf(h_im), wf = floorf(w_im);
            const int hl = (int)hf, wl = (int)wf;
            lh[t] = h_im - hf;
            lw[t] = w_im - wf;
            const int ty = hl - (y0 - kPersistHalo), tx = wl - (x0 - kPersistHalo);
            if ((unsigned)ty < (unsigned)(kPersistBoxH - 1) && (unsigned)tx < (unsigned)(kPersistBoxW - 1)) {
                idx[t] = ty * kPersistBoxW + tx;
            } else {
                const int n = t < G::REF ? t : t - 1;
                slow |= 1u << n;
                idx[t] = hl * W + wl;
                const bool top = hl >= 0, bot = hl + 1 <= H - 1, lef = wl >= 0, rig = wl + 1 <= W - 1;
                cmask |= ((top && lef) ? 1u : 0u) << (4 * n + 0);
                cmask |= ((top && rig) ? 1u : 0u) << (4 * n + 1);
                cmask |= ((bot && lef) ? 1u : 0u) << (4 * n + 2);
                cmask |= ((bot && rig) ? 1u : 0u) << (4 * n + 3);
            }
        }
        if (flags & kPreserve) dp = __ldg(dep + q);
        if (conf) cf = __ldg(conf + q);
    }

    for (int t = 1; t <= T; ++t) {
        const float *src_prev;
        float *src_next = nullptr;
        if (conf) {
            src_prev = src + (long)((t - 1) % S) * BP;
            if (t < T) src_next = src + (long)(t % S) * BP;
        } else {
            src_prev = t == 1 ? src : list_feat + (long)(t - 2) * BP;
        }
        const float *im = src_prev + (long)b * P;
        // stage the halo'd tile of the plane: coalesced L2 reads (ld.global.cg -- the plane was
        // written by other SMs in this launch), zero outside the image (= the sampler's padding)
        for (int i = tid; i < kPersistBoxH * kPersistBoxW; i += kPersistBlock) {
            const int by = i / kPersistBoxW, bx = i - by * kPersistBoxW;
            const int yy = y0 - kPersistHalo + by, xx = x0 - kPersistHalo + bx;
            box[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldcg(im + yy * W + xx) : 0.f;
        }
        __syncthreads();
        if (active) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < G::KK; ++k) {
                float v;
                if (k == G::REF) {
                    v = box[idx[k]];
                } else {
                    const int n = k < G::REF ? k : k - 1;
                    float v1, v2, v3, v4;
                    if (slow & (1u << n)) {
                        const float *p = im + idx[k];
                        v1 = (cmask >> (4 * n + 0)) & 1u ? __ldcg(p) : 0.f;
                        v2 = (cmask >> (4 * n + 1)) & 1u ? __ldcg(p + 1) : 0.f;
                        v3 = (cmask >> (4 * n + 2)) & 1u ? __ldcg(p + W) : 0.f;
                        v4 = (cmask >> (4 * n + 3)) & 1u ? __ldcg(p + W + 1) : 0.f;
                    } else {
                        const float *p = box + idx[k];
                        v1 = p[0]; v2 = p[1]; v3 = p[kPersistBoxW]; v4 = p[kPersistBoxW + 1];
                    }
                    const float hh = 1.f - lh[k], hw = 1.f - lw[k];
                    v = (hh * hw) * v1 + (hh * lw[k]) * v2 + (lh[k] * hw) * v3 + (lh[k] * lw[k]) * v4;
                }
                acc += v * av[k];
            }
            if (flags & kPreserve) acc = blend_fix(acc, dp);
            if (flags & kAlwaysClip) acc = clip_keep_sign(acc);
            list_feat[(long)(t - 1) * BP + q] = acc;
            if (src_next) src_next[q] = acc * cf;
        }
        if (t < T) grid.sync();   // iteration t fully written (and fenced); also frees the box for refill
    }
}

} // namespace nlspn
