// kernels_fixed.cuh -- fixed-local (no-offset) propagation, the fork's DEFAULT configuration
// (`--offset` is off by default, src/config.py:272-275).  Reference: nlspnmodel.py:209-224 --
// the state is replicate-padded by one pixel and the K^2 = 9 shifted copies are weighted by the
// affinities and summed (CSPN-style; hard-coded to 3x3).  Replicate padding == clamping the
// neighbour index into the image, so one thread per pixel reads 9 clamped neighbours.
// Memory-bound streaming kernels; no tiles needed (neighbour reads are coalesced row segments).
#pragma once
#include "kernels_v2.cuh"

namespace nlspn {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// forward iteration: out = blend(sum_k aff_k * src[clamp(p + tap_k)]), src_next = out * conf
template <bool STREAM>
__global__ void __launch_bounds__(kBlock)
fixed_fwd_kernel(const float *__restrict__ src_prev, const float *__restrict__ aff,
                 const float *__restrict__ conf, const float *__restrict__ dep, unsigned flags,
                 int H, int W, float *__restrict__ out, float *__restrict__ src_next)
{
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    tma::grid_launch_dependents();
    const float *ab = aff + b * 9 * P + r;
    float av[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) av[t] = ld_geo<STREAM>(ab + (long)t * P);
    const long q = b * P + r;
    const float dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    const float cf = (conf && src_next) ? __ldg(conf + q) : 1.f;
    tma::grid_dependency_wait();
    const float *im = src_prev + b * P;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int hh = clampi(h - 1 + t / 3, 0, H - 1), ww = clampi(w - 1 + t % 3, 0, W - 1);
        acc += __ldg(im + hh * W + ww) * av[t];          // feat * aff, summed over dim 1 (:223-224)
    }
    if (flags & kPreserve) acc = blend_fix(acc, dp);
    if (flags & kAlwaysClip) acc = clip_keep_sign(acc);
    out[q] = acc;
    if (src_next) src_next[q] = conf ? acc * cf : acc;
}

// backward pass A: plain (unblocked) scatter planes; 9 scalar REDs to clamped neighbours
template <bool STREAM>
__global__ void __launch_bounds__(kBlock)
fixed_state_kernel(const float *__restrict__ aff, const float *__restrict__ conf,
                   const float *__restrict__ dep, const float *__restrict__ x_t,
                   const float *__restrict__ g_ext, float *__restrict__ s_in, float *__restrict__ s_out,
                   float *__restrict__ gy_out, float *__restrict__ g_conf_acc, unsigned flags, int H, int W)
{
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;
    tma::grid_launch_dependents();
    const float *ab = aff + b * 9 * P + r;
    float av[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) av[t] = ld_geo<STREAM>(ab + (long)t * P);
    const float gext = g_ext ? __ldg(g_ext + q) : 0.f;
    const float cf = conf ? __ldg(conf + q) : 1.f;
    const bool need_x = (s_in && conf) || (flags & kAlwaysClip);
    const float xt = need_x ? __ldg(x_t + q) : 1.f;
    const float dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    tma::grid_dependency_wait();
    float gs = 0.f, gca = 0.f;
    if (s_in) {
        gs = __ldcg(s_in + q);
        if (conf) gca = g_conf_acc[q];
    }
    float Gx = gext;
    if (s_in) Gx += conf ? cf * gs : gs;
    if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
    if (flags & kPreserve) Gx = (1.0f - (dp > 0.f ? 1.f : 0.f)) * Gx;
    if (s_in) {
        s_in[q] = 0.f;
        if (conf) g_conf_acc[q] = gca + xt * gs;
    }
    gy_out[q] = Gx;
    if (Gx == 0.f) return;
    float *so = s_out + b * P;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int hh = clampi(h - 1 + t / 3, 0, H - 1), ww = clampi(w - 1 + t % 3, 0, W - 1);
        atomicAdd(so + hh * W + ww, Gx * av[t]);
    }
}

// backward pass B: raw affinity gradients of all T iterations, accumulators in registers
__global__ void __launch_bounds__(kBlock)
fixed_param_kernel(const float *__restrict__ src, const float *__restrict__ list_feat,
                   const float *__restrict__ gy_all, int has_conf, int H, int W, int T, long BP, long GP,
                   float *__restrict__ g_aff_acc)
{
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;
    int nb[9];
#pragma unroll
    for (int t = 0; t < 9; ++t)
        nb[t] = clampi(h - 1 + t / 3, 0, H - 1) * W + clampi(w - 1 + t % 3, 0, W - 1);
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int t = T; t >= 1; --t) {
        const float gy = __ldg(gy_all + (long)(t - 1) * GP + q);
        if (gy == 0.f) continue;
        const float *im;
        if (has_conf) im = src + (long)(t - 1) * BP + b * P;
        else im = (t == 1 ? src : list_feat + (long)(t - 2) * BP) + b * P;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] += gy * __ldg(im + nb[k]);
    }
    float *gab = g_aff_acc + b * 9 * P + r;
#pragma unroll
    for (int k = 0; k < 9; ++k) gab[(long)k * P] = acc[k];
}

} // namespace nlspn
