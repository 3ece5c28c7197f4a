// kernels_step.cuh -- ONE fused propagation iteration, forward and backward (sm_100a).
//
// SURVEY 8(b) "nlspn_step_fwd/_bwd" / 8(f) row f2: the fork's GRU mode re-estimates the affinities
// between iterations (nlspnmodel.py:365-373), so the T-iteration kernels (geometry resident for all
// t) do not apply; the loop body nlspnmodel.py:350-361 then runs as one call per iteration:
//
//   forward   out = clip(blend(G(src_prev; offset, aff))),  src_next = out * conf_fixed
//   backward  G = g_out + conf_fixed * g_src_next ; g_conf = out * g_src_next ; gy = clipmask * (1-m) * G
//             g_aff[k] = gy * bil_k(src_prev) ; g_offset[k] = gy * aff_k * d bil_k / d coord
//             g_src_prev = scatter_k(gy * aff_k * corner weights)
//
// The same backward kernel serves the single-step DCN drop-in (boundary B1,
// modulated_deform_conv_cuda.cu:124-280) with all K^2 taps deformable and a per-tap weight.
//
// Structure: the gather source arrives as a TMA box (as in iter_fwd_tiled_kernel); the scatter is the
// one-vector-RED-per-tap form of kernels_v2.cuh (four phase-shifted 2x2-blocked copies of the target),
// collected into the plain [B,1,H,W] gradient by scatter_collect_kernel.  Measured motivation:
// tools/red_bench.cu, 4x RED.F32 0.410 ms vs 1x RED.F32x4 0.141 ms per 8 KITTI frames.
#pragma once
#include "kernels_fixed.cuh"
#include "kernels_tiled.cuh"

namespace nlspn {

// ---------------------------------------------------------------------------------------------
// DCN forward (B1), tiled: all KK taps deformable, weight[KK], bias[1]
// (modulated_deform_conv_cuda.cu:92-118: im2col + GEMV + bias).  grid = (tiles_x, tiles_y, B).
// ---------------------------------------------------------------------------------------------
template <int K, int TH>
__global__ void __launch_bounds__(kTileW * TH)
dcn_fwd_tiled_kernel(const __grid_constant__ CUtensorMap in_map, const float *__restrict__ input,
                     const float *__restrict__ offset, const float *__restrict__ mask,
                     const float *__restrict__ weight, const float *__restrict__ bias, int H, int W,
                     float *__restrict__ out)
{
    using G = Geo<K>;
    using TG = TileGeo<K, TH>;
    constexpr int kHalo = TG::R;
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
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&bar, TG::BoxBytes);
        tma::load_3d(box, &in_map, &bar, x0 - kHalo, y0 - kHalo, (int)b);
    }
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    float oh[G::KK], ow[G::KK], av[G::KK];
    if (inside) {
        const float *ob = offset + b * 2 * G::KK * P + r;
        const float *ab = mask + b * G::KK * P + r;
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            av[t] = __ldg(ab + (long)t * P);
            oh[t] = __ldg(ob + (long)(2 * t) * P);
            ow[t] = __ldg(ob + (long)(2 * t + 1) * P);
        }
    }
    tma::mbar_wait(&bar, 0);
    if (!inside) return;
    const float *im = input + b * P;
    float acc = __ldg(bias);
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const float h_im = (float)(h - G::PAD + t / K) + oh[t];
        const float w_im = (float)(w - G::PAD + t % K) + ow[t];
        float v = 0.f;
        if (tap_valid(h_im, w_im, H, W)) v = quad_value(box_quad<K, TH>(box, y0, x0, im, H, W, h_im, w_im));
        // the reference multiplies value * mask into the column, then the GEMV applies the weight
        acc += __ldg(weight + t) * (v * av[t]);
    }
    out[b * P + r] = acc;
}

// gy of one step (and the confidence gradient, stored): see the file header
template <bool DCN>
__device__ __forceinline__ float step_bwd_gy(const float *__restrict__ conf, const float *__restrict__ dep,
                                             const float *__restrict__ x_out, const float *__restrict__ g_out,
                                             const float *__restrict__ g_src_next, unsigned flags, long q,
                                             float *__restrict__ g_conf)
{
    float Gx = g_out ? __ldg(g_out + q) : 0.f;
    if (DCN) return Gx;
    const float gsn = g_src_next ? __ldg(g_src_next + q) : 0.f;
    const float xt = (g_src_next || (flags & kAlwaysClip)) ? __ldg(x_out + q) : 1.f;
    if (g_src_next) Gx += conf ? __ldg(conf + q) * gsn : gsn;
    if (g_conf) g_conf[q] = g_src_next ? xt * gsn : 0.f;                // d(out * c)/dc, nlspnmodel.py:351
    if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
    if (flags & kPreserve) Gx = (1.0f - (__ldg(dep + q) > 0.f ? 1.f : 0.f)) * Gx;
    return Gx;
}

// The K^2 taps of one pixel's backward.  quad_at(h_im, w_im) -> Quad fetches the four corner values.
//   BLOCKED: `so` = this image's four phase planes, one vector RED per tap; else `so` = plain plane, scalar REDs.
template <int K, bool DCN, bool BLOCKED, typename QuadFn>
__device__ __forceinline__ void step_bwd_taps(float gy, const float (&av)[K * K], const float (&oh)[K * K],
                                              const float (&ow)[K * K], const float *__restrict__ weight, int h,
                                              int w, int H, int W, float centre, QuadFn quad_at,
                                              float *__restrict__ so, const ScatterGeo &sg,
                                              float *__restrict__ gob, float *__restrict__ gab)
{
    using G = Geo<K>;
    const int P = H * W;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const float col = DCN ? __ldg(weight + t) * gy : gy;
        if (!DCN && t == G::REF) {
            gab[(long)t * P] = col * centre;
            gob[(long)(2 * t) * P] = 0.f;
            gob[(long)(2 * t + 1) * P] = 0.f;
            if (col != 0.f) atomicAdd(BLOCKED ? so + scatter_cell(sg, 0, 0, h + 1, w + 1) : so + h * W + w, col * av[t]);
            continue;
        }
        const float h_im = (float)(h - G::PAD + t / K) + oh[t];
        const float w_im = (float)(w - G::PAD + t % K) + ow[t];
        float ga = 0.f, gh = 0.f, gw = 0.f;
        if (tap_valid(h_im, w_im, H, W)) {
            const Quad qd = quad_at(h_im, w_im);
            ga = col * quad_value(qd);                                   // cuh:314-315
            const float top = col * av[t];
            // mdmcn_get_coordinate_weight, cuh:101-122 (expressions kept literal)
            const float wl1 = (float)(qd.wl + 1) - w_im, wl0 = w_im - (float)qd.wl;
            const float hl1 = (float)(qd.hl + 1) - h_im, hl0 = h_im - (float)qd.hl;
            const float dh = -1.f * wl1 * qd.v1 + -1.f * wl0 * qd.v2 + wl1 * qd.v3 + wl0 * qd.v4;
            const float dw = -1.f * hl1 * qd.v1 + hl1 * qd.v2 + -1.f * hl0 * qd.v3 + hl0 * qd.v4;
            gh = dh * top;
            gw = dw * top;
            if (top != 0.f) {
                // mdmcn_get_gradient_weight, cuh:71-79
                const float th = (float)(qd.hl + 1) - h_im, bh = (h_im + 1.f) - (float)(qd.hl + 1);
                const float lw_ = (float)(qd.wl + 1) - w_im, rw = (w_im + 1.f) - (float)(qd.wl + 1);
                if (BLOCKED) {   // one vector RED into the phase plane of the footprint (padding absorbs the border)
                    const int Y = qd.hl + 1, X = qd.wl + 1;
                    const int sy = Y & 1, sx = X & 1;
                    float4 *blk = reinterpret_cast<float4 *>(
                        so + (long)(sy * 2 + sx) * sg.plane + ((long)((Y + sy) >> 1) * sg.Wb + ((X + sx) >> 1)) * 4);
                    atomicAdd(blk, make_float4(th * lw_ * top, th * rw * top, bh * lw_ * top, bh * rw * top));
                } else {
                    const bool topv = qd.hl >= 0, botv = qd.hl + 1 <= H - 1;
                    const bool lefv = qd.wl >= 0, rigv = qd.wl + 1 <= W - 1;
                    float *sp = so + (long)qd.hl * W + qd.wl;
                    if (topv && lefv) atomicAdd(sp, th * lw_ * top);
                    if (topv && rigv) atomicAdd(sp + 1, th * rw * top);
                    if (botv && lefv) atomicAdd(sp + W, bh * lw_ * top);
                    if (botv && rigv) atomicAdd(sp + W + 1, bh * rw * top);
                }
            }
        }
        gab[(long)t * P] = ga;
        gob[(long)(2 * t) * P] = gh;
        gob[(long)(2 * t + 1) * P] = gw;
    }
}

// ---------------------------------------------------------------------------------------------
// Single-iteration backward, tiled.  grid = (tiles_x, tiles_y, B), block = (32, TH).
//   DCN = false  NLSPN step: centre tap has a structurally zero offset (nlspnmodel.py:256); its
//                offset-gradient pair is written as zeros; conf / dep / clip handling as above.
//   DCN = true   boundary B1: gy = g_out, col_k = weight[k] * gy (the mm at cu:221), all taps deformable.
// planes: the four blocked scatter copies [B][sg.image], zero on entry.
// ---------------------------------------------------------------------------------------------
template <int K, int TH, bool DCN>
__global__ void __launch_bounds__(kTileW * TH)
step_bwd_tiled_kernel(const __grid_constant__ CUtensorMap src_map, const float *__restrict__ src_prev,
                      const float *__restrict__ offset, const float *__restrict__ aff,
                      const float *__restrict__ conf, const float *__restrict__ dep,
                      const float *__restrict__ x_out, const float *__restrict__ g_out,
                      const float *__restrict__ g_src_next, const float *__restrict__ weight, unsigned flags,
                      int H, int W, float *__restrict__ planes, float *__restrict__ g_off,
                      float *__restrict__ g_aff, float *__restrict__ g_conf)
{
    using G = Geo<K>;
    using TG = TileGeo<K, TH>;
    constexpr int kHalo = TG::R;
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
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&bar, TG::BoxBytes);
        tma::load_3d(box, &src_map, &bar, x0 - kHalo, y0 - kHalo, (int)b);
    }
    const int w = x0 + threadIdx.x, h = y0 + threadIdx.y;
    const bool inside = w < W && h < H;
    const int r = inside ? h * W + w : 0;
    const long q = b * P + r;
    // ---- every global load first
    float oh[G::KK], ow[G::KK], av[G::KK];
    float gy = 0.f;
    if (inside) {
        const float *ob = offset + b * 2 * G::KK * P + r;
        const float *ab = aff + b * G::KK * P + r;
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            av[t] = __ldg(ab + (long)t * P);
            oh[t] = ow[t] = 0.f;
            if (DCN || t != G::REF) {
                oh[t] = __ldg(ob + (long)(2 * t) * P);
                ow[t] = __ldg(ob + (long)(2 * t + 1) * P);
            }
        }
        gy = step_bwd_gy<DCN>(conf, dep, x_out, g_out, g_src_next, flags, q, g_conf);
    }
    tma::mbar_wait(&bar, 0);
    if (!inside) return;

    const ScatterGeo sg = scatter_geo(H, W);
    const float *im = src_prev + b * P;
    const float centre = box[(threadIdx.y + kHalo) * TG::BoxW + threadIdx.x + kHalo];
    step_bwd_taps<K, DCN, true>(gy, av, oh, ow, weight, h, w, H, W, centre,
                                [&](float h_im, float w_im) { return box_quad<K, TH>(box, y0, x0, im, H, W, h_im, w_im); },
                                planes + b * sg.image, sg, g_off + b * 2 * G::KK * P + r, g_aff + b * G::KK * P + r);
}

// Direct-gather form of the same step (no TMA box, scalar REDs straight into the plain gradient plane):
// serves shapes the tiled kernel cannot (W % 4 != 0) and the "dcn_blocked = 0" comparison.
// g_src_prev must be zero on entry.  grid = (ceil(P/256), B).
template <int K, bool DCN>
__global__ void __launch_bounds__(kBlock)
step_bwd_direct_kernel(const float *__restrict__ src_prev, const float *__restrict__ offset,
                       const float *__restrict__ aff, const float *__restrict__ conf,
                       const float *__restrict__ dep, const float *__restrict__ x_out,
                       const float *__restrict__ g_out, const float *__restrict__ g_src_next,
                       const float *__restrict__ weight, unsigned flags, int H, int W,
                       float *__restrict__ g_src_prev, float *__restrict__ g_off, float *__restrict__ g_aff,
                       float *__restrict__ g_conf)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;
    float oh[G::KK], ow[G::KK], av[G::KK];
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        av[t] = __ldg(ab + (long)t * P);
        oh[t] = ow[t] = 0.f;
        if (DCN || t != G::REF) {
            oh[t] = __ldg(ob + (long)(2 * t) * P);
            ow[t] = __ldg(ob + (long)(2 * t + 1) * P);
        }
    }
    const float gy = step_bwd_gy<DCN>(conf, dep, x_out, g_out, g_src_next, flags, q, g_conf);
    const float *im = src_prev + b * P;
    const ScatterGeo sg = scatter_geo(H, W);
    step_bwd_taps<K, DCN, false>(gy, av, oh, ow, weight, h, w, H, W, DCN ? 0.f : __ldg(im + r),
                                 [&](float h_im, float w_im) { return load_quad(im, H, W, h_im, w_im); },
                                 g_src_prev + b * P, sg, g_off + b * 2 * G::KK * P + r, g_aff + b * G::KK * P + r);
}

// plain plane <- sum of the four phase copies (out-of-image corners fell into padding cells)
__global__ void __launch_bounds__(kBlock)
scatter_collect_kernel(const float *__restrict__ planes, int H, int W, float *__restrict__ out)
{
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const ScatterGeo sg = scatter_geo(H, W);
    const float *sl = planes + b * sg.image;
    const int h = r / W, w = r - h * W;
    float cv[4];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) cv[ph] = __ldcs(sl + scatter_cell(sg, ph >> 1, ph & 1, h + 1, w + 1));
    out[b * P + r] = ((cv[0] + cv[1]) + cv[2]) + cv[3];
}

// ---------------------------------------------------------------------------------------------
// Fixed-local (no-offset) single-iteration backward, nlspnmodel.py:209-224 transposed.
// g_src_prev must be zero on entry (coalesced scalar REDs to the clamped neighbours).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
fixed_step_bwd_kernel(const float *__restrict__ src_prev, const float *__restrict__ aff,
                      const float *__restrict__ conf, const float *__restrict__ dep,
                      const float *__restrict__ x_out, const float *__restrict__ g_out,
                      const float *__restrict__ g_src_next, unsigned flags, int H, int W,
                      float *__restrict__ g_src_prev, float *__restrict__ g_aff, float *__restrict__ g_conf)
{
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;
    const float *ab = aff + b * 9 * P + r;
    const float *im = src_prev + b * P;
    float av[9], sv[9];
    int nb[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        nb[t] = clampi(h - 1 + t / 3, 0, H - 1) * W + clampi(w - 1 + t % 3, 0, W - 1);
        av[t] = __ldg(ab + (long)t * P);
        sv[t] = __ldg(im + nb[t]);
    }
    float Gx = g_out ? __ldg(g_out + q) : 0.f;
    const float gsn = g_src_next ? __ldg(g_src_next + q) : 0.f;
    const float xt = (g_src_next || (flags & kAlwaysClip)) ? __ldg(x_out + q) : 1.f;
    if (g_src_next) Gx += conf ? __ldg(conf + q) * gsn : gsn;
    if (g_conf) g_conf[q] = g_src_next ? xt * gsn : 0.f;
    if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
    if (flags & kPreserve) Gx = (1.0f - (__ldg(dep + q) > 0.f ? 1.f : 0.f)) * Gx;
    float *gab = g_aff + b * 9 * P + r;
    float *so = g_src_prev + b * P;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        gab[(long)t * P] = Gx * sv[t];
        if (Gx != 0.f) atomicAdd(so + nb[t], Gx * av[t]);
    }
}

} // namespace nlspn
