// kernels_v2.cuh -- the two-pass backward of the NLSPN path (sm_100a).
//
// The reference's backward re-reads and re-writes the offset/affinity gradient accumulators
// every iteration (autograd adds T tensors; 24N B/px*iter) and scatters the state gradient with
// 4 scalar atomics per tap.  Two observations remove both costs:
//
//  (1) The state-gradient recursion  gs_{t-1} = M^T gy_t  needs only the sampling GEOMETRY
//      (offsets, affinities) -- not the sampled values.  Pass A (bwd_state_kernel, one launch
//      per iteration) therefore does no gather at all: it forms gy_t, stores it, and scatters.
//      The scatter target is kept in four copies of the plane stored as 2x2 blocks with the
//      four possible block phases; every bilinear footprint is then exactly ONE aligned 16-byte
//      block in ONE of the copies, i.e. ONE red.global.add.v4.f32 (REDG.E.ADD.F32x4) per tap
//      instead of 4 scalar REDs.  Measured on B200 (tools/red_bench.cu, 8x KITTI frame, iid
//      offsets): 0.410 ms for 4x RED.F32 vs 0.141 ms for 1x RED.F32x4 vs 0.121 ms for a plain
//      128-bit store.  Out-of-image corners land in padding cells that are never read, so the
//      per-corner guards of cuh:37-48 disappear from the scatter.
//
//  (2) The parameter gradients  d/d offset, d/d affinity  are sums over t of per-pixel terms
//      gy_t[p] * f(src_{t-1} around p).  Pass B (bwd_param_kernel, ONE launch) keeps a pixel's
//      offsets, affinities and all 3N accumulators in registers and loops over t itself,
//      gathering from the saved src planes.  The accumulators never touch memory until the
//      fused prologue-backward epilogue writes g_guidance once.
//
// Per pixel*iteration this moves ~124 B (pass A) + ~12 B (pass B) through HBM instead of the
// 36N+36 = 324 B of the per-iteration formulation (K=3).
#pragma once
#include "kernels_v1.cuh"
#include "tma.cuh"

namespace nlspn {

// ======================================================================================
// Pass A: one backward iteration of the state gradient.
//   gs   = sum of the four phase copies of S_in at this pixel (then zeroed)
//   G    = g_list[t] + conf * gs ;  g_conf_acc += x_t * gs        (nlspnmodel.py:351)
//   gy   = (1-m) * G  [clip mask]   -> stored for pass B            (nlspnmodel.py:357,361)
//   S_out[footprint(tap)] += gy * aff[tap] * corner weights          (cuh:229-252, 71-79)
// ======================================================================================
// load with an optional "stream once" (evict-first) hint: used for offsets/affinities when the
// image group is too large to stay L2-resident across iterations
template <bool STREAM>
__device__ __forceinline__ float ld_geo(const float *p)
{
    return STREAM ? __ldcs(p) : __ldg(p);
}

// Zeroing of the scatter planes.  Two sets (s_zero == nullptr): every thread clears the four cells it just
// read -- four 4-byte stores that fill half of each 32-byte sector they touch (1 sector per pixel on the
// SM->L2 write port, which is the port the REDs saturate).  Three sets (s_zero != nullptr): the set that
// iteration t+1 consumed is cleared LINEARLY with one 16-byte store per thread (0.5 sector per pixel)
// while this iteration reads the second set and scatters into the third.
__device__ __forceinline__ void zero_set_linear(float *__restrict__ set_img, long n4, long i0, long stride)
{
    float4 *p = reinterpret_cast<float4 *>(set_img);
    for (long i = i0; i < n4; i += stride) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int K, bool STREAM, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
bwd_state_kernel(const float *__restrict__ offset, const float *__restrict__ aff,
                 const float *__restrict__ conf, const float *__restrict__ dep,
                 const float *__restrict__ x_t, const float *__restrict__ g_ext,
                 float *__restrict__ s_in, float *__restrict__ s_out, float *__restrict__ gy_out,
                 float *__restrict__ g_conf_acc, unsigned flags, int H, int W, float *__restrict__ s_zero)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    const long b = blockIdx.y;
    const ScatterGeo sg = scatter_geo(H, W);
    if (r >= P) {   // tail threads only help clearing the third set
        if (s_zero) {
            tma::grid_dependency_wait();
            zero_set_linear(s_zero + b * sg.image, sg.image / 4, r, (long)gridDim.x * kBlock);
        }
        return;
    }
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;

    // ---- stage 1: issue EVERY load before the first store, so one memory round trip covers
    // the streamed geometry, the four scatter cells and the per-pixel planes.
    // PDL: the loads that do not depend on the previous backward iteration go first; the next
    // launch may start prefetching its own geometry while this grid drains.
    tma::grid_launch_dependents();
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
    float oh[G::KK], ow[G::KK], av[G::KK];
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        av[t] = ld_geo<STREAM>(ab + (long)t * P);
        if (t != G::REF) {
            oh[t] = ld_geo<STREAM>(ob + (long)(2 * t) * P);
            ow[t] = ld_geo<STREAM>(ob + (long)(2 * t + 1) * P);
        }
    }
    const float gext = g_ext ? __ldg(g_ext + q) : 0.f;
    const float cf = conf ? __ldg(conf + q) : 1.f;
    const bool need_x = (s_in && conf) || (flags & kAlwaysClip);
    const float xt = need_x ? __ldg(x_t + q) : 1.f;
    const float dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    tma::grid_dependency_wait();   // s_in / s_out / g_conf_acc belong to the previous launch
    float *si = s_in ? s_in + b * sg.image : nullptr;
    long cell[4];
    float cv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        cell[ph] = scatter_cell(sg, ph >> 1, ph & 1, h + 1, w + 1);
        if (si) cv[ph] = __ldcg(si + cell[ph]);
    }
    const float gca = (si && conf) ? g_conf_acc[q] : 0.f;

    // ---- stage 2: arithmetic
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

    // ---- stage 3: plain stores
    if (s_zero) {
        zero_set_linear(s_zero + b * sg.image, sg.image / 4, r, (long)gridDim.x * kBlock);
    } else if (si) {
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) si[cell[ph]] = 0.f;
    }
    if (si && conf) g_conf_acc[q] = gca + xt * gs;
    gy_out[q] = gy;
    if (gy == 0.f) return; // fixed pixels (and exact zeros) contribute nothing to the scatter

    // ---- stage 4: one vector RED per tap
    float *so = s_out + b * sg.image;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const float top = gy * av[t];
        if (t == G::REF) {
            atomicAdd(so + cell[0], top);
            continue;
        }
        const float h_im = (float)(h - G::PAD + t / K) + oh[t];
        const float w_im = (float)(w - G::PAD + t % K) + ow[t];
        if (!tap_valid(h_im, w_im, H, W)) continue;
        float hf, wf;
        int hl, wl;
        floor_small(h_im, hf, hl);
        floor_small(w_im, wf, wl);
        // mdmcn_get_gradient_weight, cuh:71-79 (expressions kept literal; (float)(hl+1) == hf + 1 exactly)
        const float h1 = hf + 1.f, w1 = wf + 1.f;
        const float th = h1 - h_im, bh = (h_im + 1.f) - h1;
        const float lw_ = w1 - w_im, rw = (w_im + 1.f) - w1;
        const int Y = hl + 1, X = wl + 1;         // padded coordinates of the top-left corner
        const int sy = Y & 1, sx = X & 1;
        float4 *blk = reinterpret_cast<float4 *>(
            so + (long)(sy * 2 + sx) * sg.plane + ((long)((Y + sy) >> 1) * sg.Wb + ((X + sx) >> 1)) * 4);
        atomicAdd(blk, make_float4(th * lw_ * top, th * rw * top, bh * lw_ * top, bh * rw * top));
    }
}

// ======================================================================================
// Pass B: offset / affinity gradients of ALL T iterations in one launch.
// Each thread owns one pixel and a chunk of C taps (blockIdx.z selects the chunk; K=3 is a
// single chunk of all 9 taps, K=5 three chunks, K=7 six): the chunk's offsets, affinities and
// 3C accumulators stay in registers while the thread replays t = T..1 (the order autograd
// accumulates in), gathering from the saved src planes.  Results are written ONCE: offset
// gradients straight into g_guidance's offset channels, raw affinity gradients into g_aff_acc
// for the normalisation backward (final_bwd_kernel).
// ======================================================================================
constexpr int kParamBlock = 128;

template <int K, int C>
__global__ void __launch_bounds__(kParamBlock)
bwd_param_kernel(const float *__restrict__ offset, const float *__restrict__ aff,
                 const float *__restrict__ src, const float *__restrict__ list_feat,
                 const float *__restrict__ gy_all, int has_conf, int H, int W, int T, long BP, long GP,
                 float *__restrict__ g_guidance, float *__restrict__ g_aff_acc)
{
    using G = Geo<K>;
    constexpr int NCH = (G::KK + C - 1) / C;
    const int P = H * W;
    const int r = blockIdx.x * kParamBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int k0 = NCH == 1 ? 0 : (int)blockIdx.z * C;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
    float oh[C], ow[C], av[C], acc_h[C], acc_w[C], acc_a[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int k = k0 + c;
        oh[c] = ow[c] = av[c] = 0.f;
        acc_h[c] = acc_w[c] = acc_a[c] = 0.f;
        if (k < G::KK) {
            av[c] = __ldg(ab + (long)k * P);
            if (k != G::REF) {
                oh[c] = __ldg(ob + (long)(2 * k) * P);
                ow[c] = __ldg(ob + (long)(2 * k + 1) * P);
            }
        }
    }
    for (int t = T; t >= 1; --t) {
        const float gy = __ldg(gy_all + (long)(t - 1) * GP + q);
        if (gy == 0.f) continue;
        const float *im;
        if (has_conf) im = src + (long)(t - 1) * BP + b * P;
        else im = (t == 1 ? src : list_feat + (long)(t - 2) * BP) + b * P;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int k = k0 + c;
            if (k >= G::KK) continue;
            if (k == G::REF) {
                acc_a[c] += gy * __ldg(im + r);
                continue;
            }
            const float h_im = (float)(h - G::PAD + k / K) + oh[c];
            const float w_im = (float)(w - G::PAD + k % K) + ow[c];
            if (!tap_valid(h_im, w_im, H, W)) continue;
            const Quad qd = load_quad(im, H, W, h_im, w_im);
            acc_a[c] += gy * quad_value(qd);                       // cuh:314-315
            const float top = gy * av[c];
            // mdmcn_get_coordinate_weight, cuh:101-122 (expressions kept literal)
            const float wl1 = (float)(qd.wl + 1) - w_im, wl0 = w_im - (float)qd.wl;
            const float hl1 = (float)(qd.hl + 1) - h_im, hl0 = h_im - (float)qd.hl;
            const float dh = -1.f * wl1 * qd.v1 + -1.f * wl0 * qd.v2 + wl1 * qd.v3 + wl0 * qd.v4;
            const float dw = -1.f * hl1 * qd.v1 + hl1 * qd.v2 + -1.f * hl0 * qd.v3 + hl0 * qd.v4;
            acc_h[c] += dh * top;
            acc_w[c] += dw * top;
        }
    }
    float *ggb = g_guidance + b * 3 * G::N * P + r;
    float *gab = g_aff_acc + b * G::KK * P + r;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int k = k0 + c;
        if (k >= G::KK) continue;
        gab[(long)k * P] = acc_a[c];
        if (k != G::REF) {
            const int n = k < G::REF ? k : k - 1;
            ggb[(long)(2 * n) * P] = acc_h[c];
            ggb[(long)(2 * n + 1) * P] = acc_w[c];
        }
    }
}

} // namespace nlspn
