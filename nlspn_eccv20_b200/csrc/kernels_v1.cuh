// kernels_v1.cuh -- straightforward one-thread-per-pixel kernels (sm_100a).
//
// These are the CORRECTNESS BASELINE kernels: the prologue, a direct-gather forward iteration,
// the reverse-replay backward iteration and the fused final/prologue-backward kernel.  They are
// also what the single-step DCN boundary (B1) runs.  The tiled/persistent kernels in
// kernels_tiled.cuh replace the two iteration kernels on the NLSPN path when shapes allow.
#pragma once
#include "common.cuh"

namespace nlspn {

enum : unsigned { kPreserve = 1u, kAlwaysClip = 2u, kNoOffset = 4u, kBlendPre = 8u, kConfSampled = 16u,
                  kLegacy = 32u, kDeterministic = 0x200u };
enum : int { kAS = 0, kASS = 1, kTC = 2, kTGASS = 3 };
constexpr int kGammaSlots = 64;   // partial sums of d loss / d gamma (final_bwd_kernel -> gamma_reduce_kernel)

// Upstream conf_prop (right column; restated from the north-star prose -- parity
// unpinned): the confidence seen by neighbour n is the confidence map sampled by a 1x1 modulated
// deformable gather (pad 0, unit mask/weight, zero bias) at that neighbour's offset; the tap
// displacement is added to the offset only in --legacy mode.  Returns the sampled value and, if
// asked, the sampling coordinates.
template <int K>
__device__ __forceinline__ float sample_conf(const float *__restrict__ conf_img, const float *__restrict__ gb,
                                             int n, int h, int w, int H, int W, int P, bool legacy,
                                             float *h_out = nullptr, float *w_out = nullptr)
{
    using G = Geo<K>;
    const int t = n < G::REF ? n : n + 1;
    float oh = __ldg(gb + (long)(2 * n) * P), ow = __ldg(gb + (long)(2 * n + 1) * P);
    if (legacy) {
        oh = oh + (float)(t / K - G::PAD);
        ow = ow + (float)(t % K - G::PAD);
    }
    const float h_im = (float)h + oh, w_im = (float)w + ow;
    if (h_out) {
        *h_out = h_im;
        *w_out = w_im;
    }
    if (!tap_valid(h_im, w_im, H, W)) return 0.f;
    return quad_value(load_quad(conf_img, H, W, h_im, w_im));
}

// ======================================================================================
// Prologue.  nlspnmodel.py:252-259 (_off_insert), :179-201 (_affinity_normalization),
// :261-269 (_aff_insert), :328-334 (mask_fix / confidence), :341-351 (first blend + premul).
// ======================================================================================
template <int K>
__global__ void __launch_bounds__(kBlock)
prologue_fwd_kernel(const float *__restrict__ guidance, const float *__restrict__ conf,
                    const float *__restrict__ init, const float *__restrict__ dep,
                    const float *__restrict__ gamma_ptr, int affinity, unsigned flags, int H, int W,
                    float *__restrict__ offset, float *__restrict__ aff, float *__restrict__ conf_out,
                    float *__restrict__ src0)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    // fixed-local mode (args.offset False): guidance holds the N raw affinities only
    const bool no_off = (flags & kNoOffset) != 0;
    const float *gb = guidance + b * (no_off ? 1 : 3) * G::N * P + r;
    float *ob = offset + b * 2 * G::KK * P + r;
    float *ab = aff + b * G::KK * P + r;
    const int aff_ch0 = no_off ? 0 : 2 * G::N;   // first raw-affinity channel of guidance

#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        if (no_off) break;
        if (t == G::REF) {
            ob[(long)(2 * t) * P] = 0.f;
            ob[(long)(2 * t + 1) * P] = 0.f;
        } else {
            const int n = t < G::REF ? t : t - 1;
            ob[(long)(2 * t) * P] = __ldg(gb + (long)(2 * n) * P);
            ob[(long)(2 * t + 1) * P] = __ldg(gb + (long)(2 * n + 1) * P);
        }
    }

    float a[G::N];
    float abs_sum = 0.f;
    const bool use_tanh = affinity == kTC || affinity == kTGASS;
    const float gamma = __ldg(gamma_ptr);
    const float g = affinity == kTGASS ? gamma + 1e-8f : gamma;
    const bool sampled = (flags & kConfSampled) != 0;     // upstream conf_prop (parity unpinned)
    const int hh0 = r / W, ww0 = r - hh0 * W;
    // all loads ahead of the first tanhf (its branches would fence them into small groups)
    const long q = b * P + r;
    const bool preserve = (flags & kPreserve) != 0;
#pragma unroll
    for (int n = 0; n < G::N; ++n) a[n] = __ldg(gb + (long)(aff_ch0 + n) * P);
    const float d = preserve ? __ldg(dep + q) : 0.f;
    float x = __ldg(init + q);
    float c = (conf && !sampled) ? __ldg(conf + q) : 1.f;
#pragma unroll
    for (int n = 0; n < G::N; ++n) {
        float v = a[n];
        if (use_tanh) v = tanhf(v) / g;
        if (sampled) v *= sample_conf<K>(conf + b * P, gb, n, hh0, ww0, H, W, P, (flags & kLegacy) != 0);
        a[n] = v;
        abs_sum += fabsf(v);
    }
    abs_sum += 1e-4f;
    if ((affinity == kASS || affinity == kTGASS) && abs_sum < 1.0f) abs_sum = 1.0f;
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < G::N; ++n) {
        if (affinity != kTC) a[n] = a[n] / abs_sum;
        sum += a[n];
    }
#pragma unroll
    for (int t = 0; t < G::KK; ++t)
        ab[(long)t * P] = t == G::REF ? 1.0f - sum : a[t < G::REF ? t : t - 1];

    if (preserve) x = blend_fix(x, d);
    if (flags & kAlwaysClip) x = fmaxf(x, 0.f);
    if (conf && !sampled) {
        if (preserve) {
            const float m = d > 0.f ? 1.f : 0.f;
            c = (1.0f - m) * c + m;
        }
        conf_out[q] = c;
        x = x * c;
    }
    src0[q] = x;
}

// ======================================================================================
// One propagation iteration, direct gather from global memory.
// CENTER_ZERO = true : NLSPN path; the centre tap has a structurally zero offset
//                      (nlspnmodel.py:256) so it is the pixel's own value -- one coalesced
//                      load instead of four gathered ones; weight = 1, bias = 0.
// CENTER_ZERO = false: general single step (boundary B1): all KK taps deformable, weight
//                      and bias applied (modulated_deform_conv_cuda.cu:112).
// Epilogue (NLSPN path): blend, clamp, append to list_feat, pre-multiply for the next gather.
// ======================================================================================
template <int K, bool CENTER_ZERO>
__global__ void __launch_bounds__(kBlock)
iter_fwd_kernel(const float *__restrict__ src_prev, const float *__restrict__ offset,
                const float *__restrict__ aff, const float *__restrict__ conf,
                const float *__restrict__ dep, const float *__restrict__ weight,
                const float *__restrict__ bias, unsigned flags, int H, int W,
                float *__restrict__ out, float *__restrict__ src_next)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const float *im = src_prev + b * P;
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;

    float acc = (!CENTER_ZERO && bias) ? __ldg(bias) : 0.f;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const int i = t / K, j = t % K;
        const float a = __ldg(ab + (long)t * P);
        float v;
        if (CENTER_ZERO && t == G::REF) {
            v = __ldg(im + r);
        } else {
            const float oh = __ldg(ob + (long)(2 * t) * P);
            const float ow = __ldg(ob + (long)(2 * t + 1) * P);
            const float h_im = (float)(h - G::PAD + i) + oh;
            const float w_im = (float)(w - G::PAD + j) + ow;
            v = 0.f;
            if (tap_valid(h_im, w_im, H, W)) v = quad_value(load_quad(im, H, W, h_im, w_im));
        }
        if (!CENTER_ZERO && weight)
            acc += __ldg(weight + t) * (v * a);
        else
            acc += v * a;
    }

    const long q = b * P + r;
    if (CENTER_ZERO) {
        if (flags & kBlendPre) {
            // upstream order: the list gets the raw gather, the NEXT gather reads the blended state
            out[q] = acc;
            if (src_next) src_next[q] = (flags & kPreserve) ? blend_fix(acc, __ldg(dep + q)) : acc;
        } else {
            if (flags & kPreserve) acc = blend_fix(acc, __ldg(dep + q));
            if (flags & kAlwaysClip) acc = clip_keep_sign(acc);
            out[q] = acc;
            if (src_next) src_next[q] = conf ? acc * __ldg(conf + q) : acc;
        }
    } else {
        out[q] = acc;
    }
}

// ======================================================================================
// One backward iteration (reverse replay).  Per pixel:
//   carried gradient  G = g_list[t] + c * S_in        (S_in = scatter result of iteration t+1)
//   g_conf_acc       += x_t * S_in                     (d(x*c)/dc, nlspnmodel.py:351)
//   gy = (1-m) * G                                     (d blend, nlspnmodel.py:357)
//   per tap: grad_aff += gy * bilinear                 (cuh:314-315)
//            grad_off += gy * aff * d bilinear/d coord (cuh:316-323, mdmcn_get_coordinate_weight)
//            S_out[corner] += gy * aff * corner weight (cuh:229-252, mdmcn_get_gradient_weight)
// grad_off / grad_aff accumulate over the T iterations in place (`first` stores instead of
// adding, so no memset); S_in is zeroed after it is read so the two scatter planes ping-pong.
// CENTER_ZERO as in the forward kernel.  For CENTER_ZERO=false (B1) g_off is laid out
// [B,2KK,P]; for the NLSPN path it is the offset part of g_guidance, [B,3N,P] with N pairs.
// ======================================================================================
template <int K, bool CENTER_ZERO>
__global__ void __launch_bounds__(kBlock)
iter_bwd_kernel(const float *__restrict__ src_prev, const float *__restrict__ offset,
                const float *__restrict__ aff, const float *__restrict__ conf,
                const float *__restrict__ dep, const float *__restrict__ x_t,
                const float *__restrict__ g_ext, const float *__restrict__ weight,
                float *__restrict__ s_in, float *__restrict__ s_out, float *__restrict__ g_off,
                long g_off_batch_stride, float *__restrict__ g_aff, float *__restrict__ g_conf_acc,
                unsigned flags, int first, int H, int W)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    const int h = r / W, w = r - h * W;

    float Gx = g_ext ? __ldg(g_ext + q) : 0.f;
    if (s_in) {
        const float gs = s_in[q];
        s_in[q] = 0.f;
        if (conf) {
            Gx += __ldg(conf + q) * gs;
            g_conf_acc[q] += __ldg(x_t + q) * gs;
        } else {
            Gx += gs;
        }
    }
    if (CENTER_ZERO) {
        if ((flags & kAlwaysClip) && was_clipped(__ldg(x_t + q))) Gx = 0.f;
        if (flags & kPreserve) Gx = (1.0f - (__ldg(dep + q) > 0.f ? 1.f : 0.f)) * Gx;
    }
    const float gy = Gx;

    const float *im = src_prev + b * P;
    float *so = s_out + b * P;
    const float *ob = offset + b * 2 * G::KK * P + r;
    const float *ab = aff + b * G::KK * P + r;
    float *gob = g_off + b * g_off_batch_stride + r;
    float *gab = g_aff + b * G::KK * P + r;

#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const int i = t / K, j = t % K;
        const float a = __ldg(ab + (long)t * P);
        const float col = (!CENTER_ZERO && weight) ? __ldg(weight + t) * gy : gy;
        if (CENTER_ZERO && t == G::REF) {
            const float ga = col * __ldg(im + r);
            if (first) gab[(long)t * P] = ga; else gab[(long)t * P] += ga;
            atomicAdd(so + r, col * a);
            continue;
        }
        const float oh = __ldg(ob + (long)(2 * t) * P);
        const float ow = __ldg(ob + (long)(2 * t + 1) * P);
        const float h_im = (float)(h - G::PAD + i) + oh;
        const float w_im = (float)(w - G::PAD + j) + ow;
        float ga = 0.f, gh = 0.f, gw = 0.f;
        if (tap_valid(h_im, w_im, H, W)) {
            const Quad qd = load_quad(im, H, W, h_im, w_im);
            ga = col * quad_value(qd);
            const float top = col * a;
            // mdmcn_get_coordinate_weight, cuh:101-122 (expressions kept literal)
            const float wl1 = (float)(qd.wl + 1) - w_im, wl0 = w_im - (float)qd.wl;
            const float hl1 = (float)(qd.hl + 1) - h_im, hl0 = h_im - (float)qd.hl;
            const float dh = -1.f * wl1 * qd.v1 + -1.f * wl0 * qd.v2 + wl1 * qd.v3 + wl0 * qd.v4;
            const float dw = -1.f * hl1 * qd.v1 + hl1 * qd.v2 + -1.f * hl0 * qd.v3 + hl0 * qd.v4;
            gh = dh * top;
            gw = dw * top;
            // mdmcn_get_gradient_weight, cuh:71-79: (h+1-ah), (ah+1-h) per corner
            const float th = (float)(qd.hl + 1) - h_im;            // top rows
            const float bh = (h_im + 1.f) - (float)(qd.hl + 1);    // bottom rows
            const float lw_ = (float)(qd.wl + 1) - w_im;           // left cols
            const float rw = (w_im + 1.f) - (float)(qd.wl + 1);    // right cols
            const bool topv = qd.hl >= 0, botv = qd.hl + 1 <= H - 1;
            const bool lefv = qd.wl >= 0, rigv = qd.wl + 1 <= W - 1;
            float *sp = so + (long)qd.hl * W + qd.wl;
            if (topv && lefv) atomicAdd(sp, th * lw_ * top);
            if (topv && rigv) atomicAdd(sp + 1, th * rw * top);
            if (botv && lefv) atomicAdd(sp + W, bh * lw_ * top);
            if (botv && rigv) atomicAdd(sp + W + 1, bh * rw * top);
        }
        // pair index inside g_off: NLSPN path skips the centre pair
        const int pr = CENTER_ZERO ? (t < G::REF ? t : t - 1) : t;
        if (first) {
            gob[(long)(2 * pr) * P] = gh;
            gob[(long)(2 * pr + 1) * P] = gw;
            gab[(long)t * P] = ga;
        } else {
            gob[(long)(2 * pr) * P] += gh;
            gob[(long)(2 * pr + 1) * P] += gw;
            gab[(long)t * P] += ga;
        }
    }
}

// ======================================================================================
// Final backward kernel: consumes the last scatter plane (gradient wrt src[0]) and runs the
// backward of the prologue (formulas: SURVEY 3.2, derived from nlspnmodel.py:185-197,262-267).
// g_guidance's offset part already holds the accumulated offset gradients.
// ======================================================================================
// LAYOUT of s_in: 0 = plain [B,H,W] plane, 1 = four phase copies of the 2x2-blocked plane (kernels_v2.cuh),
// 2 = one plane padded by `pad` cells on every side with row pitch `pitch` (kernels_local.cuh)
template <int K, int LAYOUT, bool SAMPLED = false>
__global__ void __launch_bounds__(kBlock)
final_bwd_kernel(const float *__restrict__ guidance, const float *__restrict__ init,
                 const float *__restrict__ dep, const float *__restrict__ conf,
                 const float *__restrict__ s_in, const float *__restrict__ g_aff,
                 const float *__restrict__ g_conf_acc, const float *__restrict__ g_off_ext,
                 const float *__restrict__ g_aff_ext, const float *__restrict__ gamma_ptr, int affinity,
                 unsigned flags, int H, int W, float *__restrict__ g_init, float *__restrict__ g_guidance,
                 float *__restrict__ g_conf, double *__restrict__ g_gamma,
                 const float *__restrict__ conf_raw = nullptr, const float *__restrict__ gy_first = nullptr,
                 const float *__restrict__ aff_norm = nullptr, const float *__restrict__ f_last = nullptr,
                 int pad = 0, int pitch = 0, long pad_plane = 0)
{
    // gy_first / aff_norm / f_last: the gather-form pass A (kernels_gather.cuh) keeps the centre tap and
    // the overflowing taps out of the blocked planes; the consumer of the last planes adds
    // gy_1[p] * aff_ref[p] and the overflow plane here.
    // conf_raw: only for kConfSampled (upstream conf_prop): the raw confidence the prologue sampled;
    // its gradient is SCATTERED into g_conf (zeroed by the host) with the bilinear corner weights.
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    const long b = blockIdx.y;
    double local_gamma = 0.0;
    if (r < P) {
        const long q = b * P + r;
        const bool preserve = (flags & kPreserve) != 0;
        const float d = preserve ? __ldg(dep + q) : 0.f;
        const float m = d > 0.f ? 1.f : 0.f;
        float gs;
        if (LAYOUT == 2) {
            const int hh = r / W, ww = r - hh * W;
            gs = __ldg(s_in + b * pad_plane + (long)(hh + pad) * pitch + (ww + pad));
        } else if (LAYOUT == 1) {   // four phase copies of the 2x2-blocked scatter plane (kernels_v2.cuh)
            const ScatterGeo sg = scatter_geo(H, W);
            const float *sl = s_in + b * sg.image;
            const int hh = r / W, ww = r - hh * W;
            gs = 0.f;
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) gs += sl[scatter_cell(sg, ph >> 1, ph & 1, hh + 1, ww + 1)];
        } else {
            gs = __ldg(s_in + q);
        }
        if (f_last) gs += __ldg(f_last + q);
        if (gy_first) gs += __ldg(gy_first + q) * __ldg(aff_norm + (b * G::KK + G::REF) * P + r);
        float x0 = __ldg(init + q);
        if (preserve) x0 = blend_fix(x0, d);
        const bool clipped = (flags & kAlwaysClip) && x0 < 0.f;
        if (flags & kAlwaysClip) x0 = fmaxf(x0, 0.f);
        float Gx;
        if (conf) {
            const float gc = __ldg(g_conf_acc + q) + x0 * gs;
            g_conf[q] = preserve ? (1.0f - m) * gc : gc;
            Gx = __ldg(conf + q) * gs;
        } else {
            Gx = gs;
        }
        if (clipped) Gx = 0.f;
        g_init[q] = preserve ? (1.0f - m) * Gx : Gx;

        const bool no_off = (flags & kNoOffset) != 0;
        const int aff_ch0 = no_off ? 0 : 2 * G::N;
        const float *gb = guidance + b * (no_off ? 1 : 3) * G::N * P + r;
        float *ggb = g_guidance + b * (no_off ? 1 : 3) * G::N * P + r;
        if (g_off_ext && !no_off) {
            const float *eb = g_off_ext + b * 2 * G::KK * P + r;
#pragma unroll
            for (int n = 0; n < G::N; ++n) {
                const int t = n < G::REF ? n : n + 1;
                ggb[(long)(2 * n) * P] += __ldg(eb + (long)(2 * t) * P);
                ggb[(long)(2 * n + 1) * P] += __ldg(eb + (long)(2 * t + 1) * P);
            }
        }
        const float *gab = g_aff + b * G::KK * P + r;
        const float *eab = g_aff_ext ? g_aff_ext + b * G::KK * P + r : nullptr;
        const bool use_tanh = affinity == kTC || affinity == kTGASS;
        const float gamma = __ldg(gamma_ptr);
        const float g = affinity == kTGASS ? gamma + 1e-8f : gamma;
        // Three IEEE divisions per pixel instead of 4N: the denominators g, s and s^2 are shared by
        // all N affinities, and a gradient tolerates the last-bit difference of x * (1/d) vs x / d
        // (the FORWARD prologue keeps true divisions: its outputs are API-visible values).
        const float inv_g = 1.0f / g;
        constexpr bool sampled = SAMPLED;   // compile-time: the common (fork) path carries none of this
        const int ph = r / W, pw = r - ph * W;
        float a[G::N], th[G::N], Gh[G::N];
        // every load of the affinity part is issued before the first tanhf: its branches otherwise fence the loads
        // into small groups (ncu, round 2: 2.9 TB/s with long-scoreboard 8.9 per issue on the dependent FADDs)
        float raw[G::N], gav[G::KK];
#pragma unroll
        for (int n = 0; n < G::N; ++n) raw[n] = __ldg(gb + (long)(aff_ch0 + n) * P);
#pragma unroll
        for (int t = 0; t < G::KK; ++t) gav[t] = __ldg(gab + (long)t * P);
        if (eab) {
#pragma unroll
            for (int t = 0; t < G::KK; ++t) gav[t] += __ldg(eab + (long)t * P);
        }
        float s0 = 0.f;
#pragma unroll
        for (int n = 0; n < G::N; ++n) {
            const float rr = raw[n];
            th[n] = use_tanh ? tanhf(rr) : 0.f;
            a[n] = use_tanh ? th[n] * inv_g : rr;
            if (sampled) a[n] *= sample_conf<K>(conf_raw + b * P, gb, n, ph, pw, H, W, P, (flags & kLegacy) != 0);
            s0 += fabsf(a[n]);
        }
        s0 += 1e-4f;
        bool clamped = false;
        float s = s0;
        if (affinity == kASS || affinity == kTGASS) {
            clamped = s0 < 1.0f;
            // The clamp `s[s < 1] = 1` (nlspnmodel.py:193-194) is a kink: within a few ulp of 1 the DECISION must be
            // the forward's, and th * (1/g) above is not bit-identical to the prologue's tanhf(r) / g.  Near the
            // kink the sum is therefore re-formed with the prologue's exact expressions and order (soak case 121 of
            // the randomized sweep: one pixel in 170 k with s = 1 - 1 ulp took the other branch).
            if (fabsf(s0 - 1.0f) < 1e-5f) {
                float se = 0.f;
#pragma unroll
                for (int n = 0; n < G::N; ++n) {
                    float v = raw[n];
                    if (use_tanh) v = tanhf(v) / g;
                    if (sampled) v *= sample_conf<K>(conf_raw + b * P, gb, n, ph, pw, H, W, P, (flags & kLegacy) != 0);
                    se += fabsf(v);
                }
                se += 1e-4f;
                clamped = se < 1.0f;
            }
            if (clamped) s = 1.0f;
        }
        const float inv_s = 1.0f / s;
        const float Gref = gav[G::REF];
        float dot = 0.f, gsum = 0.f;
#pragma unroll
        for (int n = 0; n < G::N; ++n) {
            const int t = n < G::REF ? n : n + 1;
            const float gv = gav[t];
            Gh[n] = gv - Gref;
            dot += Gh[n] * a[n];
        }
        const float corr = clamped ? 0.f : dot * inv_s * inv_s;
#pragma unroll
        for (int n = 0; n < G::N; ++n) {
            float da;
            if (affinity == kTC) {
                da = Gh[n];
            } else {
                const float sg = a[n] > 0.f ? 1.f : (a[n] < 0.f ? -1.f : 0.f);
                da = Gh[n] * inv_s - sg * corr;
            }
            if (sampled) {
                // a_n = u_n * c_n with u_n = tanh(r_n)/g and c_n the sampled confidence
                float hc, wc;
                const float cs = sample_conf<K>(conf_raw + b * P, gb, n, ph, pw, H, W, P, (flags & kLegacy) != 0,
                                                &hc, &wc);
                const float un = use_tanh ? th[n] * inv_g : raw[n];
                const float gcs = da * un;            // d loss / d c_n
                if (tap_valid(hc, wc, H, W) && gcs != 0.f) {
                    // grad_input of the 1x1 gather: mdmcn_get_gradient_weight, cuh:56-81
                    const int hl = (int)floorf(hc), wl = (int)floorf(wc);
                    const float tw = (float)(hl + 1) - hc, bw = (hc + 1.f) - (float)(hl + 1);
                    const float lw_ = (float)(wl + 1) - wc, rw = (wc + 1.f) - (float)(wl + 1);
                    float *gp = g_conf + b * P + (long)hl * W + wl;
                    const bool tp = hl >= 0, bt = hl + 1 <= H - 1, lf = wl >= 0, rg = wl + 1 <= W - 1;
                    if (tp && lf) atomicAdd(gp, tw * lw_ * gcs);
                    if (tp && rg) atomicAdd(gp + 1, tw * rw * gcs);
                    if (bt && lf) atomicAdd(gp + W, bw * lw_ * gcs);
                    if (bt && rg) atomicAdd(gp + W + 1, bw * rw * gcs);
                }
                da *= cs;                              // d loss / d u_n
            }
            float dr = da;
            if (use_tanh) {
                dr = da * (1.f - th[n] * th[n]) * inv_g;
                gsum += da * th[n];
            }
            ggb[(long)(aff_ch0 + n) * P] = dr;
        }
        // d a_n / d gamma = -tanh(r_n) / g^2 ; one fp64 conversion per thread (fp64 ALUs are scarce)
        local_gamma = -(double)gsum / ((double)g * (double)g);
    }
    // block reduction of the gamma gradient, one double atomic per block
    if (affinity == kTGASS || affinity == kTC) {
        __shared__ double red[kBlock / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_gamma += __shfl_xor_sync(0xffffffffu, local_gamma, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local_gamma;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int i = 0; i < kBlock / 32; ++i) tot += red[i];
            // spread over kGammaSlots addresses: thousands of same-address fp64 atomics serialise in L2.
            // Deterministic mode: g_gamma is a per-block array, summed in a fixed order by gamma_reduce_det_kernel.
            if (flags & kDeterministic) g_gamma[blockIdx.x + blockIdx.y * gridDim.x] = tot;
            else if (tot != 0.0) atomicAdd(g_gamma + (blockIdx.x + blockIdx.y * gridDim.x) % kGammaSlots, tot);
        }
    }
}

// sums the kGammaSlots partial gamma gradients into the caller's scalar (one warp, fixed order)
__global__ void gamma_reduce_kernel(const double *__restrict__ slots, double *__restrict__ g_gamma)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < kGammaSlots; i += 32) v += slots[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *g_gamma = v;
}

// grad_weight / grad_bias of the single-step operator (modulated_deform_conv_cuda.cu:248,271-272)
template <int K>
__global__ void __launch_bounds__(kBlock)
dcn_wb_grad_kernel(const float *__restrict__ input, const float *__restrict__ offset,
                   const float *__restrict__ mask, const float *__restrict__ gout, int H, int W,
                   float *__restrict__ gw, float *__restrict__ gbias)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    const long b = blockIdx.y;
    float part[G::KK + 1];
#pragma unroll
    for (int t = 0; t <= G::KK; ++t) part[t] = 0.f;
    if (r < P) {
        const int h = r / W, w = r - h * W;
        const float g = __ldg(gout + b * P + r);
        const float *im = input + b * P;
        const float *ob = offset + b * 2 * G::KK * P + r;
        const float *mb = mask + b * G::KK * P + r;
        part[G::KK] = g;
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
            const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
            float v = 0.f;
            if (tap_valid(h_im, w_im, H, W)) v = quad_value(load_quad(im, H, W, h_im, w_im));
            part[t] = g * (v * __ldg(mb + (long)t * P));
        }
    }
    __shared__ float red[kBlock / 32][G::KK + 1];
#pragma unroll
    for (int t = 0; t <= G::KK; ++t) {
        float v = part[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
    }
    __syncthreads();
    if (threadIdx.x <= G::KK) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < kBlock / 32; ++i) tot += red[i][threadIdx.x];
        atomicAdd(threadIdx.x == G::KK ? gbias : gw + threadIdx.x, tot);
    }
}

// debug: integer corners per tap
template <int K>
__global__ void __launch_bounds__(kBlock)
debug_indices_kernel(const float *__restrict__ offset, int H, int W, int32_t *__restrict__ idx)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const float *ob = offset + b * 2 * G::KK * P + r;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
        const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
        int32_t *o = idx + ((b * G::KK + t) * 3) * P + r;
        o[0] = (int32_t)floorf(h_im);
        o[P] = (int32_t)floorf(w_im);
        o[2 * (long)P] = tap_valid(h_im, w_im, H, W) ? 1 : 0;
    }
}

} // namespace nlspn
