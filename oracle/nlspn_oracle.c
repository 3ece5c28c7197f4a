/*
 * nlspn_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * NLSPN propagation path.  Nothing the product ships may call into this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg load it.
 *
 * What it restates (reference paths are relative to /root/reference):
 *   sampling, gradients, border rules  src/model/deformconv/src/cuda/modulated_deform_im2col_cuda.cuh
 *   host op (C=1 case)                 src/model/deformconv/src/cuda/modulated_deform_conv_cuda.cu:19-121,124-280
 *   module logic                       src/model/nlspnmodel.py:179-201,252-269,323-377
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every entry point
 * against the .npz fixtures in tests/golden, which oracle/gen_golden.py produced by running the
 * reference's own unmodified Python (NLSPNModel.forward and its helper methods)
 * on CPU with torchvision.ops.deform_conv2d standing in for the CUDA-only DCN
 * extension (the stand-in BASELINE.json's north_star prescribes).
 *
 * The file is compiled twice: -DREAL=float -DSUF=f32 and -DREAL=double -DSUF=f64.
 * Compile with -ffp-contract=off so fp32 results do not depend on FMA fusion.
 *
 * Domain: C_in = C_out = 1, groups = deformable_groups = 1, stride 1, dilation 1,
 * pad = (K-1)/2 -- exactly what nlspnmodel.py:107-121,205-208 passes.
 * Layouts are the reference's: NCHW contiguous; offset channel 2*tap = dh,
 * 2*tap+1 = dw, tap = i*K + j (cuh:171-172).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#define SUF f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

#define AFF_AS 0
#define AFF_ASS 1
#define AFF_TC 2
#define AFF_TGASS 3

/* ---- cuh:24-54 : zero-padded bilinear sample ------------------------------------ */
static REAL sample_bilinear(const REAL *im, int H, int W, REAL h, REAL w)
{
    int hl = (int)floor(h), wl = (int)floor(w);
    int hh_i = hl + 1, wh_i = wl + 1;
    REAL lh = h - hl, lw = w - wl;
    REAL hh = 1 - lh, hw = 1 - lw;
    REAL v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (hl >= 0 && wl >= 0) v1 = im[hl * W + wl];
    if (hl >= 0 && wh_i <= W - 1) v2 = im[hl * W + wh_i];
    if (hh_i <= H - 1 && wl >= 0) v3 = im[hh_i * W + wl];
    if (hh_i <= H - 1 && wh_i <= W - 1) v4 = im[hh_i * W + wh_i];
    REAL w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* ---- cuh:56-81 : weight with which sample (ah,aw) touched integer pixel (h,w) ---- */
static REAL corner_weight(REAL ah, REAL aw, int h, int w, int H, int W)
{
    if (ah <= -1 || ah >= H || aw <= -1 || aw >= W) return 0;
    int hl = (int)floor(ah), wl = (int)floor(aw);
    int hh_i = hl + 1, wh_i = wl + 1;
    REAL r = 0;
    if (h == hl && w == wl) r = (h + 1 - ah) * (w + 1 - aw);
    if (h == hl && w == wh_i) r = (h + 1 - ah) * (aw + 1 - w);
    if (h == hh_i && w == wl) r = (ah + 1 - h) * (w + 1 - aw);
    if (h == hh_i && w == wh_i) r = (ah + 1 - h) * (aw + 1 - w);
    return r;
}

/* ---- cuh:83-125 : d(bilinear)/d(coordinate); dir 0 = h, 1 = w --------------------- */
static REAL coord_weight(REAL ah, REAL aw, int H, int W, const REAL *im, int dir)
{
    if (ah <= -1 || ah >= H || aw <= -1 || aw >= W) return 0;
    int hl = (int)floor(ah), wl = (int)floor(aw);
    int hh_i = hl + 1, wh_i = wl + 1;
    REAL r = 0;
    if (dir == 0) {
        if (hl >= 0 && wl >= 0) r += -1 * (wl + 1 - aw) * im[hl * W + wl];
        if (hl >= 0 && wh_i <= W - 1) r += -1 * (aw - wl) * im[hl * W + wh_i];
        if (hh_i <= H - 1 && wl >= 0) r += (wl + 1 - aw) * im[hh_i * W + wl];
        if (hh_i <= H - 1 && wh_i <= W - 1) r += (aw - wl) * im[hh_i * W + wh_i];
    } else {
        if (hl >= 0 && wl >= 0) r += -1 * (hl + 1 - ah) * im[hl * W + wl];
        if (hl >= 0 && wh_i <= W - 1) r += (hl + 1 - ah) * im[hl * W + wh_i];
        if (hh_i <= H - 1 && wl >= 0) r += -1 * (ah - hl) * im[hh_i * W + wl];
        if (hh_i <= H - 1 && wh_i <= W - 1) r += (ah - hl) * im[hh_i * W + wh_i];
    }
    return r;
}

/* ---- one modulated deformable step, forward -------------------------------------
 * cuh:127-194 (im2col, validity test at :180) followed by the GEMV + bias of
 * modulated_deform_conv_cuda.cu:112.  wgt has K*K entries, bias one. */
void FN(dcn_step_fwd)(const REAL *x, const REAL *off, const REAL *msk, const REAL *wgt,
                      const REAL *bias, int B, int H, int W, int K, REAL *out)
{
    const int pad = (K - 1) / 2, KK = K * K;
    const long P = (long)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h) {
            const REAL *im = x + b * P;
            const REAL *ob = off + (long)b * 2 * KK * P;
            const REAL *mb = msk + (long)b * KK * P;
            for (int w = 0; w < W; ++w) {
                REAL acc = bias ? bias[0] : 0;
                for (int i = 0; i < K; ++i)
                    for (int j = 0; j < K; ++j) {
                        const int t = i * K + j;
                        const REAL oh = ob[(2 * t) * P + h * W + w];
                        const REAL ow = ob[(2 * t + 1) * P + h * W + w];
                        const REAL m = mb[t * P + h * W + w];
                        /* integer part first, one floating add: cuh:178-179 */
                        const REAL hi = (h - pad + i) + oh;
                        const REAL wi = (w - pad + j) + ow;
                        REAL v = 0;
                        if (hi > -1 && wi > -1 && hi < H && wi < W)
                            v = sample_bilinear(im, H, W, hi, wi);
                        acc += (wgt ? wgt[t] : 1) * (v * m);
                    }
                out[b * P + h * W + w] = acc;
            }
        }
}

/* ---- one modulated deformable step, backward ------------------------------------
 * columns = weight^T * grad_out (cu:221); col2im_coord (cuh:256-328) -> grad_offset,
 * grad_mask; col2im (cuh:196-254) -> grad_input; second im2col + addmm/addmv
 * (cu:248,271-272) -> grad_weight, grad_bias (optional, may be NULL).
 * All outputs are OVERWRITTEN except when accumulate != 0 (then += for
 * grad_offset/grad_mask; grad_input is always overwritten). */
void FN(dcn_step_bwd)(const REAL *x, const REAL *off, const REAL *msk, const REAL *wgt,
                      const REAL *gout, int B, int H, int W, int K, int accumulate,
                      REAL *gin, REAL *goff, REAL *gmsk, REAL *gwgt, REAL *gbias)
{
    const int pad = (K - 1) / 2, KK = K * K;
    const long P = (long)H * W;
    memset(gin, 0, sizeof(REAL) * (size_t)B * P);
    if (gwgt) memset(gwgt, 0, sizeof(REAL) * KK);
    if (gbias) gbias[0] = 0;
    /* scatter is serial inside one image (the reference uses atomics: order-free) */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const REAL *im = x + b * P;
        const REAL *ob = off + (long)b * 2 * KK * P;
        const REAL *mb = msk + (long)b * KK * P;
        const REAL *g = gout + b * P;
        REAL *gi = gin + b * P;
        REAL *gob = goff + (long)b * 2 * KK * P;
        REAL *gmb = gmsk + (long)b * KK * P;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int t = 0; t < KK; ++t) {
                    const int i = t / K, j = t % K;
                    const long q = (long)h * W + w;
                    const REAL oh = ob[(2 * t) * P + q], ow = ob[(2 * t + 1) * P + q];
                    const REAL m = mb[t * P + q];
                    const REAL col = (wgt ? wgt[t] : 1) * g[q];
                    const REAL hi = (h - pad + i) + oh;
                    const REAL wi = (w - pad + j) + ow;
                    /* --- col2im_coord (cuh:308-326) --- */
                    REAL ih = hi, iw = wi, mval = 0;
                    if (ih <= -1 || iw <= -1 || ih >= H || iw >= W) {
                        ih = iw = -2;
                    } else {
                        mval = col * sample_bilinear(im, H, W, ih, iw);
                    }
                    const REAL dh = coord_weight(ih, iw, H, W, im, 0) * col * m;
                    const REAL dw = coord_weight(ih, iw, H, W, im, 1) * col * m;
                    if (accumulate) {
                        gob[(2 * t) * P + q] += dh;
                        gob[(2 * t + 1) * P + q] += dw;
                        gmb[t * P + q] += mval;
                    } else {
                        gob[(2 * t) * P + q] = dh;
                        gob[(2 * t + 1) * P + q] = dw;
                        gmb[t * P + q] = mval;
                    }
                    /* --- col2im (cuh:229-252): truncation + 5x5 search --- */
                    const REAL top = col * m;
                    const int ch = (int)hi, cw = (int)wi;
                    for (int dy = -2; dy <= 2; ++dy)
                        for (int dx = -2; dx <= 2; ++dx) {
                            const int yy = ch + dy, xx = cw + dx;
                            if (yy >= 0 && yy < H && xx >= 0 && xx < W &&
                                fabs(hi - yy) < 1 && fabs(wi - xx) < 1) {
                                gi[yy * W + xx] += corner_weight(hi, wi, yy, xx, H, W) * top;
                            }
                        }
                }
    }
    if (gwgt || gbias) { /* serial reductions, double accumulators are NOT used: REAL */
        for (int b = 0; b < B; ++b)
            for (long q = 0; q < P; ++q) {
                const int h = (int)(q / W), w = (int)(q % W);
                const REAL gq = gout[b * P + q];
                if (gbias) gbias[0] += gq;
                if (gwgt)
                    for (int t = 0; t < KK; ++t) {
                        const int i = t / K, j = t % K;
                        const REAL oh = off[((long)b * 2 * KK + 2 * t) * P + q];
                        const REAL ow = off[((long)b * 2 * KK + 2 * t + 1) * P + q];
                        const REAL hi = (h - pad + i) + oh, wi = (w - pad + j) + ow;
                        REAL v = 0;
                        if (hi > -1 && wi > -1 && hi < H && wi < W)
                            v = sample_bilinear(x + b * P, H, W, hi, wi);
                        gwgt[t] += gq * (v * msk[((long)b * KK + t) * P + q]);
                    }
            }
    }
}

/* ---- debug: integer corners per tap (exact-index test) ---------------------------
 * idx[b, tap, 0, h, w] = floor(h_im), idx[b, tap, 1, h, w] = floor(w_im),
 * idx[b, tap, 2, h, w] = 1 if the tap passes the validity test of cuh:180 else 0. */
void FN(dcn_debug_indices)(const REAL *off, int B, int H, int W, int K, int *idx)
{
    const int pad = (K - 1) / 2, KK = K * K;
    const long P = (long)H * W;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < KK; ++t)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const long q = (long)h * W + w;
                    const REAL oh = off[((long)b * 2 * KK + 2 * t) * P + q];
                    const REAL ow = off[((long)b * 2 * KK + 2 * t + 1) * P + q];
                    const REAL hi = (h - pad + t / K) + oh, wi = (w - pad + t % K) + ow;
                    int *o = idx + (((long)b * KK + t) * 3) * P + q;
                    o[0] = (int)floor(hi);
                    o[P] = (int)floor(wi);
                    o[2 * P] = (hi > -1 && wi > -1 && hi < H && wi < W) ? 1 : 0;
                }
}

/* ---- prologue: nlspnmodel.py:252-259 (_off_insert), :179-201 (_affinity_normalization),
 *      :261-269 (_aff_insert), :328-334 (mask_fix, confidence fix-up) -----------------
 * guidance [B,3N,H,W] = 2N offset channels then N raw affinities (layout of :303-305).
 * Outputs offset [B,2K^2,H,W], aff [B,K^2,H,W], conf_out [B,1,H,W] (if conf given). */
void FN(nlspn_prologue_fwd)(const REAL *guidance, const REAL *conf, const REAL *dep,
                            REAL gamma, int affinity, int preserve, int B, int H, int W,
                            int K, REAL *offset, REAL *aff, REAL *conf_out)
{
    const int KK = K * K, N = KK - 1, ref = N / 2;
    const long P = (long)H * W;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const REAL *gb = guidance + (long)b * 3 * N * P;
        REAL *ob = offset + (long)b * 2 * KK * P;
        REAL *ab = aff + (long)b * KK * P;
        for (long q = 0; q < P; ++q) {
            /* _off_insert: zero pair at tap `ref` */
            for (int t = 0; t < KK; ++t) {
                if (t == ref) {
                    ob[(2 * t) * P + q] = 0;
                    ob[(2 * t + 1) * P + q] = 0;
                } else {
                    const int n = t < ref ? t : t - 1;
                    ob[(2 * t) * P + q] = gb[(2 * n) * P + q];
                    ob[(2 * t + 1) * P + q] = gb[(2 * n + 1) * P + q];
                }
            }
            /* _affinity_normalization */
            REAL a[128];
            REAL abs_sum = 0;
            for (int n = 0; n < N; ++n) {
                REAL r = gb[(2 * N + n) * P + q];
                if (affinity == AFF_TC) r = (REAL)tanh(r) / gamma;
                else if (affinity == AFF_TGASS) r = (REAL)tanh(r) / (gamma + (REAL)1e-8);
                a[n] = r;
                abs_sum += (REAL)fabs(r);
            }
            abs_sum += (REAL)1e-4;
            if ((affinity == AFF_ASS || affinity == AFF_TGASS) && abs_sum < (REAL)1.0)
                abs_sum = (REAL)1.0;
            REAL sum = 0;
            for (int n = 0; n < N; ++n) {
                if (affinity != AFF_TC) a[n] = a[n] / abs_sum;
                sum += a[n];
            }
            /* _aff_insert */
            for (int t = 0; t < KK; ++t) {
                if (t == ref) ab[t * P + q] = (REAL)1.0 - sum;
                else ab[t * P + q] = a[t < ref ? t : t - 1];
            }
            if (conf && conf_out) {
                const long bq = b * P + q;
                if (preserve && dep) {
                    const REAL m = dep[bq] > 0 ? (REAL)1 : (REAL)0;
                    conf_out[bq] = ((REAL)1.0 - m) * conf[bq] + m;
                } else {
                    conf_out[bq] = conf[bq];
                }
            }
        }
    }
}

/* ---- the T-iteration loop: nlspnmodel.py:336-363 -------------------------------------
 * offset/aff are the POST-insert tensors, conf the POST-fix-up confidence (or NULL).
 * list_feat [T,B,1,H,W] receives every iteration's state (list_pred, :363);
 * feat_result is list_feat[T-1] (un-clamped unless always_clip).
 * scratch: 2*B*H*W REALs. */
void FN(nlspn_propagate_fwd)(const REAL *feat_init, const REAL *offset, const REAL *aff,
                             const REAL *conf, const REAL *dep, int preserve, int always_clip,
                             int B, int H, int W, int K, int T, REAL *list_feat, REAL *scratch)
{
    const long P = (long)H * W, BP = (long)B * P;
    REAL *cur = scratch, *src = scratch + BP;
    for (long q = 0; q < BP; ++q) {
        REAL v = feat_init[q];
        if (preserve) {
            const REAL m = dep[q] > 0 ? (REAL)1 : (REAL)0;
            v = ((REAL)1.0 - m) * v + m * dep[q];           /* :344 */
        }
        if (always_clip && v < 0) v = 0;                    /* :348 */
        cur[q] = v;
    }
    for (int t = 0; t < T; ++t) {
        for (long q = 0; q < BP; ++q) src[q] = conf ? cur[q] * conf[q] : cur[q];   /* :351 */
        REAL *dst = list_feat + (long)t * BP;
        FN(dcn_step_fwd)(src, offset, aff, NULL, NULL, B, H, W, K, dst);           /* :205-208 */
        for (long q = 0; q < BP; ++q) {
            REAL v = dst[q];
            if (preserve) {
                const REAL m = dep[q] > 0 ? (REAL)1 : (REAL)0;
                v = ((REAL)1.0 - m) * v + m * dep[q];       /* :357 */
            }
            if (always_clip && v < 0) v = 0;                /* :361 */
            dst[q] = v;
            cur[q] = v;
        }
    }
}

/* ---- backward of the loop (what autograd does with :336-363) -------------
 * g_list [T,B,1,H,W]: upstream gradient for every list_feat[t] (the gradient wrt
 * feat_result is simply part of g_list[T-1]).  always_clip is not differentiated here.
 * Outputs (overwritten): g_init [B,1,H,W], g_offset [B,2K^2,H,W] (centre pair included,
 * as the reference's grad_offset has it), g_aff [B,K^2,H,W], g_conf [B,1,H,W] = gradient
 * wrt the POST-fix-up confidence (may be NULL when conf is NULL).
 * scratch: 3*B*H*W REALs. */
void FN(nlspn_propagate_bwd)(const REAL *feat_init, const REAL *offset, const REAL *aff,
                             const REAL *conf, const REAL *dep, int preserve,
                             const REAL *list_feat, const REAL *g_list, int B, int H, int W,
                             int K, int T, REAL *g_init, REAL *g_offset, REAL *g_aff,
                             REAL *g_conf, REAL *scratch)
{
    const int KK = K * K;
    const long P = (long)H * W, BP = (long)B * P;
    REAL *gx = scratch, *src = scratch + BP, *gs = scratch + 2 * BP;
    memset(gx, 0, sizeof(REAL) * BP);
    memset(g_offset, 0, sizeof(REAL) * (size_t)B * 2 * KK * P);
    memset(g_aff, 0, sizeof(REAL) * (size_t)B * KK * P);
    if (g_conf) memset(g_conf, 0, sizeof(REAL) * BP);
    for (int t = T - 1; t >= 0; --t) {
        /* state that entered iteration t (x_{t-1}); for t == 0 the blended init */
        for (long q = 0; q < BP; ++q) {
            REAL xprev;
            if (t > 0) xprev = list_feat[(long)(t - 1) * BP + q];
            else {
                xprev = feat_init[q];
                if (preserve) {
                    const REAL m = dep[q] > 0 ? (REAL)1 : (REAL)0;
                    xprev = ((REAL)1.0 - m) * xprev + m * dep[q];
                }
            }
            src[q] = conf ? xprev * conf[q] : xprev;
            REAL g = gx[q] + g_list[(long)t * BP + q];
            if (preserve) g = ((REAL)1.0 - (dep[q] > 0 ? (REAL)1 : (REAL)0)) * g;   /* d(:357) */
            gx[q] = g;
        }
        FN(dcn_step_bwd)(src, offset, aff, NULL, gx, B, H, W, K, 1, gs, g_offset, g_aff, NULL, NULL);
        for (long q = 0; q < BP; ++q) {
            if (conf) {
                REAL xp;
                if (t > 0) xp = list_feat[(long)(t - 1) * BP + q];
                else {
                    xp = feat_init[q];
                    if (preserve) {
                        const REAL m = dep[q] > 0 ? (REAL)1 : (REAL)0;
                        xp = ((REAL)1.0 - m) * xp + m * dep[q];
                    }
                }
                if (g_conf) g_conf[q] += xp * gs[q];         /* d(x*c)/dc */
                gx[q] = conf[q] * gs[q];                      /* d(x*c)/dx */
            } else {
                gx[q] = gs[q];
            }
        }
    }
    for (long q = 0; q < BP; ++q) {
        REAL g = gx[q];
        if (preserve) g = ((REAL)1.0 - (dep[q] > 0 ? (REAL)1 : (REAL)0)) * g;       /* d(:344) */
        g_init[q] = g;
    }
}

/* ---- prologue backward (formulas of SURVEY 3.2, derived from nlspnmodel.py:185-197,262-267)
 * g_offset [B,2K^2,H,W], g_aff [B,K^2,H,W], g_conf_fixed [B,1,H,W] (or NULL) are the
 * gradients wrt the prologue's OUTPUTS; produces g_guidance [B,3N,H,W], g_conf [B,1,H,W]
 * (wrt the RAW confidence) and returns d(loss)/d(gamma) in *g_gamma (double-accumulated
 * here only to make the scalar reduction order-independent). */
void FN(nlspn_prologue_bwd)(const REAL *guidance, const REAL *dep, REAL gamma, int affinity,
                            int preserve, const REAL *g_offset, const REAL *g_aff,
                            const REAL *g_conf_fixed, int B, int H, int W, int K,
                            REAL *g_guidance, REAL *g_conf, double *g_gamma)
{
    const int KK = K * K, N = KK - 1, ref = N / 2;
    const long P = (long)H * W;
    double gg = 0;
    for (int b = 0; b < B; ++b) {
        const REAL *gb = guidance + (long)b * 3 * N * P;
        REAL *ggb = g_guidance + (long)b * 3 * N * P;
        const REAL *gob = g_offset + (long)b * 2 * KK * P;
        const REAL *gab = g_aff + (long)b * KK * P;
        for (long q = 0; q < P; ++q) {
            for (int n = 0; n < N; ++n) {
                const int t = n < ref ? n : n + 1;
                ggb[(2 * n) * P + q] = gob[(2 * t) * P + q];
                ggb[(2 * n + 1) * P + q] = gob[(2 * t + 1) * P + q];
            }
            REAL a[128], th[128], Gh[128];
            const REAL g = (affinity == AFF_TGASS) ? gamma + (REAL)1e-8 : gamma;
            REAL s0 = 0;
            for (int n = 0; n < N; ++n) {
                const REAL r = gb[(2 * N + n) * P + q];
                if (affinity == AFF_TC || affinity == AFF_TGASS) {
                    th[n] = (REAL)tanh(r);
                    a[n] = th[n] / g;
                } else {
                    th[n] = 0;
                    a[n] = r;
                }
                s0 += (REAL)fabs(a[n]);
            }
            s0 += (REAL)1e-4;
            int clamped = 0;
            REAL s = s0;
            if ((affinity == AFF_ASS || affinity == AFF_TGASS) && s0 < (REAL)1.0) {
                s = (REAL)1.0;
                clamped = 1;
            }
            const REAL Gref = gab[ref * P + q];
            REAL dot = 0;
            for (int n = 0; n < N; ++n) {
                const int t = n < ref ? n : n + 1;
                Gh[n] = gab[t * P + q] - Gref;
                dot += Gh[n] * a[n];
            }
            for (int n = 0; n < N; ++n) {
                REAL da;
                if (affinity == AFF_TC) da = Gh[n];
                else {
                    da = Gh[n] / s;
                    if (!clamped) {
                        const REAL sg = a[n] > 0 ? (REAL)1 : (a[n] < 0 ? (REAL)-1 : (REAL)0);
                        da -= sg * dot / (s * s);
                    }
                }
                REAL dr = da;
                if (affinity == AFF_TC || affinity == AFF_TGASS) {
                    dr = da * ((REAL)1 - th[n] * th[n]) / g;
                    gg += -(double)da * (double)th[n] / ((double)g * (double)g);
                }
                ggb[(2 * N + n) * P + q] = dr;
            }
            if (g_conf && g_conf_fixed) {
                const long bq = b * P + q;
                REAL m = 0;
                if (preserve && dep) m = dep[bq] > 0 ? (REAL)1 : (REAL)0;
                g_conf[bq] = ((REAL)1.0 - m) * g_conf_fixed[bq];
            }
        }
    }
    *g_gamma = gg;
}
