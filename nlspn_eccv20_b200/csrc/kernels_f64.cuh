// kernels_f64.cuh -- double-precision instantiation of the single-step operator (boundary B1).
//
// The reference dispatches its modulated deformable convolution over float AND double
// (AT_DISPATCH_FLOATING_TYPES, modulated_deform_conv_cuda.cu:93,224), and its own tests run
// gradcheck in double (deformconv/test.py).  The fp32 kernels are the product path; these two
// kernels exist so that a double call on the drop-in `DCN` module works and so that
// torch.autograd.gradcheck can validate the derivative formulas on the GPU.  One thread per
// pixel, all K*K taps deformable, weight and bias applied; same boundary rules as the fp32
// kernels (cuh:37-48 corner guards, :180 validity, :88-92 zero coordinate gradient outside).
// Not tuned: double atomics, direct gathers.
#pragma once
#include "common.cuh"

namespace nlspn {

struct QuadD {
    double v1, v2, v3, v4, lh, lw;
    int hl, wl;
};

__device__ __forceinline__ bool tap_valid_d(double h_im, double w_im, int H, int W)
{
    return h_im > -1.0 && w_im > -1.0 && h_im < (double)H && w_im < (double)W;
}

__device__ __forceinline__ QuadD load_quad_d(const double *__restrict__ im, int H, int W, double h_im, double w_im)
{
    QuadD q;
    const double hf = floor(h_im), wf = floor(w_im);
    q.hl = (int)hf;
    q.wl = (int)wf;
    q.lh = h_im - hf;
    q.lw = w_im - wf;
    const bool top = q.hl >= 0, bot = q.hl + 1 <= H - 1, lef = q.wl >= 0, rig = q.wl + 1 <= W - 1;
    const double *p = im + (long)q.hl * W + q.wl;
    q.v1 = (top && lef) ? p[0] : 0.0;
    q.v2 = (top && rig) ? p[1] : 0.0;
    q.v3 = (bot && lef) ? p[W] : 0.0;
    q.v4 = (bot && rig) ? p[W + 1] : 0.0;
    return q;
}

__device__ __forceinline__ double quad_value_d(const QuadD &q)
{
    const double hh = 1.0 - q.lh, hw = 1.0 - q.lw;
    return (hh * hw) * q.v1 + (hh * q.lw) * q.v2 + (q.lh * hw) * q.v3 + (q.lh * q.lw) * q.v4;
}

// out = bias + sum_t weight[t] * mask[t] * bilinear(input; p_t)     (cuh:127-194 + cu:92-118)
template <int K>
__global__ void __launch_bounds__(kBlock)
dcn_fwd_f64_kernel(const double *__restrict__ input, const double *__restrict__ offset,
                   const double *__restrict__ mask, const double *__restrict__ weight,
                   const double *__restrict__ bias, int H, int W, double *__restrict__ out)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const double *im = input + b * P;
    const double *ob = offset + b * 2 * G::KK * P + r;
    const double *mb = mask + b * G::KK * P + r;
    double acc = bias[0];
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        const double h_im = (double)(h - G::PAD + t / K) + ob[(long)(2 * t) * P];
        const double w_im = (double)(w - G::PAD + t % K) + ob[(long)(2 * t + 1) * P];
        double v = 0.0;
        if (tap_valid_d(h_im, w_im, H, W)) v = quad_value_d(load_quad_d(im, H, W, h_im, w_im));
        acc += weight[t] * (v * mb[(long)t * P]);
    }
    out[b * P + r] = acc;
}

// grad_input (scatter, cuh:196-254), grad_offset (cuh:256-328), grad_mask (cuh:314-315),
// grad_weight / grad_bias (cu:236-262); outputs pre-zeroed by the host where they are accumulated.
template <int K>
__global__ void __launch_bounds__(kBlock)
dcn_bwd_f64_kernel(const double *__restrict__ input, const double *__restrict__ offset,
                   const double *__restrict__ mask, const double *__restrict__ weight,
                   const double *__restrict__ gout, int H, int W, double *__restrict__ g_input,
                   double *__restrict__ g_offset, double *__restrict__ g_mask,
                   double *__restrict__ g_weight, double *__restrict__ g_bias)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    const long b = blockIdx.y;
    double part[G::KK + 1];
#pragma unroll
    for (int t = 0; t <= G::KK; ++t) part[t] = 0.0;
    if (r < P) {
        const int h = r / W, w = r - h * W;
        const double g = gout[b * P + r];
        const double *im = input + b * P;
        double *gi = g_input + b * P;
        const double *ob = offset + b * 2 * G::KK * P + r;
        const double *mb = mask + b * G::KK * P + r;
        double *gob = g_offset + b * 2 * G::KK * P + r;
        double *gmb = g_mask + b * G::KK * P + r;
        part[G::KK] = g;
#pragma unroll
        for (int t = 0; t < G::KK; ++t) {
            const double h_im = (double)(h - G::PAD + t / K) + ob[(long)(2 * t) * P];
            const double w_im = (double)(w - G::PAD + t % K) + ob[(long)(2 * t + 1) * P];
            const double a = mb[(long)t * P];
            const double col = weight[t] * g;
            double ga = 0.0, gh = 0.0, gw = 0.0;
            if (tap_valid_d(h_im, w_im, H, W)) {
                const QuadD qd = load_quad_d(im, H, W, h_im, w_im);
                const double val = quad_value_d(qd);
                ga = col * val;
                part[t] = g * (val * a);
                const double top = col * a;
                const double wl1 = (double)(qd.wl + 1) - w_im, wl0 = w_im - (double)qd.wl;
                const double hl1 = (double)(qd.hl + 1) - h_im, hl0 = h_im - (double)qd.hl;
                gh = (-1.0 * wl1 * qd.v1 + -1.0 * wl0 * qd.v2 + wl1 * qd.v3 + wl0 * qd.v4) * top;
                gw = (-1.0 * hl1 * qd.v1 + hl1 * qd.v2 + -1.0 * hl0 * qd.v3 + hl0 * qd.v4) * top;
                const double th = (double)(qd.hl + 1) - h_im, bh = (h_im + 1.0) - (double)(qd.hl + 1);
                const double lw_ = (double)(qd.wl + 1) - w_im, rw = (w_im + 1.0) - (double)(qd.wl + 1);
                const bool topv = qd.hl >= 0, botv = qd.hl + 1 <= H - 1, lefv = qd.wl >= 0, rigv = qd.wl + 1 <= W - 1;
                double *sp = gi + (long)qd.hl * W + qd.wl;
                if (topv && lefv) atomicAdd(sp, th * lw_ * top);
                if (topv && rigv) atomicAdd(sp + 1, th * rw * top);
                if (botv && lefv) atomicAdd(sp + W, bh * lw_ * top);
                if (botv && rigv) atomicAdd(sp + W + 1, bh * rw * top);
            }
            gob[(long)(2 * t) * P] = gh;
            gob[(long)(2 * t + 1) * P] = gw;
            gmb[(long)t * P] = ga;
        }
    }
    __shared__ double red[kBlock / 32][G::KK + 1];
#pragma unroll
    for (int t = 0; t <= G::KK; ++t) {
        double v = part[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
    }
    __syncthreads();
    if (threadIdx.x <= G::KK) {
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < kBlock / 32; ++i) tot += red[i][threadIdx.x];
        atomicAdd(threadIdx.x == G::KK ? g_bias : g_weight + threadIdx.x, tot);
    }
}

} // namespace nlspn
