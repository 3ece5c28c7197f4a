// common.cuh -- shared device helpers for the NLSPN propagation kernels (sm_100a).
//
// Arithmetic contract (SURVEY 0.4 / 7 hard-part 3): the sampling coordinate is formed as
// (float)(h - pad + i) + offset with the integer part added first and ONE floating add, the
// fractional weights as lh = h_im - floor(h_im), hh = 1 - lh, exactly like the reference
// (modulated_deform_im2col_cuda.cuh:28-35,157-158,178-179).  Re-associating any of this
// breaks the 1e-4 m bound at KITTI width.  No fast-math anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nlspn {

constexpr int kBlock = 256;

template <int K>
struct Geo {
    static constexpr int KK = K * K;
    static constexpr int N = KK - 1;
    static constexpr int REF = N / 2;       // nlspnmodel.py:91
    static constexpr int PAD = (K - 1) / 2; // nlspnmodel.py:117
};

// in-range test of modulated_deform_im2col_cuda.cuh:180 (NaN -> false)
__device__ __forceinline__ bool tap_valid(float h_im, float w_im, int H, int W)
{
    return h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
}

// floor() of a coordinate known to satisfy |x| < 2^22 (every coordinate that passed tap_valid()),
// returning it both as float and as int WITHOUT the F2I / I2F conversions, which run on the
// quarter-rate XU pipe (ncu: 30 % of pass A's stall samples sat on F2I.FLOOR).  x + 1.5*2^23
// rounds x to the nearest integer exactly and leaves that integer in the low mantissa bits;
// one compare turns round-to-nearest into floor.  Exact for all |x| < 2^22.
__device__ __forceinline__ void floor_small(float x, float &xf, int &xi)
{
    const float t = x + 12582912.f;                       // 0x4B400000
    int ri = __float_as_int(t) - 0x4B400000;
    float rf = t - 12582912.f;
    if (rf > x) {
        rf -= 1.f;
        ri -= 1;
    }
    xf = rf;
    xi = ri;
}

// The four corner values of a zero-padded bilinear sample read straight from global memory
// with the per-corner guards of cuh:37-48.  Precondition: tap_valid().
struct Quad {
    float v1, v2, v3, v4; // (hl,wl) (hl,wl+1) (hl+1,wl) (hl+1,wl+1)
    float lh, lw;         // fractional parts
    int hl, wl;
};

__device__ __forceinline__ Quad load_quad(const float *__restrict__ im, int H, int W, float h_im,
                                          float w_im)
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
    q.v1 = (top && lef) ? __ldg(p) : 0.f;
    q.v2 = (top && rig) ? __ldg(p + 1) : 0.f;
    q.v3 = (bot && lef) ? __ldg(p + W) : 0.f;
    q.v4 = (bot && rig) ? __ldg(p + W + 1) : 0.f;
    return q;
}

// cuh:50-52
__device__ __forceinline__ float quad_value(const Quad &q)
{
    const float hh = 1.f - q.lh, hw = 1.f - q.lw;
    const float w1 = hh * hw, w2 = hh * q.lw, w3 = q.lh * hw, w4 = q.lh * q.lw;
    return w1 * q.v1 + w2 * q.v2 + w3 * q.v3 + w4 * q.v4;
}

// input-preserving blend, nlspnmodel.py:344,357 -- literal, so that fixed pixels are exact
__device__ __forceinline__ float blend_fix(float v, float dep)
{
    const float m = dep > 0.f ? 1.f : 0.f;
    return (1.f - m) * v + m * dep;
}

// Geometry of the blocked scatter planes of one image (backward pass A, kernels_v2.cuh).
// The gradient plane is kept in FOUR copies made of 2x2-pixel blocks, one per block phase
// (sy, sx) in {0,1}^2, so that any bilinear footprint (rows Y, Y+1; cols X, X+1 in padded
// coordinates Y = y+1, X = x+1) is exactly one aligned 16-byte block of the copy with
// sy = Y&1, sx = X&1.  The two column phases are interleaved block by block: the footprints at
// X = 2m and X = 2m+1 (neighbouring lanes when the offsets are spatially smooth) are the two
// halves of ONE 32-byte sector, which halves the sectors a warp-wide vector RED ships when the
// offsets are coherent (iid offsets: no change).
struct ScatterGeo {
    int Hb, Wb;          // block rows / block-pair columns
    long plane;          // floats per row-phase plane = Hb * Wb * 8
    long image;          // floats per image = 2 * plane
};

__host__ __device__ inline ScatterGeo scatter_geo(int H, int W)
{
    ScatterGeo g;
    g.Hb = H / 2 + 2;
    g.Wb = W / 2 + 2;
    g.plane = (long)g.Hb * g.Wb * 8;
    g.image = 2 * g.plane;
    return g;
}

// first float of the block that holds padded cell (Y, X) in phase (sy, sx); (X + sx) / 2 - sx >= 0
// for every footprint (sx = X&1) and for every reader (X >= 1)
__device__ __forceinline__ long scatter_block(const ScatterGeo &g, int sy, int sx, int Y, int X)
{
    const int by = (Y + sy) >> 1;
    const int bxp = ((X + sx) >> 1) - sx;
    return (long)sy * g.plane + (((long)by * g.Wb + bxp) * 2 + sx) * 4;
}

// address of padded cell (Y, X) = (y+1, x+1) in phase (sy, sx)
__device__ __forceinline__ long scatter_cell(const ScatterGeo &g, int sy, int sx, int Y, int X)
{
    return scatter_block(g, sy, sx, Y, X) + ((Y + sy) & 1) * 2 + ((X + sx) & 1);
}

} // namespace nlspn
