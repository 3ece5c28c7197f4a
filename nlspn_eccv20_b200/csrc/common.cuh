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

#include <mutex>
#include <utility>
#include <vector>

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

// float part only (same precondition)
__device__ __forceinline__ float floor_small_f(float x)
{
    const float t = x + 12582912.f;
    float rf = t - 12582912.f;
    if (rf > x) rf -= 1.f;
    return rf;
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

// always_clip (nlspnmodel.py:346-348,359-361): torch.clamp(x, min=0) passes the gradient where x >= 0, i.e. also at
// x == 0 exactly.  The forward therefore stores a clipped value as -0.0f and everything else with a + sign on
// zero; the backward recognises "was clipped" by that bit pattern alone (no extra plane).  -0.0f compares equal
// to 0 and multiplies/adds like 0, so the forward arithmetic is unchanged.
__device__ __forceinline__ float clip_keep_sign(float v)
{
    return v < 0.f ? __int_as_float((int)0x80000000u) : v + 0.f;
}
__device__ __forceinline__ bool was_clipped(float stored) { return __float_as_uint(stored) == 0x80000000u; }

// geometry of the blocked scatter planes of one image
struct ScatterGeo {
    int Hb, Wb;          // blocks per column / row
    long plane;          // floats per phase plane (multiple of 4)
    long image;          // floats per image = 4 * plane
};

__host__ __device__ inline ScatterGeo scatter_geo(int H, int W)
{
    ScatterGeo g;
    g.Hb = H / 2 + 2;
    g.Wb = W / 2 + 2;
    g.plane = (long)g.Hb * g.Wb * 4;
    g.image = 4 * g.plane;
    return g;
}

// address of padded cell (Y, X) = (y+1, x+1) in phase plane (sy, sx)
__device__ __forceinline__ long scatter_cell(const ScatterGeo &g, int sy, int sx, int Y, int X)
{
    const int by = (Y + sy) >> 1, ly = (Y + sy) & 1;
    const int bx = (X + sx) >> 1, lx = (X + sx) & 1;
    return (long)(sy * 2 + sx) * g.plane + ((long)by * g.Wb + bx) * 4 + ly * 2 + lx;
}

// Host: opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE (per context), so it
// is remembered per (kernel, device): a process that drives several GPUs (nn.DataParallel threads) must set it on each.
inline cudaError_t ensure_dynamic_smem(const void *kernel, int bytes)
{
    static std::mutex mu;
    static std::vector<std::pair<const void *, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    for (const auto &d : done)
        if (d.first == kernel && d.second == dev) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.emplace_back(kernel, dev);
    return e;
}

} // namespace nlspn
