// kernels_det.cuh -- DETERMINISTIC pass A of the backward (NLSPN_FLAG_DETERMINISTIC): bit-identical gradients from
// run to run, any K.
//
// Why the default forms are not: the reference's col2im scatters with atomicAdd (cuh:229-252) and admits that two
// calls differ (deformconv/test.py:627-631); our RED form (kernels_v2.cuh), the tile-local transpose (kernels_local.cuh:
// footprint ranks come from shared-memory atomics, overlapping regions meet in a TMA reduction) and the tabulated
// gather (kernels_gather.cuh: slot assignment by ATOMG) all sum in an order that depends on scheduling.
//
// Here the transposed operator is tabulated EXACTLY, as a CSR matrix over destination cells, once per backward
// call:
//     det_count_kernel   every in-image corner of every valid tap counts one entry for its cell (integer atomics:
//                        the COUNTS do not depend on order)
//     det_scan_*         exclusive prefix sum of the counts -> row starts (three small kernels, fixed order)
//     det_fill_kernel    entry {key = source pixel << 8 | tap << 2 | corner, w = corner weight * affinity} into the
//                        next free slot of its row (order arbitrary)
//     det_sort_kernel    one warp per row: every entry's rank = number of smaller keys in the row (keys are unique),
//                        written back in place -> rows are sorted by (source pixel, tap, corner): a canonical order
// and an iteration is ONE kernel without any scatter:
//     det_state_kernel   gs[q] = gy_{t+1}[q] * aff_ref[q] + sum over row(q), in row order, of gy_{t+1}[src] * w;
//                        then G, gy_t and the confidence gradient exactly as bwd_state_kernel's stage 1.
// Every sum runs in a fixed order inside one thread, so the result is a pure function of the inputs.  The gamma
// gradient takes per-block partial sums and a fixed-order tree instead of fp64 atomics.
//
// Value note: an entry carries (corner weight * affinity) pre-multiplied, so a contribution is gy * (w * a) where the
// reference forms w * (gy * a) (cuh:226-250): last-bit differences against the default forms, same tolerance class.
// Cost: 8 B per corner instead of 12 B per tap of geometry: 4N x 8 = 256 B (K = 3) / 768 B (K = 5) per pixel and
// iteration; the table build is about one to two iterations' worth.
#pragma once
#include "kernels_v2.cuh"

namespace nlspn {

constexpr int kDetScanTile = 1024;          // cells per block of the scan (256 threads x 4)
constexpr int kDetMaxRow = 8 * 32;          // row length one warp sorts in registers; longer rows: serial fallback

struct DetEntry {
    unsigned key;      // (source pixel within the image) << 8 | tap << 2 | corner
    float w;           // corner weight (cuh:71-79) * affinity of that tap
};

// visits the in-image corners of every valid neighbour tap of pixel (h, w): f(cell, tap, corner, weight)
template <int K, typename F>
__device__ __forceinline__ void det_for_each_corner(const float *__restrict__ ob, int P, int h, int w, int H, int W, F f)
{
    using G = Geo<K>;
#pragma unroll
    for (int t = 0; t < G::KK; ++t) {
        if (t == G::REF) continue;
        const float h_im = (float)(h - G::PAD + t / K) + __ldg(ob + (long)(2 * t) * P);
        const float w_im = (float)(w - G::PAD + t % K) + __ldg(ob + (long)(2 * t + 1) * P);
        if (!tap_valid(h_im, w_im, H, W)) continue;
        float hf, wf;
        int hl, wl;
        floor_small(h_im, hf, hl);
        floor_small(w_im, wf, wl);
        // mdmcn_get_gradient_weight, cuh:71-79 (literal)
        const float h1 = hf + 1.f, w1 = wf + 1.f;
        const float th = h1 - h_im, bh = (h_im + 1.f) - h1;
        const float lw_ = w1 - w_im, rw = (w_im + 1.f) - w1;
        const bool topv = hl >= 0, botv = hl + 1 <= H - 1, lefv = wl >= 0, rigv = wl + 1 <= W - 1;
        const int c0 = hl * W + wl;
        if (topv && lefv) f(c0, t, 0, th * lw_);
        if (topv && rigv) f(c0 + 1, t, 1, th * rw);
        if (botv && lefv) f(c0 + W, t, 2, bh * lw_);
        if (botv && rigv) f(c0 + W + 1, t, 3, bh * rw);
    }
}

template <int K>
__global__ void __launch_bounds__(kBlock)
det_count_kernel(const float *__restrict__ offset, int H, int W, unsigned *__restrict__ count)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    unsigned *cb = count + b * P;
    det_for_each_corner<K>(offset + b * 2 * G::KK * P + r, P, h, w, H, W,
                           [&](int cell, int, int, float) { atomicAdd(cb + cell, 1u); });
}

// ---- exclusive prefix sum of n counts -> start[0..n] (start[n] = total): tile sums, scan of the tile sums by one
// block, tile-local scan + base.  All sums are integers in a fixed order.
__global__ void __launch_bounds__(256)
det_scan_tiles_kernel(const unsigned *__restrict__ count, long n, unsigned *__restrict__ tile_sum)
{
    __shared__ unsigned red[8];
    const long i0 = (long)blockIdx.x * kDetScanTile + threadIdx.x * 4;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < n) s += count[i0 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int i = 0; i < 8; ++i) tot += red[i];
        tile_sum[blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(1024)
det_scan_sums_kernel(unsigned *__restrict__ tile_sum, long ntiles)
{
    // one block: in-place exclusive scan of the tile sums, 1024 at a time with a running carry
    __shared__ unsigned buf[1024];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long base = 0; base < ntiles; base += 1024) {
        const long i = base + threadIdx.x;
        const unsigned v = i < ntiles ? tile_sum[i] : 0u;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {      // Hillis-Steele inclusive scan
            const unsigned u = threadIdx.x >= (unsigned)o ? buf[threadIdx.x - o] : 0u;
            __syncthreads();
            buf[threadIdx.x] += u;
            __syncthreads();
        }
        const unsigned carry = carry_s;
        if (i < ntiles) tile_sum[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + buf[1023];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
det_scan_apply_kernel(unsigned *__restrict__ count, long n, const unsigned *__restrict__ tile_sum,
                      unsigned *__restrict__ start)
{
    // start[i] = exclusive prefix; count[] is cleared on the way (it becomes the fill cursor)
    __shared__ unsigned wsum[8];
    const long i0 = (long)blockIdx.x * kDetScanTile + threadIdx.x * 4;
    unsigned c[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = i0 + k < n ? count[i0 + k] : 0u;
        s += c[k];
    }
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= (unsigned)o) inc += u;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned base = tile_sum[blockIdx.x];
    for (int i = 0; i < (int)(threadIdx.x >> 5); ++i) base += wsum[i];
    unsigned run = base + inc - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) {
            start[i0 + k] = run;
            count[i0 + k] = 0u;
        }
        run += c[k];
    }
    if (i0 <= n - 1 && n - 1 < i0 + 4) start[n] = run;     // the thread that owns the last cell writes the total
}

template <int K>
__global__ void __launch_bounds__(kBlock)
det_fill_kernel(const float *__restrict__ offset, const float *__restrict__ aff, int H, int W,
                const unsigned *__restrict__ start, unsigned *__restrict__ cursor, DetEntry *__restrict__ entries)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const int h = r / W, w = r - h * W;
    const unsigned *sb = start + b * P;
    unsigned *cb = cursor + b * P;
    const float *ab = aff + b * G::KK * P + r;
    det_for_each_corner<K>(offset + b * 2 * G::KK * P + r, P, h, w, H, W, [&](int cell, int t, int c, float wgt) {
        const unsigned pos = sb[cell] + atomicAdd(cb + cell, 1u);
        DetEntry e;
        e.key = ((unsigned)r << 8) | ((unsigned)t << 2) | (unsigned)c;
        e.w = wgt * __ldg(ab + (long)t * P);
        entries[pos] = e;
    });
}

// one warp per row: rank by counting (keys are unique), in place
__global__ void __launch_bounds__(kBlock)
det_sort_kernel(const unsigned *__restrict__ start, long ncells, DetEntry *__restrict__ entries)
{
    __shared__ unsigned keys[kBlock / 32][kDetMaxRow];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const long cell = (long)blockIdx.x * (kBlock / 32) + wp;
    if (cell >= ncells) return;
    const unsigned s0 = start[cell], L = start[cell + 1] - s0;
    if (L <= 1) return;
    DetEntry *row = entries + s0;
    if (L > (unsigned)kDetMaxRow) {
        // rare (far beyond the mean row length of 4N): insertion sort by one lane
        if (lane == 0) {
            for (unsigned i = 1; i < L; ++i) {
                const DetEntry e = row[i];
                unsigned j = i;
                while (j > 0 && row[j - 1].key > e.key) {
                    row[j] = row[j - 1];
                    --j;
                }
                row[j] = e;
            }
        }
        return;
    }
    DetEntry mine[kDetMaxRow / 32];
#pragma unroll
    for (int k = 0; k < kDetMaxRow / 32; ++k) {
        const unsigned i = lane + 32 * k;
        if (i < L) {
            mine[k] = row[i];
            keys[wp][i] = mine[k].key;
        }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kDetMaxRow / 32; ++k) {
        const unsigned i = lane + 32 * k;
        if (i < L) {
            unsigned rank = 0;
            for (unsigned j = 0; j < L; ++j) rank += keys[wp][j] < mine[k].key ? 1u : 0u;
            row[rank] = mine[k];        // every entry was loaded before the first store (the __syncwarp above)
        }
    }
}

// ======================================================================================
// One backward iteration, deterministic form.  gy_next = gy_{t+1} (nullptr at t = T).  With gy_out == nullptr the
// kernel only writes the gathered plane s_out (the gradient that reaches x_0, consumed by final_bwd_kernel).
// ======================================================================================
template <int K>
__global__ void __launch_bounds__(kBlock)
det_state_kernel(const unsigned *__restrict__ start, const DetEntry *__restrict__ entries,
                 const float *__restrict__ aff, const float *__restrict__ conf, const float *__restrict__ dep,
                 const float *__restrict__ x_t, const float *__restrict__ g_ext, const float *__restrict__ gy_next,
                 float *__restrict__ gy_out, float *__restrict__ s_out, float *__restrict__ g_conf_acc,
                 unsigned flags, int H, int W)
{
    using G = Geo<K>;
    const int P = H * W;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= P) return;
    const long b = blockIdx.y;
    const long q = b * P + r;
    float gs = 0.f;
    if (gy_next) {
        const float *gyb = gy_next + b * P;
        gs = __ldg(gyb + r) * __ldg(aff + (b * G::KK + G::REF) * P + r);      // centre tap: structurally zero offset
        const unsigned s0 = start[q], s1 = start[q + 1];
        const uint2 *row = reinterpret_cast<const uint2 *>(entries);
        for (unsigned i = s0; i < s1; ++i) {
            const uint2 e = __ldg(row + i);
            gs += __ldg(gyb + (e.x >> 8)) * __uint_as_float(e.y);
        }
    }
    if (!gy_out) {
        s_out[q] = gs;
        return;
    }
    // nlspnmodel.py:351,357,361 backwards; identical to bwd_state_kernel's stage 1
    const float gext = g_ext ? __ldg(g_ext + q) : 0.f;
    const float cf = conf ? __ldg(conf + q) : 1.f;
    const bool need_x = (gy_next && conf) || (flags & kAlwaysClip);
    const float xt = need_x ? __ldg(x_t + q) : 1.f;
    const float dp = (flags & kPreserve) ? __ldg(dep + q) : 0.f;
    float Gx = gext;
    if (flags & kBlendPre) {
        if (gy_next) Gx += (flags & kPreserve) ? (1.0f - (dp > 0.f ? 1.f : 0.f)) * gs : gs;
    } else {
        if (gy_next) Gx += conf ? cf * gs : gs;
        if ((flags & kAlwaysClip) && was_clipped(xt)) Gx = 0.f;
        if (flags & kPreserve) Gx = (1.0f - (dp > 0.f ? 1.f : 0.f)) * Gx;
    }
    if (gy_next && conf) g_conf_acc[q] += xt * gs;
    gy_out[q] = Gx;
}

// fixed-order sum of n per-block partial gamma gradients (one block)
__global__ void __launch_bounds__(256)
gamma_reduce_det_kernel(const double *__restrict__ part, long n, double *__restrict__ g_gamma)
{
    __shared__ double red[256];
    double v = 0.0;
    for (long i = threadIdx.x; i < n; i += 256) v += part[i];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < (unsigned)o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *g_gamma = red[0];
}

} // namespace nlspn
