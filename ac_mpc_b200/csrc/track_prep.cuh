// track_prep.cuh -- the track side of the MPC step:
//   utils/load.py:30-35              remove_near_duplicate_points: drop map points closer than 1e-4 m to their predecessor
//                                    (applied to the three lines of a loaded map, load.py:9-27)
//   perception/utils.py:107-119      smooth_track_with_polyfit: x = poly(y) least-squares fit, 500-point scan for the point
//                                    nearest the ego origin, resample num_points from there to max(y)
//   perception/tracks.py:247-252     _calculate_centre_track: (left + right) / 2 with 10 origin points prepended, degree-2 fit
//   SURVEY.md section 8d             instance -> get_control input: the next `lookahead` metres of a map centre line, resampled
//                                    to H points and expressed in the (perturbed) ego frame -- on the device, so that a track
//                                    sweep uploads 20 bytes per instance instead of 1200
// The fit is NOT numpy's algorithm (SVD least squares of the column-scaled Vandermonde matrix): it builds the discrete
// orthogonal polynomials of the sample abscissae by their three-term recurrence (Forsythe), which needs nothing but
// warp reductions, stores no matrix and does not square the condition number.  Same polynomial for full-rank input
// (tests: 1e-9 m); with fewer than degree + 1 distinct abscissae numpy returns the minimum-norm solution under a
// RankWarning, this kernel fits the highest degree the data determine and says so in status[].
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace acmpc {
namespace trk {

constexpr unsigned kAll = 0xffffffffu;
constexpr int kDupThreads = 256;
constexpr int kMaxDegree = 3;
constexpr int kScanPoints = 500;       // perception/utils.py:112

// status[]: ACMPC_TRACK_OK / ACMPC_TRACK_EMPTY (len(track) == 0: the stub line of perception/utils.py:108-111) /
// ACMPC_TRACK_RANK_DEFICIENT (degree reduced to what the abscissae determine) -- include/acmpc_b200.h

// ---- remove_near_duplicate_points ---------------------------------------------------------------------------------
__device__ __forceinline__ bool keep_point(const double* __restrict__ xy, int i, int M, double tol)
{
    if (i >= M) return false;
    if (i == 0) return true;
    const double dx = xy[2 * i] - xy[2 * i - 2], dy = xy[2 * i + 1] - xy[2 * i - 1];
    return hypot(dx, dy) > tol;      // dists > 0.0001, load.py:34 (NaN compares false, as in numpy)
}

__global__ void near_duplicate_count_kernel(const double* __restrict__ xy, int M, double tol, int* __restrict__ cta_counts)
{
    const int i = blockIdx.x * kDupThreads + threadIdx.x;
    const int c = __syncthreads_count(keep_point(xy, i, M, tol));
    if (threadIdx.x == 0) cta_counts[blockIdx.x] = c;
}

__global__ void near_duplicate_scatter_kernel(const double* __restrict__ xy, int M, double tol,
                                              const int* __restrict__ cta_counts, double* __restrict__ out,
                                              int* __restrict__ kept)
{
    __shared__ int base, warp_off[kDupThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * kDupThreads + threadIdx.x;
    const bool keep = keep_point(xy, i, M, tol);
    const unsigned mask = __ballot_sync(kAll, keep);
    if (lane == 0) warp_off[warp] = __popc(mask);
    if (warp == 0) {                   // rows kept by the CTAs before this one
        int s = 0;
        for (int c = lane; c < (int)blockIdx.x; c += 32) s += cta_counts[c];
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(kAll, s, d);
        if (lane == 0) base = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < kDupThreads / 32; ++w) {
            const int c = warp_off[w];
            warp_off[w] = run, run += c;
        }
        if (blockIdx.x == gridDim.x - 1) kept[0] = base + run;
    }
    __syncthreads();
    if (keep) {
        const int pos = base + warp_off[warp] + __popc(mask & ((1u << lane) - 1u));
        out[2 * pos] = xy[2 * i], out[2 * pos + 1] = xy[2 * i + 1];
    }
}

// ---- smooth_track_with_polyfit ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kAll, v, d);
    return v;
}

struct OrthoFit {
    double a[kMaxDegree + 1], b[kMaxDegree + 1], c[kMaxDegree + 1];   // recurrence (alpha, beta) and coefficients
    int degree;
    // p_k(t) for k = 0..degree by p_{k+1} = (t - a_k) p_k - b_k p_{k-1}; returns sum c_k p_k
    __device__ __forceinline__ double eval(double t) const
    {
        double pm = 0.0, p = 1.0, s = c[0];
        for (int k = 0; k < degree; ++k) {
            const double pn = (t - a[k]) * p - b[k] * pm;
            pm = p, p = pn;
            s += c[k + 1] * p;
        }
        return s;
    }
    __device__ __forceinline__ double basis(double t, int k) const
    {
        double pm = 0.0, p = 1.0;
        for (int j = 0; j < k; ++j) {
            const double pn = (t - a[j]) * p - b[j] * pm;
            pm = p, p = pn;
        }
        return p;
    }
};

// np.linspace(start, stop, num)[j] (numpy/_core/function_base.py: arange * step + start, last sample = stop), no FMA
__device__ __forceinline__ double linspace_at(double start, double stop, int num, int j)
{
    if (num < 2) return start;
    if (j == num - 1) return stop;
    const double div = (double)(num - 1), delta = stop - start, step = delta / div;
    if (step == 0.0) return __dadd_rn(__dmul_rn((double)j / div, delta), start);
    return __dadd_rn(__dmul_rn((double)j, step), start);
}

// One warp per track.  points[offsets[b] .. offsets[b+1]) rows of (x, y); out[b, num_points, 2].
// pad_origin != 0: the track is preceded by `pad_origin` virtual points (x of its first row, 0) -- the
// origin_points of tracks.py:249-251 -- without materialising the concatenation.
__global__ void polyfit_resample_kernel(const double* __restrict__ points, const int* __restrict__ offsets, int B,
                                        int num_points, int degree, int pad_origin, double* __restrict__ out,
                                        int* __restrict__ status, int* __restrict__ start_index)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const int lo = offsets[b], m = offsets[b + 1] - lo;
    double* o = out + (size_t)b * num_points * 2;
    if (m <= 0) {                                   // perception/utils.py:108-111
        for (int j = lane; j < num_points; j += 32)
            o[2 * j] = linspace_at(0.0, 0.1, num_points, j), o[2 * j + 1] = linspace_at(0.0, 2.0, num_points, j);
        if (lane == 0) {
            if (status) status[b] = ACMPC_TRACK_EMPTY;
            if (start_index) start_index[b] = 0;
        }
        return;
    }
    const double* p = points + (size_t)lo * 2;
    const int pad = pad_origin > 0 ? pad_origin : 0, n = m + pad;
    const double x_pad = p[0];
    auto px = [&](int i) { return i < pad ? x_pad : p[2 * (i - pad)]; };
    auto py = [&](int i) { return i < pad ? 0.0 : p[2 * (i - pad) + 1]; };

    double ymax = -INFINITY;
    for (int i = lane; i < n; i += 32) ymax = fmax(ymax, py(i));      // np.max propagates NaN; inputs are finite
    for (int d = 16; d; d >>= 1) ymax = fmax(ymax, __shfl_xor_sync(kAll, ymax, d));

    OrthoFit f;
    f.degree = degree;
    for (int k = 0; k <= kMaxDegree; ++k) f.a[k] = f.b[k] = f.c[k] = 0.0;
    double g_prev = 1.0;
    int st = ACMPC_TRACK_OK;
    const double rcond = (double)n * 2.220446049250313e-16;            // np.polyfit: rcond = len(x) * eps
    for (int k = 0; k <= degree; ++k) {
        double g = 0.0, gt = 0.0, gf = 0.0, raw = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double t = py(i), pk = f.basis(t, k);
            g += pk * pk, gt += t * pk * pk, gf += px(i) * pk;
            double r = 1.0;
            for (int j = 0; j < k; ++j) r *= t * t;
            raw += r;
        }
        g = warp_sum(g), gt = warp_sum(gt), gf = warp_sum(gf), raw = warp_sum(raw);
        if (k > 0 && !(g > rcond * rcond * raw)) {   // column k lies in the span of the lower ones
            f.degree = k - 1, st = ACMPC_TRACK_RANK_DEFICIENT;
            break;
        }
        f.c[k] = gf / g, f.a[k] = gt / g;
        f.b[k] = k > 0 ? g / g_prev : 0.0;
        g_prev = g;
    }

    // ynew = linspace(0, ymax, 500); start_index = argmin(||(poly(ynew), ynew)||), first minimum
    double best = INFINITY;
    int arg = 0x7fffffff;
    for (int j = lane; j < kScanPoints; j += 32) {
        const double y = linspace_at(0.0, ymax, kScanPoints, j), x = f.eval(y);
        const double r = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
        if (r < best) best = r, arg = j;
    }
    for (int d = 16; d; d >>= 1) {
        const double ob = __shfl_xor_sync(kAll, best, d);
        const int oa = __shfl_xor_sync(kAll, arg, d);
        if (ob < best || (ob == best && oa < arg)) best = ob, arg = oa;
    }
    if (arg == 0x7fffffff) arg = 0;
    const double y0 = linspace_at(0.0, ymax, kScanPoints, arg);
    for (int j = lane; j < num_points; j += 32) {
        const double y = linspace_at(y0, ymax, num_points, j);
        o[2 * j] = f.eval(y), o[2 * j + 1] = y;
    }
    if (lane == 0) {
        if (status) status[b] = st;
        if (start_index) start_index[b] = arg;
    }
}

// (left + right) / 2, tracks.py:248; rows [B, N, 2]
__global__ void midline_kernel(const double* __restrict__ left, const double* __restrict__ right, size_t count,
                               double* __restrict__ mid)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) mid[i] = (left[i] + right[i]) / 2.0;
}

// ---- instance extraction -------------------------------------------------------------------
// One thread per output point (b, k).  centreline [M, 2] closed loop sampled every `ds` metres.
__global__ void extract_paths_kernel(const double* __restrict__ cl, int M, const int* __restrict__ index,
                                     const double* __restrict__ offset_lat, const double* __restrict__ offset_psi, int B,
                                     int H, double lookahead, double ds, double step_s, double step_w,
                                     double* __restrict__ paths)
{   // step_s = lookahead / (H - 1), step_w = (6 - 10) / (H - 1): the np.linspace steps, divided once on the host
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * H) return;
    const int b = t / H, k = t - b * H;
    int i = index[b] % M;
    if (i < 0) i += M;
    const int i1 = i + 1 == M ? 0 : i + 1;
    const double ox = cl[2 * i], oy = cl[2 * i + 1];
    // (cos, sin) of the tangent heading th = atan2(ty, tx) is the normalised tangent itself; th + psi by the angle
    // addition formulas -- no atan2 and one small-argument sincos per thread instead of atan2 + two sincos
    const double tx = cl[2 * i1] - ox, ty = cl[2 * i1 + 1] - oy, rn = 1.0 / sqrt(tx * tx + ty * ty);
    const double cs = tx * rn, sn = ty * rn;
    const double lat = offset_lat ? offset_lat[b] : 0.0, psi = offset_psi ? offset_psi[b] : 0.0;
    const double gx = ox + lat * -sn, gy = oy + lat * cs;             // ego origin: `lat` metres along the left normal
    double sp, cp;
    sincos(psi, &sp, &cp);
    const double se = sn * cp + cs * sp, ce = cs * cp - sn * sp;
    const double s = (k == H - 1 ? lookahead : __dmul_rn((double)k, step_s)) / ds;
    const double fl = floor(s), frac = s - fl;
    int ia = (int)((i + (long long)fl) % M);
    const int ib = ia + 1 == M ? 0 : ia + 1;
    const double qx = cl[2 * ia] * (1.0 - frac) + cl[2 * ib] * frac, qy = cl[2 * ia + 1] * (1.0 - frac) + cl[2 * ib + 1] * frac;
    const double rx = qx - gx, ry = qy - gy;
    double* o = paths + (size_t)t * 3;
    o[0] = rx * se + ry * -ce;                                        // x right
    o[1] = rx * ce + ry * se;                                         // y forward
    o[2] = k == H - 1 ? 6.0 : __dadd_rn(__dmul_rn((double)k, step_w), 10.0);   // np.linspace(10, 6, H), controller.py:264
}

}   // namespace trk
}   // namespace acmpc
