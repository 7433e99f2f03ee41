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

#include <type_traits>

namespace acmpc {
namespace trk {

constexpr unsigned kAll = 0xffffffffu;
constexpr int kDupThreads = 256;
constexpr int kMaxDegree = 3;
constexpr int kScanPoints = 500;       // perception/utils.py:112

// status[]: ACMPC_TRACK_OK / ACMPC_TRACK_EMPTY (len(track) == 0: the stub line of perception/utils.py:108-111) /
// ACMPC_TRACK_RANK_DEFICIENT (degree reduced to what the abscissae determine) -- include/acmpc_b200.h

// ---- remove_near_duplicate_points ---------------------------------------------------------------------------------
// rows of `ld` doubles; columns 0 and 1 are (x, y), further columns (z, width ...) travel with their row like numpy's
// track[is_not_duplicated]
__device__ __forceinline__ bool keep_point(const double* __restrict__ xy, int i, int M, int ld, double tol)
{
    if (i >= M) return false;
    if (i == 0) return true;
    const double dx = xy[(size_t)ld * i] - xy[(size_t)ld * (i - 1)], dy = xy[(size_t)ld * i + 1] - xy[(size_t)ld * (i - 1) + 1];
    return hypot(dx, dy) > tol;      // dists > 0.0001, load.py:34 (NaN compares false, as in numpy)
}

__global__ void near_duplicate_count_kernel(const double* __restrict__ xy, int M, int ld, double tol, int* __restrict__ cta_counts)
{
    const int i = blockIdx.x * kDupThreads + threadIdx.x;
    const int c = __syncthreads_count(keep_point(xy, i, M, ld, tol));
    if (threadIdx.x == 0) cta_counts[blockIdx.x] = c;
}

__global__ void near_duplicate_scatter_kernel(const double* __restrict__ xy, int M, int ld, double tol,
                                              const int* __restrict__ cta_counts, double* __restrict__ out,
                                              int* __restrict__ kept)
{
    __shared__ int base, warp_off[kDupThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * kDupThreads + threadIdx.x;
    const bool keep = keep_point(xy, i, M, ld, tol);
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
        for (int c = 0; c < ld; ++c) out[(size_t)ld * pos + c] = xy[(size_t)ld * i + c];
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
    // every loop below is unrolled over the compile-time kMaxDegree so that a / b / c stay in registers
    // p_k(t) for k = 0..degree by p_{k+1} = (t - a_k) p_k - b_k p_{k-1}; returns sum c_k p_k
    __device__ __forceinline__ double eval(double t) const
    {
        double pm = 0.0, p = 1.0, s = c[0];
#pragma unroll
        for (int k = 0; k < kMaxDegree; ++k)
            if (k < degree) {
                const double pn = (t - a[k]) * p - b[k] * pm;
                pm = p, p = pn;
                s += c[k + 1] * p;
            }
        return s;
    }
    template <int K>
    __device__ __forceinline__ double basis(double t) const
    {
        double pm = 0.0, p = 1.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
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
    auto pass = [&](auto kc) {                          // one degree: three warp reductions over the track's points
        constexpr int k = decltype(kc)::value;
        if (k > degree || st != ACMPC_TRACK_OK) return;     // warp-uniform
        double g = 0.0, gt = 0.0, gf = 0.0, raw = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double t = py(i), pk = f.template basis<k>(t);
            g += pk * pk, gt += t * pk * pk, gf += px(i) * pk;
            double r = 1.0;
#pragma unroll
            for (int j = 0; j < k; ++j) r *= t * t;
            raw += r;
        }
        g = warp_sum(g), gt = warp_sum(gt), gf = warp_sum(gf), raw = warp_sum(raw);
        if (k > 0 && !(g > rcond * rcond * raw)) {          // column k lies in the span of the lower ones
            f.degree = k - 1, st = ACMPC_TRACK_RANK_DEFICIENT;
            return;
        }
        f.c[k] = gf / g, f.a[k] = gt / g;
        f.b[k] = k > 0 ? g / g_prev : 0.0;
        g_prev = g;
    };
    pass(std::integral_constant<int, 0>{});
    pass(std::integral_constant<int, 1>{});
    pass(std::integral_constant<int, 2>{});
    pass(std::integral_constant<int, 3>{});
    static_assert(kMaxDegree == 3, "one pass per degree");

    // ynew = linspace(0, ymax, 500); start_index = argmin(||(poly(ynew), ynew)||), first minimum.  The comparison is on
    // sqrt(x^2 + y^2) as numpy rounds it (two samples whose squares differ can share a square root, and argmin then
    // takes the first), but the square root is only taken for samples whose square is within a few ulp of the lane's
    // running minimum or below it.
    const double scan_step = ymax / (double)(kScanPoints - 1);
    double best = INFINITY, gate = INFINITY;
    int arg = 0x7fffffff;
    for (int j = lane; j < kScanPoints; j += 32) {
        const double y = j == kScanPoints - 1 ? ymax : __dmul_rn((double)j, scan_step);   // linspace: j * step + 0.0
        const double x = f.eval(y), r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        if (r2 <= gate) {
            const double r = sqrt(r2);
            if (r < best) best = r, arg = j, gate = r2 * (1.0 + 0x1p-48);
        }
    }
    for (int d = 16; d; d >>= 1) {
        const double ob = __shfl_xor_sync(kAll, best, d);
        const int oa = __shfl_xor_sync(kAll, arg, d);
        if (ob < best || (ob == best && oa < arg)) best = ob, arg = oa;
    }
    if (arg == 0x7fffffff) arg = 0;
    const double y0 = linspace_at(0.0, ymax, kScanPoints, arg);
    const double out_step = num_points > 1 ? (ymax - y0) / (double)(num_points - 1) : 0.0;
    for (int j = lane; j < num_points; j += 32) {
        const double y = (out_step == 0.0 || j == num_points - 1) ? linspace_at(y0, ymax, num_points, j)
                                                                  : __dadd_rn(__dmul_rn((double)j, out_step), y0);
        reinterpret_cast<double2*>(o)[j] = make_double2(f.eval(y), y);
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
// centreline [M, 2]: closed loop sampled every `ds` metres.  A CTA of kExtractThreads threads builds kExtractGroup
// instances: the per-step table (map offset, interpolation weight, width -- the same for every instance) and the
// per-instance ego frames are computed once into shared memory, then every warp walks whole instances, one output point
// per lane: two gathers from the L2-resident centre line, a 2x2 rotation and 24 contiguous bytes out.  HBM-bound on the
// [B, H, 3] store.  step_s = lookahead / (H - 1), step_w = (6 - 10) / (H - 1): the np.linspace steps, divided on the host.
constexpr int kExtractThreads = 256;
constexpr int kExtractGroup = 32;

struct EgoFrame {
    double gx, gy, se, ce;    // ego origin, sin / cos of the ego heading
    int i;                    // map index of the instance
};

__global__ void __launch_bounds__(kExtractThreads)
extract_paths_kernel(const double* __restrict__ cl, int M, const int* __restrict__ index,
                     const double* __restrict__ offset_lat, const double* __restrict__ offset_psi, int B, int H,
                     double lookahead, double ds, double step_s, double step_w, double* __restrict__ paths)
{
    extern __shared__ double smem[];
    double* frac = smem;                              // [H]
    double* width = smem + H;                         // [H]
    int* step = reinterpret_cast<int*>(smem + 2 * H); // [H], already reduced mod M
    __shared__ EgoFrame frame[kExtractGroup];
    const int b0 = blockIdx.x * kExtractGroup, nb = min(kExtractGroup, B - b0);
    for (int k = threadIdx.x; k < H; k += kExtractThreads) {
        const double s = (k == H - 1 ? lookahead : __dmul_rn((double)k, step_s)) / ds, fl = floor(s);
        frac[k] = s - fl;
        step[k] = (int)((long long)fl % M);
        width[k] = k == H - 1 ? 6.0 : __dadd_rn(__dmul_rn((double)k, step_w), 10.0);   // np.linspace(10, 6, H), controller.py:264
    }
    if (threadIdx.x < nb) {
        const int b = b0 + threadIdx.x;
        int i = index[b] % M;
        if (i < 0) i += M;
        const int i1 = i + 1 == M ? 0 : i + 1;
        const double ox = cl[2 * i], oy = cl[2 * i + 1];
        // (cos, sin) of the tangent heading atan2(ty, tx) is the normalised tangent itself; heading + psi by the angle
        // addition formulas: no atan2, one small-argument sincos
        const double tx = cl[2 * i1] - ox, ty = cl[2 * i1 + 1] - oy, rn = 1.0 / sqrt(tx * tx + ty * ty);
        const double cs = tx * rn, sn = ty * rn;
        const double lat = offset_lat ? offset_lat[b] : 0.0, psi = offset_psi ? offset_psi[b] : 0.0;
        double sp, cp;
        sincos(psi, &sp, &cp);
        EgoFrame f;
        f.gx = ox + lat * -sn, f.gy = oy + lat * cs;                  // `lat` metres along the left normal
        f.se = sn * cp + cs * sp, f.ce = cs * cp - sn * sp;
        f.i = i;
        frame[threadIdx.x] = f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* tile = smem + 3 * H + (size_t)warp * 3 * H;   // this warp's [H, 3] staging tile (after the table; 2.5 H doubles used)
    for (int g = warp; g < nb; g += kExtractThreads / 32) {
        const EgoFrame f = frame[g];
        for (int k = lane; k < H; k += 32) {
            int ia = f.i + step[k];
            if (ia >= M) ia -= M;
            const int ib = ia + 1 == M ? 0 : ia + 1;
            const double2 pa = reinterpret_cast<const double2*>(cl)[ia], pb = reinterpret_cast<const double2*>(cl)[ib];
            const double w = frac[k];
            const double rx = pa.x * (1.0 - w) + pb.x * w - f.gx, ry = pa.y * (1.0 - w) + pb.y * w - f.gy;
            tile[3 * k] = rx * f.se + ry * -f.ce;                     // x right
            tile[3 * k + 1] = rx * f.ce + ry * f.se;                  // y forward
            tile[3 * k + 2] = width[k];
        }
        __syncwarp();
        double* o = paths + (size_t)(b0 + g) * H * 3;                 // 24 H contiguous bytes, 256 per store instruction
        for (int e = lane; e < 3 * H; e += 32) o[e] = tile[e];
        __syncwarp();
    }
}

}   // namespace trk
}   // namespace acmpc
