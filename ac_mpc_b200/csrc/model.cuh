// model.cuh -- the SpatialBicycleModel entry points of the reference as stand-alone batched kernels
// (/root/reference/src/acmpc/control/dynamics.py:23-103, spatial_mpc.py:156-168).  Inside a get_control step these
// transforms are fused into acmpc_control_kernel (mpc_warp.cuh: ControlQP::setup, control_instance); the kernels here
// exist so that the object API of the drop-in (SpatialBicycleModel.t2s / s2t / linearise, SpatialMPC.update_prediction)
// has the same surface as the reference's.  One thread per (instance, waypoint); plain elementwise arithmetic, written
// with the explicit-rounding intrinsics so that no FMA contraction moves a result away from numpy's.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace acmpc {
namespace model {

constexpr double kEps = 1e-12;   // SpatialBicycleModel._eps, dynamics.py:21
constexpr double kPiM = 3.14159265358979323846;

// dynamics.py:23-40.  way[B,3] = (x, y, psi) of the reference waypoint, state[B,3] = (x, y, psi) of the vehicle.
__global__ void t2s_kernel(const double* __restrict__ way, const double* __restrict__ state, int B, double* __restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double rx = way[3 * b], ry = way[3 * b + 1], rpsi = way[3 * b + 2];
    const double x = state[3 * b], y = state[3 * b + 1], psi = state[3 * b + 2];
    // e_y = cos(ref_psi) * (y - ref_y) - sin(ref_psi) * (x - ref_x)
    out[3 * b] = __dsub_rn(__dmul_rn(cos(rpsi), __dsub_rn(y, ry)), __dmul_rn(sin(rpsi), __dsub_rn(x, rx)));
    // e_psi = mod(psi - ref_psi + pi, 2 pi) - pi   (numpy's mod: result takes the sign of the divisor)
    double r = fmod(__dadd_rn(__dsub_rn(psi, rpsi), kPiM), 2.0 * kPiM);
    if (r < 0.0) r = __dadd_rn(r, 2.0 * kPiM);
    out[3 * b + 1] = __dsub_rn(r, kPiM);
    out[3 * b + 2] = 0.0;
}

// dynamics.py:42-63.  way[B,7,n] ReferencePath rows, states[B,n,3] -> out[B,3,n] rows X, Y, Psi.
// pred (may be NULL) [B,n,2] = SpatialMPC.update_prediction = s2t(...)[:-1].T (spatial_mpc.py:156-168).
__global__ void s2t_kernel(const double* __restrict__ way, const double* __restrict__ states, int B, int n,
                           double* __restrict__ out, double* __restrict__ pred)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n) return;
    const int b = i / n, k = i - b * n;
    const double* w = way + (size_t)b * 7 * n;
    const double xs = w[k], ys = w[n + k], psi = w[2 * n + k];
    const double ey = states[(size_t)i * 3], ep = states[(size_t)i * 3 + 1];
    const double X = __dsub_rn(xs, __dmul_rn(ey, sin(psi))), Y = __dadd_rn(ys, __dmul_rn(ey, cos(psi)));
    if (out) {
        double* o = out + (size_t)b * 3 * n;
        o[k] = X, o[n + k] = Y, o[2 * n + k] = __dadd_rn(psi, ep);
    }
    if (pred) pred[(size_t)i * 2] = X, pred[(size_t)i * 2 + 1] = Y;
}

// dynamics.py:65-103.  way[B,7,n] -> f[B,n,3], A[B,n,3,3], Bm[B,n,3,2].
__global__ void linearise_kernel(const double* __restrict__ way, int B, int n, double* __restrict__ f,
                                 double* __restrict__ A, double* __restrict__ Bm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n) return;
    const int b = i / n, k = i - b * n;
    const double* w = way + (size_t)b * 7 * n;
    const double ka = w[3 * n + k], d = w[4 * n + k], v = w[6 * n + k];
    const double vd = __dadd_rn(__dmul_rn(v, d), kEps);                      // v_ref * delta_s + eps
    const double v2d = __dadd_rn(__dmul_rn(__dmul_rn(v, v), d), kEps);       // v_ref**2 * delta_s + eps
    if (A) {
        double* a = A + (size_t)i * 9;
        a[0] = 1.0, a[1] = d, a[2] = 0.0;
        a[3] = __dmul_rn(-__dmul_rn(ka, ka), d), a[4] = 1.0, a[5] = 0.0;
        a[6] = __ddiv_rn(-ka, vd), a[7] = 0.0, a[8] = 1.0;
    }
    if (Bm) {
        double* m = Bm + (size_t)i * 6;
        m[0] = 0.0, m[1] = 0.0, m[2] = 0.0, m[3] = d, m[4] = __ddiv_rn(-1.0, v2d), m[5] = 0.0;
    }
    if (f) {
        double* o = f + (size_t)i * 3;
        o[0] = 0.0, o[1] = 0.0, o[2] = __ddiv_rn(1.0, vd);
    }
}

}  // namespace model
}  // namespace acmpc
