// mpc_warp.cuh -- one warp solves one MPC step (SpatialMPC.get_control), warp-parallel over the horizon.
//
// Reference path being replaced (file:line under /root/reference/src/acmpc/control/):
//   spatial_mpc.py:125-154   construct_waypoints          -> build_waypoints()
//   solvers/speed_profile.py:26-59,131-150 + osqp        -> SpeedQP
//   dynamics.py:23-40 (t2s), :65-103 (linearise)          -> ControlQP::setup()
//   solvers/control.py:26-79,121-158 + osqp               -> ControlQP
//   spatial_mpc.py:193-212, dynamics.py:42-63 (s2t)       -> solve_instance() tail
//
// Execution model (revision r1b):
//   * lane l owns the C = ceil(H/32) CONTIGUOUS horizon stages C*l .. C*l+C-1; C is a template
//     parameter so that all per-stage state is register arrays.
//   * stage s owns  x_s = (e_y, e_psi, t),  u_{s-1} = (v, kappa_cmd),  the equality block
//     "m*x_s + A_{s-1} x_{s-1} + B_{s-1} u_{s-1} = b_s", the bound rows of x_s and u_{s-1}, the entries of
//     A_s (columns of x_s in block s+1) and of B_{s-1}.  With this ownership one ADMM iteration
//     exchanges only two 3-vectors between neighbouring stages (register shuffles).
//   * the ADMM iterates (x, z, y) stay in REGISTERS for the whole solve; the scaled problem data and the
//     factorisation live in shared memory, one conflict-free column per lane.
//   * OSQP's linear system is solved in reduced form  (P + sigma I + A' diag(rho) A) x~ = rhs:  the inputs
//     are eliminated analytically, the remaining SPD block-tridiagonal system (3x3 blocks) is factorised
//     once per rho (block LDL') and each solve is two warp-parallel affine scans (Kogge-Stone over lanes
//     with precomputed 3x3 prefix products; the backward scan reads the transposed forward matrices).
//   * the speed-profile QP is the scalar version of the same scheme and lives entirely in registers.
// Everything else (Ruiz equilibration, rho classes, relaxation, projection, dual update, unscaled
// residuals, infeasibility certificates, adaptive rho) follows OSQP 0.6.x step for step, because the
// reference's answer is defined by OSQP's iterate at its termination check.
//
// The same source is compiled (a) by nvcc for sm_100a into the product library and (b) by g++ with
// -DACMPC_EMULATE into tests/_emul (a 32-lane lock-step emulation, TEST ONLY, see simt.cuh).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/acmpc_b200.h"
#include "simt.cuh"

namespace acmpc {

constexpr double kInfty = 1e30;
constexpr double kMinScaling = 1e-4;
constexpr double kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6;
constexpr double kRhoMax = 1e6;
constexpr double kRhoEqOverIneq = 1e3;
constexpr double kRhoTol = 1e-4;
constexpr double kBig = kInfty * kMinScaling;   // "infinite bound" threshold in scaled space
constexpr double kPi = 3.14159265358979323846;
constexpr int kLevels = 5;                      // Kogge-Stone levels over 32 lanes

// ------------------------------------------------------------------------------------------------
// On-chip data map of one instance (one warp).
//
// TENSOR MEMORY (lane private, 64 doubles = 128 columns per stage; stage (lane, j) of the instance starts
// at double column j*64 of the lane's row).  Everything the ADMM iteration reads every time lives here and
// is fetched as 16-double chunks (one tcgen05.ld.32x32b.x32 each):
//   chunk F  [ 0,16)  rho[5] iv ik rb31 rb22 rinv[5] - -        rewritten by every factorisation
//   chunk G  [16,32)  s[5] m[3] b22 b31 a11 a12 a21 a22 a31 a33 scaled constraint matrix
//   chunk H  [32,48)  q[2] be[3] lb[5] ub[5] -                  scaled cost / bounds
//   chunk NS [48,64)  N_s[9] Sigma_s^{-1}[6] -                  block LDL' factor
// SHARED MEMORY (per warp): cold per-stage fields (termination checks, outputs), one column per lane,
// field f of stage (lane, j) at S[(f*C + j)*32 + lane]; then the scratch / scan region, read ACROSS lanes:
// the raw path slice (TMA destination), the serial factorisation's work arrays, and the lane-level prefix
// products of the two scans.
// ------------------------------------------------------------------------------------------------
enum : int {
    T_STRIDE = 64,
    T_F = 0, T_G = 16, T_H = 32, T_NS = 48,
    // offsets inside the chunks
    FC_RHO = 0, FC_IV = 5, FC_IK = 6, FC_RB31 = 7, FC_RB22 = 8, FC_RINV = 9,   // [0,9) is what the rhs phase reads
    GC_S = 0, GC_M = 5, GC_B22 = 8, GC_B31 = 9, GC_A = 10,   // a11 a12 a21 a22 a31 a33
    HC_Q = 0, HC_BE = 2, HC_LB = 5, HC_UB = 10,
    NC_N = 0, NC_SI = 9
};
enum : int {
    F_XS = 0, F_YS, F_PSI,          // ReferencePath rows needed by the rollout (paths.py:4-72)
    K_P,                            // 5: diagonal of P
    K_DI = K_P + 5,                 // 5: 1/D
    K_EEI = K_DI + 5,               // 3: 1/E of the equality block
    K_EBI = K_EEI + 3,              // 5: 1/E of the bound rows
    K_FIELDS = K_EBI + 5,
    // work arrays of the factorisations inside the scratch region (per-stage columns like the fields above)
    W_SS = 0,                       // 6: S_ss -> Sigma_s^{-1}
    W_OFF = 6,                      // 9: S_{s,s-1} -> N_s
    W_FIELDS = 15
};

// Warm-start record of one instance (the reference's three persistent OSQP objects, spatial_mpc.py:43-58:
// speed-profile solver, localised speed-profile solver, control solver): header, then the SCALED iterates
// exactly as OSQP keeps them between solves, one 32-lane row per (field, stage-in-lane).
enum : int {
    WR_RHO = 0,       // 3: rho of slot 0 (speed), 1 (localised speed), 2 (control)
    WR_VALID = 3,     // 3: 1.0 once the slot holds a solve (0.0 = the object does not exist yet: cold setup)
    WR_HEADER = 8,
    WR_SPEED_FIELDS = 5,    // x za zb ya yb
    WR_CTRL_FIELDS = 21     // x[5] zb[5] yb[5] ye[3] ze[3]
};

template <int C>
struct Layout {
    static constexpr int kWarmDoubles = WR_HEADER + (2 * WR_SPEED_FIELDS + WR_CTRL_FIELDS) * C * 32;
    // C = 3 (horizons 65..96, e.g. the Spa H = 80 sweep): with the C = 2 recipe one instance needs 384 tensor-memory
    // columns (-> a 512-column allocation: ONE CTA per SM) and 33 KB of shared memory.  The "split" layout brings it back
    // to two CTAs = 8 instances per SM: the cold per-stage fields (read at termination checks, refactorisations and for
    // the outputs only) live in a per-warp slice of GLOBAL memory (L2-resident), tensor memory keeps chunks F, G of all
    // three stages and H of the first two (128 doubles per lane = 256 columns), and H of the third stage plus the three NS
    // chunks sit in shared memory as lane-private columns next to the scan matrices, which drop their zero padding.
    static constexpr bool kSplit = (C == 3);
    static constexpr bool kColdGlobal = kSplit;
    static constexpr int kColdDoubles = K_FIELDS * C * 32;          // cold per-stage fields of one instance
    // scan matrices: every element keeps 48 lanes; lanes 32..47 stay zero, so the backward scan reads
    // "lane + 2^L" without a bounds test and whatever its raw shuffle delivers there is cancelled
    // (C = 1 and the split layout keep 32 lanes and the bounds test: the padding would cost them a CTA per SM)
    static constexpr int kScanLanes = (C == 1 || kSplit) ? 32 : 48;
    static constexpr int kScanDoubles = 9 * kLevels * kScanLanes;
    static constexpr int kScratch = (W_FIELDS * C * 32 > kScanDoubles) ? W_FIELDS * C * 32 : kScanDoubles;
    static constexpr int kHotSmemChunks = kSplit ? 4 : 0;           // 16-double chunks per lane kept in shared memory
    static constexpr int kHotSmemDoubles = kHotSmemChunks * 16 * 32;
    // shared memory per warp, control phase: [cold fields unless global | scratch / scan region | hot chunks]
    static constexpr int kDoubles = (kColdGlobal ? 0 : kColdDoubles) + kScratch + kHotSmemDoubles;
    static constexpr int kSpeedDoubles = (2 * C * 32 > 3 * 32 * C) ? 2 * C * 32 : 3 * 32 * C;   // speed phase: raw path / LDL work
    static constexpr int kTmemDoubles = kSplit ? 128 : T_STRIDE * C;   // tensor memory per lane
    static constexpr int kTmemCols = (2 * kTmemDoubles <= 128) ? 128 : ((2 * kTmemDoubles <= 256) ? 256 : 512);
    // where chunk `ch` (T_F, T_G, T_H, T_NS) of the lane's stage j lives
    AC_HD static constexpr bool in_smem(int ch, int j) { return kSplit && (ch == T_NS || (ch == T_H && j == 2)); }
    AC_HD static constexpr int smem_slot(int ch, int j) { return ch == T_NS ? j : 3; }
    AC_HD static constexpr int tmem_off(int ch, int j)
    {
        return !kSplit ? j * T_STRIDE + ch : (ch == T_H ? 96 + 16 * j : 32 * j + (ch == T_G ? 16 : 0));
    }
};
template <int C>
constexpr int smem_doubles()
{
    return Layout<C>::kDoubles;
}

AC_DEV double limit_scaling_u(double v)
{
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}
AC_DEV VD limit_scaling(const VD& v)
{
    VD t = vsel(v < VD(kMinScaling), VD(1.0), v);
    return vsel(t > VD(kMaxScaling), VD(kMaxScaling), t);
}
AC_DEV double clampu(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
AC_DEV VD vclamp(const VD& v, const VD& lo, const VD& hi) { return vsel(v < lo, lo, vsel(v > hi, hi, v)); }
AC_DEV VD inv_sqrt(const VD& v) { return vrsqrt(v); }

// numpy's mod for a positive modulus
AC_DEV VD np_mod(const VD& a, double b)
{
    VD r = vfmod(a, b);
    return vsel(r < VD(0.0), r + VD(b), r);
}
AC_DEV double np_mod_u(double a, double b)
{
    double r = fmod(a, b);
    return r < 0.0 ? r + b : r;
}

// OSQP constraint classes (set_rho_vec): 0 inequality, 1 equality, 2 loose
AC_DEV VI row_class(const VD& l, const VD& u)
{
    VB loose = (l < VD(-kBig)) & (u > VD(kBig));
    VB eq = (u - l) < VD(kRhoTol);
    return vseli(loose, vi_all(2), vseli(eq, vi_all(1), vi_all(0)));
}

struct RhoSet {
    double rho, rho_eq, rinv, rinv_eq;
    AC_MEM void set(double r)
    {
        rho = r;
        rho_eq = kRhoEqOverIneq * r;
        rinv = 1.0 / rho;
        rinv_eq = 1.0 / rho_eq;
    }
    AC_MEM VD of(const VI& cls, int sh) const
    {
        return vsel(vi_field_is(cls, sh, 0), VD(rho), vsel(vi_field_is(cls, sh, 1), VD(rho_eq), VD(kRhoMin)));
    }
    AC_MEM VD inv_of(const VI& cls, int sh) const
    {
        return vsel(vi_field_is(cls, sh, 0), VD(rinv), vsel(vi_field_is(cls, sh, 1), VD(rinv_eq), VD(1.0 / kRhoMin)));
    }
};

// norms gathered at a termination check (update_info + compute_*_tol + compute_rho_estimate)
struct Norms {
    double pri, dua, nz, nAx, nq, nAty, nPx;               // unscaled (termination)
    double s_pri, s_dua, s_z, s_Ax, s_q, s_Aty, s_Px;      // scaled (rho estimate)
    double xtPx, qtx, sc;                                  // scaled x'Px, q'x, SC(y): OSQP 1.x duality-gap test only
};

struct SolveInfo {
    int status, iter, rho_updates;
    double pri_res, dua_res, obj_val;
};

AC_DEV double rho_estimate(const Norms& N, double rho)
{
    double p = N.s_pri / (fmax(N.s_z, N.s_Ax) + 1e-10);
    double d = N.s_dua / (fmax(N.s_q, fmax(N.s_Aty, N.s_Px)) + 1e-10);
    double r = rho * sqrt(p / (d + 1e-10));
    return clampu(r, kRhoMin, kRhoMax);
}

// OSQP overwrites info.obj_val when a certificate fires (check_termination): +-OSQP_INFTY, NaN if non-convex
AC_DEV double final_obj(int status, double obj)
{
    if (status == ACMPC_PRIMAL_INFEASIBLE || status == ACMPC_PRIMAL_INFEASIBLE_INACCURATE) return kInfty;
    if (status == ACMPC_DUAL_INFEASIBLE || status == ACMPC_DUAL_INFEASIBLE_INACCURATE) return -kInfty;
    if (status == ACMPC_NON_CVX) return NAN;
    return obj;
}

// the 12 max-norms of one termination check -> Norms
AC_DEV void finish_norms(VD (&v)[12], double cinv, double nq_unscaled, double nq_scaled, Norms& N)
{
    double m[12];
    for (int t = 0; t < 12; ++t) m[t] = wmax(v[t]);
    N.pri = m[0], N.nz = m[1], N.nAx = m[2];
    N.dua = cinv * m[3], N.nAty = m[4], N.nPx = m[5], N.nq = nq_unscaled;
    N.s_dua = m[6], N.s_pri = m[7], N.s_z = m[8], N.s_Ax = m[9], N.s_Aty = m[10], N.s_Px = m[11];
    N.s_q = nq_scaled;
}
AC_DEV void acc_row(VD (&v)[12], const VD& ax, const VD& z, const VD& einv)
{
    VD res = ax - z;
    v[0] = vmax(v[0], vabs(einv * res)), v[7] = vmax(v[7], vabs(res));
    v[1] = vmax(v[1], vabs(einv * z)), v[8] = vmax(v[8], vabs(z));
    v[2] = vmax(v[2], vabs(einv * ax)), v[9] = vmax(v[9], vabs(ax));
}
AC_DEV void acc_col(VD (&v)[12], const VD& q, const VD& px, const VD& aty, const VD& dinv)
{
    VD dr = q + px + aty;
    v[3] = vmax(v[3], vabs(dinv * dr)), v[6] = vmax(v[6], vabs(dr));
    v[4] = vmax(v[4], vabs(dinv * aty)), v[10] = vmax(v[10], vabs(aty));
    v[5] = vmax(v[5], vabs(dinv * px)), v[11] = vmax(v[11], vabs(px));
}
// is_primal_infeasible's projection of delta_y by bound type
AC_DEV VD project_dy(const VD& dy, const VD& lo, const VD& hi)
{
    VB up_inf = hi > VD(kBig), lo_inf = lo < VD(-kBig);
    VD a = vsel(lo_inf, VD(0.0), vmin(dy, VD(0.0)));   // upper infinite
    VD b = vsel(lo_inf, vmax(dy, VD(0.0)), dy);        // upper finite
    return vsel(up_inf, a, b);
}
AC_DEV VD support(const VD& dy, const VD& lo, const VD& hi)
{
    return hi * vmax(dy, VD(0.0)) + lo * vmin(dy, VD(0.0));
}
// SC(y) of OSQP 1.x's duality gap: the same support function over the FINITE bounds only
AC_DEV VD support_finite(const VD& y, const VD& lo, const VD& hi)
{
    return vsel(hi < VD(kBig), hi * vmax(y, VD(0.0)), VD(0.0)) + vsel(lo > VD(-kBig), lo * vmin(y, VD(0.0)), VD(0.0));
}
// OSQP 1.x check_dualgap: |x'Px + q'x + SC(y)| < eps_abs + eps_rel max(|x'Px|, |q'x|, |SC(y)|), all unscaled by 1/c
AC_DEV bool dualgap_ok(const Norms& N, double cinv, double eps_abs, double eps_rel)
{
    const double gap = cinv * (N.xtPx + N.qtx + N.sc);
    const double rel = cinv * fmax(fabs(N.xtPx), fmax(fabs(N.qtx), fabs(N.sc)));
    return fabs(gap) < eps_abs + eps_rel * rel;
}

// ------------------------------------------------------------------------------------------------
// per-instance context
// ------------------------------------------------------------------------------------------------
// Phase clocks (instrumented experiment builds only, -DACMPC_PHASE_TIMING: tools/phase_timing.py): lane 0 adds the
// cycles since the previous mark to g_phase[k].  Compiled out of the product.
#if defined(ACMPC_PHASE_TIMING) && !defined(ACMPC_EMULATE)
__device__ unsigned long long g_phase[16];
#define AC_PHASE(c, k)                                                          \
    do {                                                                        \
        const long long t_ = clock64();                                         \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase[k], (unsigned long long)(t_ - (c).tl)); \
        (c).tl = clock64();                                                     \
    } while (0)
#else
#define AC_PHASE(c, k) ((void)0)
#endif

template <int C>
struct Ctx {
    mutable long long tl;
    double* S;   // this instance's (warp's) cold per-stage fields: shared memory, or global memory in the split layout
    double* W;   // this instance's scratch / scan region in shared memory
    double* HS;  // split layout: the hot chunks kept in shared memory (lane-private columns)
    Tm tm;       // base of this instance's tensor-memory block (the warp's lane quarter; control phase only)
    int H, n;
    const acmpc_config* cfg;
    VI lane;
    AC_MEM double* col(int f, int j) const { return S + (f * C + j) * 32; }
    AC_MEM VD ld(int f, int j) const { return ld_lane(col(f, j)); }
    AC_MEM void st(int f, int j, const VD& v) const { st_lane(col(f, j), v); }
    AC_MEM double* scratch() const { return W; }
    AC_MEM double* wcol(int f, int j) const { return scratch() + (f * C + j) * 32; }
    // element e of the level-L scan matrices: scan(L, e)[lane], lanes 0 .. 47
    AC_MEM double* scan(int lvl, int e) const { return scratch() + (lvl * 9 + e) * Layout<C>::kScanLanes; }
    AC_MEM VI stage(int j) const { return lane * C + j; }
    // hot per-stage data: N doubles at offset `off` of chunk `ch` (T_F, T_G, T_H, T_NS) of own stage j -- from tensor
    // memory or, for the chunks the split layout keeps there, from the lane's shared-memory column.  ch and j are
    // compile-time constants at every call site (unrolled loops), so the routing folds away.
    template <int N>
    AC_MEM void hld(int ch, int off, int j, VD (&o)[N]) const
    {
        if (Layout<C>::in_smem(ch, j)) {
            const double* p = HS + (Layout<C>::smem_slot(ch, j) * 16 + off) * 32;
            AC_UNROLL
            for (int k = 0; k < N; ++k) o[k] = ld_lane(p + k * 32);
        } else {
            tm_ld<N>(tm, Layout<C>::tmem_off(ch, j) + off, o);
        }
    }
    AC_MEM VD hld1(int ch, int off, int j) const
    {
        if (Layout<C>::in_smem(ch, j)) return ld_lane(HS + (Layout<C>::smem_slot(ch, j) * 16 + off) * 32);
        return tm_ld1(tm, Layout<C>::tmem_off(ch, j) + off);
    }
    AC_MEM void tld(int ch, int j, VD (&o)[16]) const { hld<16>(ch, 0, j, o); }
    // several pieces of stage j at once; when all of them live in tensor memory the loads share ONE wait
    template <int N0, int N1, int N2, int N3>
    AC_MEM void hld4(int j, int c0, int f0, VD (&o0)[N0], int c1, int f1, VD (&o1)[N1], int c2, int f2, VD (&o2)[N2], int c3,
                     int f3, VD (&o3)[N3]) const
    {
        using L = Layout<C>;
        if (L::in_smem(c0, j) || L::in_smem(c1, j) || L::in_smem(c2, j) || L::in_smem(c3, j)) {
            hld<N0>(c0, f0, j, o0), hld<N1>(c1, f1, j, o1), hld<N2>(c2, f2, j, o2), hld<N3>(c3, f3, j, o3);
        } else {
            tm_ld_group4<N0, N1, N2, N3>(tm, L::tmem_off(c0, j) + f0, o0, L::tmem_off(c1, j) + f1, o1, L::tmem_off(c2, j) + f2, o2,
                                         L::tmem_off(c3, j) + f3, o3);
        }
    }
    AC_MEM void tst(int ch, int j, const VD (&v)[16]) const
    {
        if (Layout<C>::in_smem(ch, j)) {
            double* p = HS + Layout<C>::smem_slot(ch, j) * 16 * 32;
            AC_UNROLL
            for (int k = 0; k < 16; ++k) st_lane(p + k * 32, v[k]);
        } else {
            tm_st<16>(tm, Layout<C>::tmem_off(ch, j), v);
        }
    }
};

// value held by the NEXT stage (s+1) for every own stage; 0 beyond lane 31
template <int C>
AC_DEV void pull_next(const VD (&v)[C], VD (&o)[C])
{
    VD nx = shfl_down0(v[0], 1);
    AC_UNROLL
    for (int j = 0; j + 1 < C; ++j) o[j] = v[j + 1];
    o[C - 1] = nx;
}
// value held by the PREVIOUS stage (s-1); 0 before lane 0
template <int C>
AC_DEV void pull_prev(const VD (&v)[C], VD (&o)[C])
{
    VD pv = shfl_up0(v[C - 1], 1);
    AC_UNROLL
    for (int j = C - 1; j >= 1; --j) o[j] = v[j - 1];
    o[0] = pv;
}
// scalar hot-loop variants (see pull_next_raw / pull_prev_rot below)
template <int C>
AC_DEV void pull_next_raw1(const VD (&v)[C], VD (&o)[C])
{
    VD nx = shfl_down_raw(v[0], 1);
    AC_UNROLL
    for (int j = 0; j + 1 < C; ++j) o[j] = v[j + 1];
    o[C - 1] = nx;
}
template <int C>
AC_DEV void pull_prev_rot1(const VD (&v)[C], VD (&o)[C])
{
    VD pv = shfl_rot_up1(v[C - 1]);
    AC_UNROLL
    for (int j = C - 1; j >= 1; --j) o[j] = v[j - 1];
    o[0] = pv;
}
template <int C, int K>
AC_DEV void pull_next_k(const VD (&v)[C][K], VD (&o)[C][K])
{
    AC_UNROLL
    for (int e = 0; e < K; ++e) {
        VD nx = shfl_down0(v[0][e], 1);
        AC_UNROLL
        for (int j = 0; j + 1 < C; ++j) o[j][e] = v[j + 1][e];
        o[C - 1][e] = nx;
    }
}
template <int C, int K>
AC_DEV void pull_prev_k(const VD (&v)[C][K], VD (&o)[C][K])
{
    AC_UNROLL
    for (int e = 0; e < K; ++e) {
        VD pv = shfl_up0(v[C - 1][e], 1);
        AC_UNROLL
        for (int j = C - 1; j >= 1; --j) o[j][e] = v[j - 1][e];
        o[0][e] = pv;
    }
}

// hot-loop variants without the out-of-range select: the consumer multiplies what lane 31 "pulls" by the
// (zero) A of its last stage; lane 0 pulls the (zero) A-product of lane 31's last stage by rotation.
template <int C, int K>
AC_DEV void pull_next_raw(const VD (&v)[C][K], VD (&o)[C][K])
{
    AC_UNROLL
    for (int e = 0; e < K; ++e) {
        VD nx = shfl_down_raw(v[0][e], 1);
        AC_UNROLL
        for (int j = 0; j + 1 < C; ++j) o[j][e] = v[j + 1][e];
        o[C - 1][e] = nx;
    }
}
template <int C, int K>
AC_DEV void pull_prev_rot(const VD (&v)[C][K], VD (&o)[C][K])
{
    AC_UNROLL
    for (int e = 0; e < K; ++e) {
        VD pv = shfl_rot_up1(v[C - 1][e]);
        AC_UNROLL
        for (int j = C - 1; j >= 1; --j) o[j][e] = v[j - 1][e];
        o[0][e] = pv;
    }
}

// o += M v  /  o += M' v   (M row major 3x3, 9 consecutive values), written as FMA chains
AC_DEV void mv_acc9(const VD* M, const VD (&v)[3], VD (&o)[3])
{
    AC_UNROLL
    for (int r = 0; r < 3; ++r) {
        o[r] = M[3 * r] * v[0] + o[r];
        o[r] = M[3 * r + 1] * v[1] + o[r];
        o[r] = M[3 * r + 2] * v[2] + o[r];
    }
}
AC_DEV void mtv_acc9(const VD* M, const VD (&v)[3], VD (&o)[3])
{
    AC_UNROLL
    for (int r = 0; r < 3; ++r) {
        o[r] = M[r] * v[0] + o[r];
        o[r] = M[3 + r] * v[1] + o[r];
        o[r] = M[6 + r] * v[2] + o[r];
    }
}

// ReferencePath rows of the lane's stages, in registers while the QPs are assembled
template <int C>
struct PathRegs {
    VD xs[C], ys[C], psi[C], kap[C], dist[C], wid[C];
};

// spatial_mpc.py:125-154.  `W` = raw (H,3) path staged in shared memory.
template <int C>
AC_DEV void build_waypoints(const Ctx<C>& c, const double* W, PathRegs<C>& p)
{
    const int n = c.n, H = c.H;
    AC_UNROLL
    for (int j = 0; j < C; ++j) {
        VI s = c.stage(j);
        VB ok = vi_lt(s, n);
        VI ic = s * 3, in = ic + 3;
        VI ip = vseli(vi_eq(s, 0), vi_all(3 * (H - 1)), ic + (-3));
        VD cx = ld_idx_if(ok, W, ic, 0.0), cy = ld_idx_if(ok, W, ic + 1, 0.0);
        VD nx = ld_idx_if(ok, W, in, 0.0), ny = ld_idx_if(ok, W, in + 1, 0.0), nw = ld_idx_if(ok, W, in + 2, 0.0);
        VD px = ld_idx_if(ok, W, ip, 0.0), py = ld_idx_if(ok, W, ip + 1, 0.0);
        VD ax = nx - cx, ay = ny - cy, bx = cx - px, by = cy - py;
        VD ps = vatan2(ay, ax);
        VD d = vsqrt(ax * ax + ay * ay);
        VD behind = vatan2(by, bx);
        VD dang = np_mod(ps - behind + VD(kPi), 2.0 * kPi) - VD(kPi);
        p.xs[j] = cx, p.ys[j] = cy, p.wid[j] = nw;
        p.psi[j] = vsel(ok, ps, VD(0.0));
        p.dist[j] = vsel(ok, d, VD(0.0));
        p.kap[j] = vsel(ok, dang / (d + VD(1e-12)) + VD(1e-12), VD(0.0));
    }
    // kappa_0 := kappa_1
    VD kn[C];
    pull_next<C>(p.kap, kn);
    p.kap[0] = vsel(vi_eq(c.stage(0), 0), kn[0], p.kap[0]);
}

// ================================================================================================
// Speed-profile QP   min 1/2 v'v - vbar'v   s.t.  a_min <= (v_{i+1}-v_i)/(2 d_i) <= a_max,
//                                                  v_min <= v_i <= vbar_i          (speed_profile.py)
// Stage i owns v_i, its bound row and acceleration row i (i <= n-2).  All state in registers.
// ================================================================================================
template <int C>
struct SpeedQP {
    const Ctx<C>& c;
    int n;
    VD al[C], au[C], ss[C], p[C], q[C], la[C], ua[C], lb[C], ub[C], di[C], eai[C], ebi[C];
    VD rho_a[C], rinv_a[C], rho_b[C], rinv_b[C];
    VI cls[C];   // bits 0-1: acceleration row, bits 2-3: bound row
    VD x[C], za[C], zb[C], ya[C], yb[C];
    VD dx[C], dya[C], dyb[C];
    VD nl[C], dinv[C], phi[kLevels], psi[kLevels];
    double cs, cinv, nq_unscaled, nq_scaled;
    RhoSet R;

    AC_MEM explicit SpeedQP(const Ctx<C>& ctx) : c(ctx), n(ctx.n) {}

    // speed_profile.py:26-59 / :131-150, then OSQP scale_data + set_rho_vec
    AC_MEM void assemble_and_scale(const PathRegs<C>& path, double v_max_live, int localised)
    {
        const acmpc_config& g = *c.cfg;
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI s = c.stage(j);
            VB ok = vi_lt(s, n), row = vi_le(s, n - 2);
            VD ak = vabs(path.kap[j]);
            VD vdyn = vsqrt(VD(g.ay_max) / (ak + VD(1e-12)));
            vdyn = vsel(ak < VD(g.ki_min), VD(v_max_live), vdyn);
            VD v = vsel(vdyn < VD(v_max_live), vdyn, VD(v_max_live));
            v = vsel(VD(g.v_min) > v, VD(g.v_min), v);
            VD vb = v + VD(2.0);
            if (g.has_end_velocity) vb = vsel(vi_eq(s, n - 1), VD(g.end_velocity), vb);
            if (localised) vb = VD(v_max_live);
            VD h = VD(1.0) / (VD(2.0) * path.dist[j]);
            al[j] = vsel(row, -h, VD(0.0));
            au[j] = vsel(row, h, VD(0.0));
            ss[j] = vsel(ok, VD(1.0), VD(0.0));
            p[j] = vsel(ok, VD(1.0), VD(0.0));
            q[j] = vsel(ok, VD(-1.0) * vb, VD(0.0));
            di[j] = VD(1.0), eai[j] = VD(1.0), ebi[j] = VD(1.0);
            la[j] = vsel(row, VD(g.a_min), VD(0.0)), ua[j] = vsel(row, VD(g.a_max), VD(0.0));
            lb[j] = vsel(ok, VD(g.v_min), VD(0.0)), ub[j] = vsel(ok, vb, VD(0.0));
        }
        cs = 1.0;
        for (int pass = 0; pass < g.scaling; ++pass) {
            // column / row inf-norms of [P A'; A 0]
            VD aua[C], aup[C], d[C], dn[C];
            AC_UNROLL
            for (int j = 0; j < C; ++j) aua[j] = vabs(au[j]);
            pull_prev<C>(aua, aup);
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                VD m = vmax(vmax(vabs(p[j]), vabs(ss[j])), vmax(vabs(al[j]), aup[j]));
                d[j] = inv_sqrt(limit_scaling(m));
            }
            pull_next<C>(d, dn);
            VD psum = VD(0.0), qmax = VD(0.0);
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                VD ea = inv_sqrt(limit_scaling(vmax(vabs(al[j]), vabs(au[j]))));
                al[j] = (al[j] * ea) * d[j];
                au[j] = (au[j] * ea) * dn[j];
                eai[j] = eai[j] * ea;
                VD eb = inv_sqrt(limit_scaling(vabs(ss[j])));
                ss[j] = (ss[j] * eb) * d[j];
                ebi[j] = ebi[j] * eb;
                p[j] = (p[j] * d[j]) * d[j];
                q[j] = q[j] * d[j];
                di[j] = di[j] * d[j];
                psum = psum + vabs(p[j]);
                qmax = vmax(qmax, vabs(q[j]));
            }
            double ct = fmax(wsum(psum) / (double)n, limit_scaling_u(wmax(qmax)));
            ct = 1.0 / limit_scaling_u(ct);
            AC_UNROLL
            for (int j = 0; j < C; ++j) p[j] = p[j] * VD(ct), q[j] = q[j] * VD(ct);
            cs *= ct;
        }
        cinv = 1.0 / cs;
        // scale bounds, classify rows, invert D/E, norm of q
        VD nqu = VD(0.0), nqs = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            la[j] = la[j] * eai[j], ua[j] = ua[j] * eai[j];
            lb[j] = lb[j] * ebi[j], ub[j] = ub[j] * ebi[j];
            cls[j] = row_class(la[j], ua[j]) | (row_class(lb[j], ub[j]) << 2);
            di[j] = VD(1.0) / di[j];
            eai[j] = VD(1.0) / eai[j];
            ebi[j] = VD(1.0) / ebi[j];
            nqu = vmax(nqu, vabs(di[j] * q[j]));
            nqs = vmax(nqs, vabs(q[j]));
        }
        nq_unscaled = wmax(nqu);
        nq_scaled = wmax(nqs);
    }

    // reduced matrix K = P + sigma + A' rho A (tridiagonal) -> LDL' -> scan coefficients.
    // scratch: two per-stage work columns of shared memory
    AC_MEM void factor()
    {
        const double sigma = c.cfg->sigma;
        VD t[C], tp[C];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            rho_a[j] = R.of(cls[j], 0), rinv_a[j] = R.inv_of(cls[j], 0);
            rho_b[j] = R.of(cls[j], 2), rinv_b[j] = R.inv_of(cls[j], 2);
            t[j] = rho_a[j] * au[j] * au[j];
        }
        pull_prev<C>(t, tp);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD d = p[j] + VD(sigma) + rho_b[j] * ss[j] * ss[j] + rho_a[j] * al[j] * al[j] + tp[j];
            st_lane(c.wcol(0, j), d);
            st_lane(c.wcol(1, j), rho_a[j] * al[j] * au[j]);   // (i, i+1) entry, 0 without a row
        }
        warp_sync();
        // serial LDL' (every lane runs it on broadcast reads; lane 0 publishes)
        {
            double oprev = 0.0, dprev = 0.0;
            int s = 0;
            for (int l = 0; l < 32 && s < n; ++l)
                for (int j = 0; j < C && s < n; ++j, ++s) {
                    double* pd = c.wcol(0, j) + l;
                    double* po = c.wcol(1, j) + l;
                    double dd = *pd, oo = *po;
                    double lw = oprev * dprev;          // L_{s,s-1}
                    double piv = dd - lw * oprev;
                    dprev = urecip(piv);
                    oprev = oo;
                    AC_LANE0
                    {
                        *pd = dprev;
                        *po = -lw;
                    }
                }
        }
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VB ok = vi_lt(c.stage(j), n);
            dinv[j] = vsel(ok, ld_lane(c.wcol(0, j)), VD(0.0));
            nl[j] = vsel(ok, ld_lane(c.wcol(1, j)), VD(0.0));
        }
        warp_sync();
        VD f = nl[0];
        AC_UNROLL
        for (int j = 1; j < C; ++j) f = nl[j] * f;
        AC_UNROLL
        for (int L = 0; L < kLevels; ++L) {
            phi[L] = f;
            psi[L] = shfl_down0(f, 1 << L);
            f = f * shfl_up0(f, 1 << L);
        }
    }

    // x~ = K^{-1} r
    AC_MEM void kkt_solve(const VD (&r)[C], VD (&xt)[C])
    {
        VD y[C], w[C];
        VD Y = r[0];
        AC_UNROLL
        for (int j = 1; j < C; ++j) Y = r[j] + nl[j] * Y;
        AC_UNROLL
        for (int L = 0; L < kLevels; ++L) Y = Y + phi[L] * shfl_up_raw(Y, 1 << L);   // phi = 0 on lanes < 2^L
        VD yin = shfl_up_raw(Y, 1);                                                  // nl[0] = 0 on lane 0
        y[0] = r[0] + nl[0] * yin;
        AC_UNROLL
        for (int j = 1; j + 1 < C; ++j) y[j] = r[j] + nl[j] * y[j - 1];
        if (C > 1) y[C - 1] = Y;
        AC_UNROLL
        for (int j = 0; j < C; ++j) w[j] = y[j] * dinv[j];
        VD a = VD(0.0);
        if (C > 1) {
            a = w[C - 2];
            AC_UNROLL
            for (int j = C - 3; j >= 0; --j) a = w[j] + nl[j + 1] * a;
        }
        VD Z = w[C - 1] + shfl_down0(nl[0] * a, 1);
        AC_UNROLL
        for (int L = 0; L < kLevels; ++L) Z = Z + psi[L] * shfl_down_raw(Z, 1 << L);   // psi = 0 beyond lane 31
        xt[C - 1] = Z;
        AC_UNROLL
        for (int j = C - 2; j >= 0; --j) xt[j] = w[j] + nl[j + 1] * xt[j + 1];
    }

    AC_MEM void compute_norms(Norms& N)
    {
        VD v[12];
        for (int t = 0; t < 12; ++t) v[t] = VD(0.0);
        VD xn[C], t[C], tp[C];
        pull_next<C>(x, xn);
        AC_UNROLL
        for (int j = 0; j < C; ++j) t[j] = au[j] * ya[j];
        pull_prev<C>(t, tp);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD aty = ss[j] * yb[j] + al[j] * ya[j] + tp[j];
            acc_row(v, al[j] * x[j] + au[j] * xn[j], za[j], eai[j]);
            acc_row(v, ss[j] * x[j], zb[j], ebi[j]);
            acc_col(v, q[j], p[j] * x[j], aty, di[j]);
        }
        finish_norms(v, cinv, nq_unscaled, nq_scaled, N);
        if (c.cfg->check_dualgap) {
            VD a = VD(0.0), b = VD(0.0), s = VD(0.0);
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                a = a + p[j] * x[j] * x[j];
                b = b + q[j] * x[j];
                s = s + support_finite(ya[j], la[j], ua[j]) + support_finite(yb[j], lb[j], ub[j]);
            }
            N.xtPx = wsum(a), N.qtx = wsum(b), N.sc = wsum(s);
        }
    }

    AC_MEM int primal_infeasible(double eps)
    {
        VD nrm = VD(0.0), lhs = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            dya[j] = project_dy(dya[j], la[j], ua[j]);
            dyb[j] = project_dy(dyb[j], lb[j], ub[j]);
            nrm = vmax(nrm, vmax(vabs(dya[j] / eai[j]), vabs(dyb[j] / ebi[j])));
            lhs = lhs + support(dya[j], la[j], ua[j]) + support(dyb[j], lb[j], ub[j]);
        }
        double nr = wmax(nrm), lh = wsum(lhs);
        if (uni(!(nr > eps) || !(lh < -eps * nr))) return 0;
        VD t[C], tp[C], m = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) t[j] = au[j] * dya[j];
        pull_prev<C>(t, tp);
        AC_UNROLL
        for (int j = 0; j < C; ++j) m = vmax(m, vabs(di[j] * (ss[j] * dyb[j] + al[j] * dya[j] + tp[j])));
        return wmax(m) < eps * nr;
    }

    AC_MEM int dual_infeasible(double eps)
    {
        VD nrm = VD(0.0), qdx = VD(0.0), pm = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            nrm = vmax(nrm, vabs(dx[j] / di[j]));
            qdx = qdx + q[j] * dx[j];
            pm = vmax(pm, vabs(di[j] * (p[j] * dx[j])));
        }
        double nr = wmax(nrm), qd = wsum(qdx), pmx = wmax(pm);
        if (uni(!(nr > eps) || !(qd < -cs * eps * nr) || !(pmx < cs * eps * nr))) return 0;
        VD dn[C];
        pull_next<C>(dx, dn);
        VB bad = vb_all(false);
        const VD lim = VD(eps * nr);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD a = eai[j] * (al[j] * dx[j] + au[j] * dn[j]);
            bad = bad | ((ua[j] < VD(kBig)) & (a > lim)) | ((la[j] > VD(-kBig)) & (a < -lim));
            VD b = ebi[j] * (ss[j] * dx[j]);
            bad = bad | ((ub[j] < VD(kBig)) & (b > lim)) | ((lb[j] > VD(-kBig)) & (b < -lim));
        }
        return !wany(bad);
    }

    // check_termination(work, approximate): returns status or 0 (continue)
    AC_MEM int check(const Norms& N, int approximate)
    {
        const acmpc_config& g = *c.cfg;
        double k = approximate ? 10.0 : 1.0;
        if (uni(N.pri > kInfty || N.dua > kInfty)) return ACMPC_NON_CVX;
        double eps_p = k * g.eps_abs + k * g.eps_rel * fmax(N.nz, N.nAx);
        double eps_d = k * g.eps_abs + k * g.eps_rel * cinv * fmax(N.nq, fmax(N.nAty, N.nPx));
        int p_ok = uni(N.pri < eps_p), d_ok = uni(N.dua < eps_d);
        int p_inf = 0, d_inf = 0;
        if (!p_ok) p_inf = primal_infeasible(k * g.eps_prim_inf);
        if (!d_ok) d_inf = dual_infeasible(k * g.eps_dual_inf);
        const int g_ok = g.check_dualgap ? uni(dualgap_ok(N, cinv, k * g.eps_abs, k * g.eps_rel)) : 1;
        if (p_ok && d_ok && g_ok) return approximate ? ACMPC_SOLVED_INACCURATE : ACMPC_SOLVED;
        if (p_inf) return approximate ? ACMPC_PRIMAL_INFEASIBLE_INACCURATE : ACMPC_PRIMAL_INFEASIBLE;
        if (d_inf) return approximate ? ACMPC_DUAL_INFEASIBLE_INACCURATE : ACMPC_DUAL_INFEASIBLE;
        return 0;
    }

    // one ADMM iteration (osqp_solve loop body); KEEP: also record delta_x / delta_y for the certificates
    template <bool KEEP>
    AC_MEM void iterate(const double alpha, const double sigma)
    {
        VD r[C], xt[C], xn[C], t[C], tp[C];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD wa = rho_a[j] * za[j] - ya[j];
            t[j] = au[j] * wa;
            r[j] = VD(sigma) * x[j] - q[j] + ss[j] * (rho_b[j] * zb[j] - yb[j]) + al[j] * wa;
        }
        pull_prev_rot1<C>(t, tp);   // au = 0 on the last stage: lane 0 pulls a zero
        AC_UNROLL
        for (int j = 0; j < C; ++j) r[j] = r[j] + tp[j];
        kkt_solve(r, xt);
        pull_next_raw1<C>(xt, xn);   // multiplied by au = 0 beyond the last row
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            {
                VD zt = al[j] * xt[j] + au[j] * xn[j];
                VD zh = VD(alpha) * zt + VD(1.0 - alpha) * za[j];
                VD zn = vclamp(zh + rinv_a[j] * ya[j], la[j], ua[j]);
                VD dy = rho_a[j] * (zh - zn);
                za[j] = zn;
                ya[j] = ya[j] + dy;
                if (KEEP) dya[j] = dy;
            }
            {
                VD zt = ss[j] * xt[j];
                VD zh = VD(alpha) * zt + VD(1.0 - alpha) * zb[j];
                VD zn = vclamp(zh + rinv_b[j] * yb[j], lb[j], ub[j]);
                VD dy = rho_b[j] * (zh - zn);
                zb[j] = zn;
                yb[j] = yb[j] + dy;
                if (KEEP) dyb[j] = dy;
            }
            VD xnew = VD(alpha) * xt[j] + VD(1.0 - alpha) * x[j];
            if (KEEP) dx[j] = xnew - x[j];
            x[j] = xnew;
        }
    }

    // osqp_solve, cold start.  Result (unscaled v) -> vout.
    // `warm` = this instance's warm-start record or nullptr; `slot` 0 / 1 = the (un)localised solver object.
    // `use_warm`: start from the record when its slot is valid (OSQP warm_start = True after update()).
    AC_MEM void solve(SolveInfo& info, VD (&vout)[C], double* warm, int slot, bool use_warm)
    {
        const acmpc_config& g = *c.cfg;
        const double alpha = g.alpha, sigma = g.sigma;
        double* wrow = warm ? warm + WR_HEADER + slot * WR_SPEED_FIELDS * C * 32 : nullptr;
        const bool have = warm && use_warm && warm[WR_VALID + slot] != 0.0;
        R.set(have ? warm[WR_RHO + slot] : clampu(g.rho, kRhoMin, kRhoMax));
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            x[j] = za[j] = zb[j] = ya[j] = yb[j] = VD(0.0);
            dx[j] = dya[j] = dyb[j] = VD(0.0);
            if (have) {
                x[j] = ld_lane(wrow + (0 * C + j) * 32), za[j] = ld_lane(wrow + (1 * C + j) * 32);
                zb[j] = ld_lane(wrow + (2 * C + j) * 32), ya[j] = ld_lane(wrow + (3 * C + j) * 32);
                yb[j] = ld_lane(wrow + (4 * C + j) * 32);
            }
        }
        warp_sync();
        factor();
        AC_PHASE(c, 10);   // first factorisation
        Norms N;
        int status = 0, iter = 0, updates = 0;
        // osqp_solve's order: iterate -> exact check (every check_termination iterations) -> adaptive rho;
        // at the iteration limit the exact check if it has not just run, then the 10x "inaccurate" one.
        // Iterations followed by a check / rho estimate are the KEEP instantiation and the check sits in the
        // same block, so the delta vectors of the certificates never live across the loop; everything in the
        // block is straight-line code (a loop over "phases" here once let LICM hoist rho_estimate's
        // divisions into every iteration).
        bool checked = false, adapt = false, last = false;
        // countdowns instead of iter % interval (a runtime integer division per test)
        int to_check = g.check_termination > 0 ? g.check_termination : -1;
        int to_adapt = (g.adaptive_rho && g.adaptive_rho_interval > 0) ? g.adaptive_rho_interval : -1;
        for (;;) {
            ++iter;
            checked = (--to_check == 0);
            if (checked) to_check = g.check_termination;
            adapt = (--to_adapt == 0);
            if (adapt) to_adapt = g.adaptive_rho_interval;
            last = iter >= g.max_iter;
            if (checked || adapt || last) {
                iterate<true>(alpha, sigma);
                AC_PHASE(c, 11);   // iterations
                compute_norms(N);
                if (checked) status = check(N, 0);
                if (uni(status == 0) && adapt) {
                    double rn = rho_estimate(N, R.rho);
                    if (uni(rn > R.rho * g.adaptive_rho_tolerance || rn < R.rho / g.adaptive_rho_tolerance)) {
                        R.set(rn);
                        ++updates;
                        factor();
                    }
                }
                if (uni(status == 0) && last) {
                    if (!checked) status = check(N, 0);
                    if (uni(status == 0)) status = check(N, 1);
                    if (uni(status == 0)) status = ACMPC_MAX_ITER_REACHED;
                }
                AC_PHASE(c, 12);   // checks
                if (uni(status != 0)) break;
            } else {
                iterate<false>(alpha, sigma);
            }
        }
        info.status = status, info.iter = iter, info.rho_updates = updates;
        info.pri_res = N.pri, info.dua_res = N.dua;
        VD obj = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            obj = obj + (VD(0.5) * p[j] * x[j] * x[j] + q[j] * x[j]);
            vout[j] = x[j] / di[j];
        }
        info.obj_val = final_obj(status, wsum(obj) * cinv);
        if (warm) {
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                st_lane(wrow + (0 * C + j) * 32, x[j]), st_lane(wrow + (1 * C + j) * 32, za[j]);
                st_lane(wrow + (2 * C + j) * 32, zb[j]), st_lane(wrow + (3 * C + j) * 32, ya[j]);
                st_lane(wrow + (4 * C + j) * 32, yb[j]);
            }
            AC_LANE0
            {
                warm[WR_RHO + slot] = R.rho;
                warm[WR_VALID + slot] = 1.0;
            }
        }
    }
};

// ================================================================================================
// Control QP (solvers/control.py + dynamics.py:65-103).  Stage s owns x_s = (e_y, e_psi, t) as local
// variables 0..2 and u_{s-1} = (v, kappa_cmd) as local variables 3,4 (stage 0 has no input: its slots
// are all-zero dummies, as are the slots of stages >= H).
// ================================================================================================
template <int C>
struct ControlQP {
    const Ctx<C>& c;
    int n, H;
    VD x[C][5], zb[C][5], yb[C][5], ye[C][3], ze[C][3];
    VD dx[C][5], dyb[C][5], dye[C][3];
    VI cls[C];   // class of bound row j in bits 2j, 2j+1
    double cs, cinv, nq_unscaled, nq_scaled;
    RhoSet R;
    bool bad_bounds;   // some row has l > u: OSQP refuses the data (setup / update fail), nothing is solved

    AC_MEM explicit ControlQP(const Ctx<C>& ctx) : c(ctx), n(ctx.n), H(ctx.H), bad_bounds(false) {}

    // linearise + stack (dynamics.py:65-103, solvers/control.py:26-79), OSQP scale_data (Ruiz) and
    // set_rho_vec classes; the scaled problem goes to tensor memory (chunks G, H) and shared memory (cold).
    AC_MEM void setup(const PathRegs<C>& path, const VD (&vel)[C], double offset)
    {
        const acmpc_config& g = *c.cfg;
        const double eps = 1e-12;
        // t2s of the state (offset, 0, pi/2) on waypoint 0 (spatial_mpc.py:186-189, dynamics.py:23-40)
        const double psi0 = lane_value(path.psi[0], 0), xs0 = lane_value(path.xs[0], 0),
                     ys0 = lane_value(path.ys[0], 0);
        double x0[3];
        x0[0] = cos(psi0) * (0.0 - ys0) - sin(psi0) * (offset - xs0);
        x0[1] = np_mod_u((kPi / 2.0 - psi0) + kPi, 2.0 * kPi) - kPi;
        x0[2] = 0.0;
        const double margin = g.width / 2.0;
        const double kmax = tan(g.delta_max) / g.wheelbase;

        VD m[C][3], a[C][6], b[C][2], s[C][5], P[C][5], q[C][2], D[C][5], EE[C][3], EB[C][5];
        VD dp[C], kp[C], vp[C], wp[C];
        pull_prev<C>(path.dist, dp);
        pull_prev<C>(path.kap, kp);
        pull_prev<C>(vel, vp);
        pull_prev<C>(path.wid, wp);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB isx = vi_lt(st, H), hasA = vi_lt(st, n), hasU = vi_ge(st, 1) & isx;
            const VD d = path.dist[j], ka = path.kap[j], v = vel[j];
            const VD one_x = vsel(isx, VD(1.0), VD(0.0)), one_a = vsel(hasA, VD(1.0), VD(0.0)),
                     one_u = vsel(hasU, VD(1.0), VD(0.0));
            for (int r = 0; r < 3; ++r) m[j][r] = -one_x;
            a[j][0] = one_a;
            a[j][1] = vsel(hasA, d, VD(0.0));
            a[j][2] = vsel(hasA, -(ka * ka) * d, VD(0.0));
            a[j][3] = one_a;
            a[j][4] = vsel(hasA, -ka / (v * d + VD(eps)), VD(0.0));
            a[j][5] = one_a;
            b[j][0] = vsel(hasU, dp[j], VD(0.0));
            b[j][1] = vsel(hasU, VD(-1.0) / (vp[j] * vp[j] * dp[j] + VD(eps)), VD(0.0));
            for (int e = 0; e < 3; ++e) s[j][e] = one_x;
            s[j][3] = one_u, s[j][4] = one_u;
            for (int e = 0; e < 3; ++e) P[j][e] = vsel(hasA, VD(g.step_cost[e]), vsel(isx, VD(g.final_cost[e]), VD(0.0)));
            P[j][3] = vsel(hasU, VD(g.r_term[0]), VD(0.0));
            P[j][4] = vsel(hasU, VD(g.r_term[1]), VD(0.0));
            q[j][0] = vsel(hasU, VD(-g.r_term[0]) * vp[j], VD(0.0));
            q[j][1] = vsel(hasU, VD(-g.r_term[1]) * kp[j], VD(0.0));
            for (int e = 0; e < 5; ++e) D[j][e] = VD(1.0), EB[j][e] = VD(1.0);
            for (int r = 0; r < 3; ++r) EE[j][r] = VD(1.0);
        }
        // ---- Ruiz equilibration
        cs = 1.0;
        const double inv_nv = 1.0 / (double)(5 * H - 2);   // mean over the 5H - 2 columns
        for (int pass = 0; pass < g.scaling; ++pass) {
            VD pr[C][3], prp[C][3], ee[C][3], een[C][3], dd[C][5];
            // partial row norms of block s+1 from the columns of stage s
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                pr[j][0] = vmax(vabs(a[j][0]), vabs(a[j][1]));
                pr[j][1] = vmax(vabs(a[j][2]), vabs(a[j][3]));
                pr[j][2] = vmax(vabs(a[j][4]), vabs(a[j][5]));
            }
            pull_prev_k<C, 3>(pr, prp);
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                VD am[3], sm[5], pm[5];
                for (int r = 0; r < 3; ++r) am[r] = vabs(m[j][r]);
                for (int e = 0; e < 5; ++e) sm[e] = vabs(s[j][e]), pm[e] = vabs(P[j][e]);
                VD a11 = vabs(a[j][0]), a12 = vabs(a[j][1]), a21 = vabs(a[j][2]), a22 = vabs(a[j][3]),
                   a31 = vabs(a[j][4]), a33 = vabs(a[j][5]), b22 = vabs(b[j][0]), b31 = vabs(b[j][1]);
                VD cn[5];
                cn[0] = vmax(vmax(vmax(am[0], a11), vmax(a21, a31)), vmax(sm[0], pm[0]));
                cn[1] = vmax(vmax(am[1], a12), vmax(a22, vmax(sm[1], pm[1])));
                cn[2] = vmax(vmax(am[2], a33), vmax(sm[2], pm[2]));
                cn[3] = vmax(b31, vmax(sm[3], pm[3]));
                cn[4] = vmax(b22, vmax(sm[4], pm[4]));
                for (int e = 0; e < 5; ++e) dd[j][e] = inv_sqrt(limit_scaling(cn[e]));
                ee[j][0] = inv_sqrt(limit_scaling(vmax(am[0], prp[j][0])));
                ee[j][1] = inv_sqrt(limit_scaling(vmax(vmax(am[1], b22), prp[j][1])));
                ee[j][2] = inv_sqrt(limit_scaling(vmax(vmax(am[2], b31), prp[j][2])));
            }
            pull_next_k<C, 3>(ee, een);
            VD psum = VD(0.0), qmax = VD(0.0);
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                for (int r = 0; r < 3; ++r) {
                    m[j][r] = (m[j][r] * ee[j][r]) * dd[j][r];
                    EE[j][r] = EE[j][r] * ee[j][r];
                }
                a[j][0] = (a[j][0] * een[j][0]) * dd[j][0];
                a[j][1] = (a[j][1] * een[j][0]) * dd[j][1];
                a[j][2] = (a[j][2] * een[j][1]) * dd[j][0];
                a[j][3] = (a[j][3] * een[j][1]) * dd[j][1];
                a[j][4] = (a[j][4] * een[j][2]) * dd[j][0];
                a[j][5] = (a[j][5] * een[j][2]) * dd[j][2];
                b[j][0] = (b[j][0] * ee[j][1]) * dd[j][4];
                b[j][1] = (b[j][1] * ee[j][2]) * dd[j][3];
                for (int e = 0; e < 5; ++e) {
                    VD eb = inv_sqrt(limit_scaling(vabs(s[j][e])));
                    s[j][e] = (s[j][e] * eb) * dd[j][e];
                    EB[j][e] = EB[j][e] * eb;
                    D[j][e] = D[j][e] * dd[j][e];
                    P[j][e] = (P[j][e] * dd[j][e]) * dd[j][e];
                }
                // column-norm sum of P as a tree: the passes are a dependent chain, a 5C-deep running sum sat on it
                psum = psum + ((vabs(P[j][0]) + vabs(P[j][1])) + (vabs(P[j][2]) + vabs(P[j][3])) + vabs(P[j][4]));
                for (int e = 0; e < 2; ++e) {
                    q[j][e] = q[j][e] * dd[j][3 + e];
                    qmax = vmax(qmax, vabs(q[j][e]));
                }
            }
            double ct = fmax(wsum(psum) * inv_nv, limit_scaling_u(wmax(qmax)));
            ct = urecip(limit_scaling_u(ct));
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                for (int e = 0; e < 5; ++e) P[j][e] = P[j][e] * VD(ct);
                for (int e = 0; e < 2; ++e) q[j][e] = q[j][e] * VD(ct);
            }
            cs *= ct;
        }
        cinv = 1.0 / cs;
        // ---- bounds (solvers/control.py:47-70,121-149, clipped to +-OSQP_INFTY), classes, stores
        VD nqu = VD(0.0), nqs = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB isx = vi_lt(st, H), hasU = vi_ge(st, 1) & isx, first = vi_eq(st, 0);
            VD lo[5], hi[5], be[3], G[16], Hc[16];
            lo[0] = vsel(first, VD(x0[0]), (-wp[j] / VD(2.0)) + VD(margin));
            hi[0] = vsel(first, VD(x0[0]), (wp[j] / VD(2.0)) - VD(margin));
            lo[1] = VD(-kInfty), hi[1] = VD(kInfty);
            lo[2] = VD(0.01), hi[2] = VD(kInfty);
            lo[3] = VD(g.input_v_min - 0.1), hi[3] = VD(g.input_v_max + 0.1);
            lo[4] = VD(-kmax), hi[4] = VD(kmax);
            const VD b31raw = VD(-1.0) / (vp[j] * vp[j] * dp[j] + VD(eps)), f3 = VD(1.0) / (vp[j] * dp[j] + VD(eps));
            be[0] = vsel(first, VD(-x0[0]), VD(0.0) * vp[j] + VD(0.0) * kp[j] - VD(0.0));
            be[1] = vsel(first, VD(-x0[1]), VD(0.0) * vp[j] + dp[j] * kp[j] - VD(0.0));
            be[2] = vsel(first, VD(-x0[2]), b31raw * vp[j] + VD(0.0) * kp[j] - f3);
            VI bits = vi_all(0);
            for (int e = 0; e < 5; ++e) {
                VB real = (e < 3) ? isx : hasU;
                if (wany(real & (lo[e] > hi[e]))) bad_bounds = true;
                VD l = vsel(real, lo[e] * EB[j][e], VD(0.0)), u = vsel(real, hi[e] * EB[j][e], VD(0.0));
                bits = bits | (row_class(l, u) << (2 * e));
                Hc[HC_LB + e] = l, Hc[HC_UB + e] = u;
                G[GC_S + e] = s[j][e];
                VD dinv = VD(1.0) / D[j][e];
                c.st(K_P + e, j, P[j][e]), c.st(K_DI + e, j, dinv), c.st(K_EBI + e, j, VD(1.0) / EB[j][e]);
                if (e >= 3) {
                    nqu = vmax(nqu, vabs(dinv * q[j][e - 3]));
                    nqs = vmax(nqs, vabs(q[j][e - 3]));
                }
            }
            cls[j] = bits;
            for (int r = 0; r < 3; ++r) {
                Hc[HC_BE + r] = vsel(isx, be[r] * EE[j][r], VD(0.0));
                c.st(K_EEI + r, j, VD(1.0) / EE[j][r]);
                G[GC_M + r] = m[j][r];
            }
            for (int e = 0; e < 6; ++e) G[GC_A + e] = a[j][e];
            G[GC_B22] = b[j][0], G[GC_B31] = b[j][1];
            Hc[HC_Q + 0] = q[j][0], Hc[HC_Q + 1] = q[j][1], Hc[15] = VD(0.0);
            c.tst(T_G, j, G), c.tst(T_H, j, Hc);
        }
        nq_unscaled = wmax(nqu);
        nq_scaled = wmax(nqs);
        warp_sync();
    }

    // Reduced matrix, analytic elimination of the inputs, block LDL' over the states, scan matrices.
    AC_MEM void factor()
    {
        const double sigma = c.cfg->sigma, re = R.rho_eq;
        VD g2[C][2], g2n[C][2], av[C][6], ap[C][6], mv[C][3], dg[C][3];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD G[16], F[16];
            c.tld(T_G, j, G);
            const VD b22 = G[GC_B22], b31 = G[GC_B31], s3 = G[GC_S + 3], s4 = G[GC_S + 4];
            for (int e = 0; e < 5; ++e) F[FC_RHO + e] = R.of(cls[j], 2 * e), F[FC_RINV + e] = R.inv_of(cls[j], 2 * e);
            VD kv = c.ld(K_P + 3, j) + VD(sigma) + F[FC_RHO + 3] * s3 * s3 + VD(re) * b31 * b31;
            VD kk = c.ld(K_P + 4, j) + VD(sigma) + F[FC_RHO + 4] * s4 * s4 + VD(re) * b22 * b22;
            VD iv = VD(1.0) / kv, ik = VD(1.0) / kk;
            VD rb31 = VD(re) * b31, rb22 = VD(re) * b22;
            F[FC_IV] = iv, F[FC_IK] = ik, F[FC_RB31] = rb31, F[FC_RB22] = rb22, F[14] = VD(0.0), F[15] = VD(0.0);
            c.tst(T_F, j, F);
            g2[j][0] = iv * rb31 * rb31;   // gv
            g2[j][1] = ik * rb22 * rb22;   // gk
            for (int e = 0; e < 6; ++e) av[j][e] = G[GC_A + e];
            for (int r = 0; r < 3; ++r) {
                mv[j][r] = G[GC_M + r];
                const VD sr = G[GC_S + r];
                dg[j][r] = c.ld(K_P + r, j) + VD(sigma) + F[FC_RHO + r] * sr * sr + VD(re) * mv[j][r] * mv[j][r];
            }
        }
        pull_next_k<C, 2>(g2, g2n);
        pull_prev_k<C, 6>(av, ap);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            const VD m0 = mv[j][0], m1 = mv[j][1], m2 = mv[j][2];
            const VD gv = g2[j][0], gk = g2[j][1], gvn = g2n[j][0], gkn = g2n[j][1];
            const VD a11 = av[j][0], a12 = av[j][1], a21 = av[j][2], a22 = av[j][3], a31 = av[j][4], a33 = av[j][5];
            VD k00 = dg[j][0], k11 = dg[j][1], k22 = dg[j][2];
            // elimination of the own input u_{s-1}
            k11 = k11 - gk * m1 * m1;
            k22 = k22 - gv * m2 * m2;
            // A_s' rho_eq A_s and the elimination of u_s (owned by stage s+1)
            k00 = k00 + VD(re) * (a11 * a11 + a21 * a21 + a31 * a31) - (gvn * a31 * a31 + gkn * a21 * a21);
            VD k10 = VD(re) * (a11 * a12 + a21 * a22) - gkn * a21 * a22;
            k11 = k11 + VD(re) * (a12 * a12 + a22 * a22) - gkn * a22 * a22;
            VD k20 = VD(re) * (a31 * a33) - gvn * a31 * a33;
            k22 = k22 + VD(re) * (a33 * a33) - gvn * a33 * a33;
            st_lane(c.wcol(W_SS + 0, j), k00), st_lane(c.wcol(W_SS + 1, j), k10), st_lane(c.wcol(W_SS + 2, j), k11);
            st_lane(c.wcol(W_SS + 3, j), k20), st_lane(c.wcol(W_SS + 4, j), VD(0.0)), st_lane(c.wcol(W_SS + 5, j), k22);
            // S_{s,s-1}: rows of x_s, columns of x_{s-1} (entries of A_{s-1})
            const VD p11 = ap[j][0], p12 = ap[j][1], p21 = ap[j][2], p22 = ap[j][3], p31 = ap[j][4], p33 = ap[j][5];
            st_lane(c.wcol(W_OFF + 0, j), m0 * VD(re) * p11);
            st_lane(c.wcol(W_OFF + 1, j), m0 * VD(re) * p12);
            st_lane(c.wcol(W_OFF + 2, j), VD(0.0));
            st_lane(c.wcol(W_OFF + 3, j), m1 * (VD(re) * p21 - gk * p21));
            st_lane(c.wcol(W_OFF + 4, j), m1 * (VD(re) * p22 - gk * p22));
            st_lane(c.wcol(W_OFF + 5, j), VD(0.0));
            st_lane(c.wcol(W_OFF + 6, j), m2 * (VD(re) * p31 - gv * p31));
            st_lane(c.wcol(W_OFF + 7, j), VD(0.0));
            st_lane(c.wcol(W_OFF + 8, j), m2 * (VD(re) * p33 - gv * p33));
        }
        warp_sync();
        // serial block LDL': Sigma_s = S_ss - G_s S_{s,s-1}', G_s = S_{s,s-1} Sigma_{s-1}^{-1}; N_s = -G_s.
        // Every lane runs the recurrence on broadcast reads; lane 0 publishes N_s and Sigma_s^{-1} in place.
        {
            double i00 = 0, i10 = 0, i11 = 0, i20 = 0, i21 = 0, i22 = 0;
            int s = 0;
            for (int l = 0; l < 32 && s < H; ++l)
                for (int j = 0; j < C && s < H; ++j, ++s) {
                    double* SI = c.wcol(W_SS, j) + l;
                    double* NN = c.wcol(W_OFF, j) + l;
                    const int st = C * 32;   // distance between consecutive fields
                    double a00 = SI[0 * st], a10 = SI[1 * st], a11 = SI[2 * st];
                    double a20 = SI[3 * st], a21 = SI[4 * st], a22 = SI[5 * st];
                    double gg[9];
                    {
                        // structural zeros of S_{s,s-1}: (0,2) (1,2) (2,1)
                        const double o00 = NN[0 * st], o01 = NN[1 * st], o10 = NN[3 * st], o11 = NN[4 * st];
                        const double o20 = NN[6 * st], o22 = NN[8 * st];
                        gg[0] = o00 * i00 + o01 * i10, gg[1] = o00 * i10 + o01 * i11, gg[2] = o00 * i20 + o01 * i21;
                        gg[3] = o10 * i00 + o11 * i10, gg[4] = o10 * i10 + o11 * i11, gg[5] = o10 * i20 + o11 * i21;
                        gg[6] = o20 * i00 + o22 * i20, gg[7] = o20 * i10 + o22 * i21, gg[8] = o20 * i20 + o22 * i22;
                        a00 -= gg[0] * o00 + gg[1] * o01;
                        a10 -= gg[3] * o00 + gg[4] * o01;
                        a11 -= gg[3] * o10 + gg[4] * o11;
                        a20 -= gg[6] * o00 + gg[7] * o01;
                        a21 -= gg[6] * o10 + gg[7] * o11;
                        a22 -= gg[6] * o20 + gg[8] * o22;
                    }
                    // Inverse of the SPD 3x3 by cofactors and ONE reciprocal of the determinant.  The recurrence is a pure
                    // latency chain (every lane runs it, one stage after the other): an LDL'-based inverse puts three
                    // dependent reciprocals (~6 dependent FP64 ops each) on it, the cofactor form one -- measured 33 k ->
                    // 19 k cycles per factorisation at H = 50.  Same conditioning as the pivots of LDL' (the blocks are
                    // dominated by the rho_eq terms of the dynamics rows; tests hold the solve to the oracle at 1e-7).
                    const double c00 = fma(a11, a22, -(a21 * a21)), c10 = fma(a21, a20, -(a10 * a22));
                    const double c20 = fma(a10, a21, -(a11 * a20)), c11 = fma(a00, a22, -(a20 * a20));
                    const double c21 = fma(a10, a20, -(a00 * a21)), c22 = fma(a00, a11, -(a10 * a10));
                    const double rdet = urecip(fma(a00, c00, fma(a10, c10, a20 * c20)));
                    i00 = c00 * rdet, i10 = c10 * rdet, i11 = c11 * rdet;
                    i20 = c20 * rdet, i21 = c21 * rdet, i22 = c22 * rdet;
                    AC_LANE0
                    {
                        SI[0 * st] = i00, SI[1 * st] = i10, SI[2 * st] = i11;
                        SI[3 * st] = i20, SI[4 * st] = i21, SI[5 * st] = i22;
                        for (int e = 0; e < 9; ++e) NN[e * st] = -gg[e];
                    }
                }
        }
        warp_sync();
        // own stages' factor -> tensor memory; lane-level prefix products for the scans -> scratch region
        // (level L maps the carry of lane l - 2^L to lane l)
        VD F[9];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD NS[16];
            for (int e = 0; e < 9; ++e) NS[NC_N + e] = ld_lane(c.wcol(W_OFF + e, j));
            for (int e = 0; e < 6; ++e) NS[NC_SI + e] = ld_lane(c.wcol(W_SS + e, j));
            NS[15] = VD(0.0);
            c.tst(T_NS, j, NS);
            if (j == 0) {
                for (int e = 0; e < 9; ++e) F[e] = NS[NC_N + e];
            } else {
                VD T[9];
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q)
                        T[3 * r + q] = NS[3 * r] * F[q] + NS[3 * r + 1] * F[3 + q] + NS[3 * r + 2] * F[6 + q];
                for (int e = 0; e < 9; ++e) F[e] = T[e];
            }
        }
        warp_sync();   // every lane has read its work columns: the region now takes the scan matrices
        for (int L = 0; L < kLevels; ++L) {
            VD U[9], T[9];
            for (int e = 0; e < 9; ++e) {
                st_lane(c.scan(L, e), F[e]);
                // re-zero the 16 lanes past lane 31 (the factorisation's work arrays lived here)
                if (Layout<C>::kScanLanes > 32) st_idx_if(vi_lt(c.lane, 16), c.scan(L, e), c.lane + 32, VD(0.0));
                U[e] = shfl_up0(F[e], 1 << L);
            }
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q)
                    T[3 * r + q] = F[3 * r] * U[q] + F[3 * r + 1] * U[3 + q] + F[3 * r + 2] * U[6 + q];
            for (int e = 0; e < 9; ++e) F[e] = T[e];
        }
        warp_sync();
    }

    // x~ = Schur^{-1} r for the state part: two affine scans over the stages (see file header)
    AC_MEM void block_solve(const VD (&r)[C][3], VD (&xt)[C][3])
    {
        VD y[C][3], w[C][3], M[9], NS[C][16];
        if (C == 2 && !Layout<C>::in_smem(T_NS, 0)) {   // both stages' factors under one tensor-memory wait
            tm_ld_group2<16, 16>(c.tm, Layout<C>::tmem_off(T_NS, 0), NS[0], Layout<C>::tmem_off(T_NS, C - 1), NS[C - 1]);
        } else {
            AC_UNROLL
            for (int j = 0; j < C; ++j) c.tld(T_NS, j, NS[j]);
        }
        // forward: local pass with zero carry-in, scan over lanes, local fix-up
        VD Y[3] = {r[0][0], r[0][1], r[0][2]};
        AC_UNROLL
        for (int j = 1; j < C; ++j) {
            VD T[3] = {r[j][0], r[j][1], r[j][2]};
            mv_acc9(&NS[j][NC_N], Y, T);
            Y[0] = T[0], Y[1] = T[1], Y[2] = T[2];
        }
        // (the level matrices are zero on lanes < 2^L and N_0 is zero on lane 0: raw shuffles suffice)
        {
            const double* sp = c.scan(0, 0);
            AC_NOUNROLL
            for (int L = 0; L < kLevels; ++L, sp += 9 * Layout<C>::kScanLanes) {
                VD U[3];
                for (int e = 0; e < 3; ++e) U[e] = shfl_up_raw(Y[e], 1 << L);
                for (int e = 0; e < 9; ++e) M[e] = ld_lane(sp + e * Layout<C>::kScanLanes);
                mv_acc9(M, U, Y);
            }
        }
        {
            VD U[3];
            for (int e = 0; e < 3; ++e) U[e] = shfl_up_raw(Y[e], 1);
            for (int e = 0; e < 3; ++e) y[0][e] = r[0][e];
            mv_acc9(&NS[0][NC_N], U, y[0]);
        }
        AC_UNROLL
        for (int j = 1; j + 1 < C; ++j) {
            for (int e = 0; e < 3; ++e) y[j][e] = r[j][e];
            mv_acc9(&NS[j][NC_N], y[j - 1], y[j]);
        }
        if (C > 1)
            for (int e = 0; e < 3; ++e) y[C - 1][e] = Y[e];
        // diagonal
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            const VD i00 = NS[j][NC_SI + 0], i10 = NS[j][NC_SI + 1], i11 = NS[j][NC_SI + 2];
            const VD i20 = NS[j][NC_SI + 3], i21 = NS[j][NC_SI + 4], i22 = NS[j][NC_SI + 5];
            w[j][0] = i00 * y[j][0] + i10 * y[j][1] + i20 * y[j][2];
            w[j][1] = i10 * y[j][0] + i11 * y[j][1] + i21 * y[j][2];
            w[j][2] = i20 * y[j][0] + i21 * y[j][1] + i22 * y[j][2];
        }
        // backward: x_s = w_s + N_{s+1}' x_{s+1}; carry = x of the lane's LAST stage
        VD A[3] = {VD(0.0), VD(0.0), VD(0.0)};
        if (C > 1) {
            for (int e = 0; e < 3; ++e) A[e] = w[C - 2][e];
            AC_UNROLL
            for (int j = C - 3; j >= 0; --j) {
                VD T[3] = {w[j][0], w[j][1], w[j][2]};
                mtv_acc9(&NS[j + 1][NC_N], A, T);
                A[0] = T[0], A[1] = T[1], A[2] = T[2];
            }
        }
        VD Z[3];
        {
            VD Hh[3] = {VD(0.0), VD(0.0), VD(0.0)};
            if (C > 1) mtv_acc9(&NS[0][NC_N], A, Hh);
            for (int e = 0; e < 3; ++e) Z[e] = w[C - 1][e] + shfl_down0(Hh[e], 1);
        }
        {
            const double* sp = c.scan(0, 0);
            AC_NOUNROLL
            for (int L = 0; L < kLevels; ++L, sp += 9 * Layout<C>::kScanLanes) {
                VD U[3];
                if (Layout<C>::kScanLanes > 32) {
                    for (int e = 0; e < 3; ++e) U[e] = shfl_down_raw(Z[e], 1 << L);
                    const double* spl = sp + (1 << L);   // lane + 2^L: zeros past lane 31
                    for (int e = 0; e < 9; ++e) M[e] = ld_lane(spl + e * Layout<C>::kScanLanes);
                } else {
                    for (int e = 0; e < 3; ++e) U[e] = shfl_down0(Z[e], 1 << L);
                    for (int e = 0; e < 9; ++e) M[e] = ld_lane_at(sp + e * Layout<C>::kScanLanes, 1 << L);
                }
                mtv_acc9(M, U, Z);
            }
        }
        for (int e = 0; e < 3; ++e) xt[C - 1][e] = Z[e];
        AC_UNROLL
        for (int j = C - 2; j >= 0; --j) {
            for (int e = 0; e < 3; ++e) xt[j][e] = w[j][e];
            mtv_acc9(&NS[j + 1][NC_N], xt[j + 1], xt[j]);
        }
    }

    // One ADMM iteration (osqp_solve loop body): rhs, reduced solve, recover inputs, relax, project, duals.
    // FIRST: the first iteration of a solve reads the z of the equality rows from the start state (`ze`: zero on a
    // cold start, the previous problem's right-hand side on a warm start); from then on that z IS the scaled
    // right-hand side b (projection onto [b, b]), so the hot instantiation carries no ze registers.
    // KEEP: also record delta_x / delta_y for the infeasibility certificates.  Only the iterations that are
    // followed by a termination check are KEEP instantiations, and the check sits in the same block of
    // solve(), so the 13 delta values per stage never live across the loop.
    template <bool FIRST, bool KEEP>
    AC_MEM void iterate()
    {
        const double sigma = c.cfg->sigma, alpha = c.cfg->alpha, re = R.rho_eq;
        VD t[C][3], tn[C][3], ru[C][2], r[C][3], xt[C][3], Aj[C][6];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD F[9], G[16], Q[8];   // F: rho[5] iv ik rb31 rb22 (no rinv here); Q: q[2] be[3] lb[0..2]
            {
                VD f8[8], f1[1];
                c.template hld4<8, 1, 16, 8>(j, T_F, 0, f8, T_F, 8, f1, T_G, 0, G, T_H, HC_Q, Q);
                for (int k = 0; k < 8; ++k) F[k] = f8[k];
                F[8] = f1[0];
            }
            const VD z0 = FIRST ? ze[j][0] : Q[HC_BE + 0], z1 = FIRST ? ze[j][1] : Q[HC_BE + 1],
                     z2 = FIRST ? ze[j][2] : Q[HC_BE + 2];
            VD w0 = VD(re) * z0 - ye[j][0], w1 = VD(re) * z1 - ye[j][1], w2 = VD(re) * z2 - ye[j][2];
            VD wb3 = F[FC_RHO + 3] * zb[j][3] - yb[j][3], wb4 = F[FC_RHO + 4] * zb[j][4] - yb[j][4];
            ru[j][0] = VD(sigma) * x[j][3] + G[GC_S + 3] * wb3 + G[GC_B31] * w2 - Q[0];
            ru[j][1] = VD(sigma) * x[j][4] + G[GC_S + 4] * wb4 + G[GC_B22] * w1 - Q[1];
            VD pv = F[FC_IV] * ru[j][0], pk = F[FC_IK] * ru[j][1];
            t[j][0] = w0;
            t[j][1] = w1 - F[FC_RB22] * pk;
            t[j][2] = w2 - F[FC_RB31] * pv;
            // the part of the state rhs that does not need the next stage
            for (int q = 0; q < 3; ++q) {
                VD wb = F[FC_RHO + q] * zb[j][q] - yb[j][q];
                r[j][q] = VD(sigma) * x[j][q] + G[GC_S + q] * wb + G[GC_M + q] * t[j][q];
            }
            for (int e = 0; e < 6; ++e) Aj[j][e] = G[GC_A + e];
        }
        pull_next_raw<C, 3>(t, tn);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            r[j][0] = r[j][0] + (Aj[j][0] * tn[j][0] + Aj[j][2] * tn[j][1] + Aj[j][4] * tn[j][2]);
            r[j][1] = r[j][1] + (Aj[j][1] * tn[j][0] + Aj[j][3] * tn[j][1]);
            r[j][2] = r[j][2] + Aj[j][5] * tn[j][2];
        }
        block_solve(r, xt);
        // p^ = A_s x~_s goes up to stage s+1
        VD ph[C][3], cp[C][3];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            ph[j][0] = Aj[j][0] * xt[j][0] + Aj[j][1] * xt[j][1];
            ph[j][1] = Aj[j][2] * xt[j][0] + Aj[j][3] * xt[j][1];
            ph[j][2] = Aj[j][4] * xt[j][0] + Aj[j][5] * xt[j][2];
        }
        pull_prev_rot<C, 3>(ph, cp);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD F[16], G[10], Hc[16];   // of chunk G only s[5] m[3] b22 b31 are needed here (not A)
            {
                VD g8[8], g2[2];
                c.template hld4<16, 8, 2, 16>(j, T_F, 0, F, T_G, 0, g8, T_G, 8, g2, T_H, 0, Hc);
                for (int k = 0; k < 8; ++k) G[k] = g8[k];
                G[8] = g2[0], G[9] = g2[1];
            }
            VD e0 = G[GC_M + 0] * xt[j][0] + cp[j][0];
            VD e1 = G[GC_M + 1] * xt[j][1] + cp[j][1];
            VD e2 = G[GC_M + 2] * xt[j][2] + cp[j][2];
            VD utv = F[FC_IV] * (ru[j][0] - F[FC_RB31] * e2);
            VD utk = F[FC_IK] * (ru[j][1] - F[FC_RB22] * e1);
            VD zte[3] = {e0, e1 + G[GC_B22] * utk, e2 + G[GC_B31] * utv};
            VD xv[5] = {xt[j][0], xt[j][1], xt[j][2], utv, utk};
            // equality rows: l == u == b
            for (int q = 0; q < 3; ++q) {
                VD zn = Hc[HC_BE + q];   // projection onto [b, b]
                VD zh = VD(alpha) * zte[q] + VD(1.0 - alpha) * (FIRST ? ze[j][q] : zn);
                VD dy = VD(re) * (zh - zn);
                ye[j][q] = ye[j][q] + dy;
                if (KEEP) dye[j][q] = dy;
            }
            for (int e = 0; e < 5; ++e) {
                VD zt = G[GC_S + e] * xv[e];
                VD zh = VD(alpha) * zt + VD(1.0 - alpha) * zb[j][e];
                VD zn = vclamp(zh + F[FC_RINV + e] * yb[j][e], Hc[HC_LB + e], Hc[HC_UB + e]);
                VD dy = F[FC_RHO + e] * (zh - zn);
                zb[j][e] = zn;
                yb[j][e] = yb[j][e] + dy;
                if (KEEP) dyb[j][e] = dy;
                VD xn = VD(alpha) * xv[e] + VD(1.0 - alpha) * x[j][e];
                if (KEEP) dx[j][e] = xn - x[j][e];
                x[j][e] = xn;
            }
        }
    }

    // A v on the equality rows owned by each stage (v: 5 local variables per stage)
    AC_MEM void eq_rows(const VD (&v)[C][5], VD (&o)[C][3])
    {
        VD ph[C][3], cp[C][3], G[C][16];
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            c.tld(T_G, j, G[j]);
            ph[j][0] = G[j][GC_A + 0] * v[j][0] + G[j][GC_A + 1] * v[j][1];
            ph[j][1] = G[j][GC_A + 2] * v[j][0] + G[j][GC_A + 3] * v[j][1];
            ph[j][2] = G[j][GC_A + 4] * v[j][0] + G[j][GC_A + 5] * v[j][2];
        }
        pull_prev_k<C, 3>(ph, cp);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            o[j][0] = G[j][GC_M + 0] * v[j][0] + cp[j][0];
            o[j][1] = G[j][GC_M + 1] * v[j][1] + cp[j][1] + G[j][GC_B22] * v[j][4];
            o[j][2] = G[j][GC_M + 2] * v[j][2] + cp[j][2] + G[j][GC_B31] * v[j][3];
        }
    }
    // A' (ve, vb) on the columns owned by each stage
    AC_MEM void at_cols(const VD (&ve)[C][3], const VD (&vb)[C][5], VD (&o)[C][5])
    {
        VD vn[C][3];
        pull_next_k<C, 3>(ve, vn);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD G[16];
            c.tld(T_G, j, G);
            o[j][0] = G[GC_S + 0] * vb[j][0] + G[GC_M + 0] * ve[j][0] +
                      (G[GC_A + 0] * vn[j][0] + G[GC_A + 2] * vn[j][1] + G[GC_A + 4] * vn[j][2]);
            o[j][1] = G[GC_S + 1] * vb[j][1] + G[GC_M + 1] * ve[j][1] + (G[GC_A + 1] * vn[j][0] + G[GC_A + 3] * vn[j][1]);
            o[j][2] = G[GC_S + 2] * vb[j][2] + G[GC_M + 2] * ve[j][2] + G[GC_A + 5] * vn[j][2];
            o[j][3] = G[GC_S + 3] * vb[j][3] + G[GC_B31] * ve[j][2];
            o[j][4] = G[GC_S + 4] * vb[j][4] + G[GC_B22] * ve[j][1];
        }
    }

    AC_MEM void compute_norms(Norms& N)
    {
        VD v[12];
        for (int t = 0; t < 12; ++t) v[t] = VD(0.0);
        VD ax[C][3], aty[C][5];
        eq_rows(x, ax);
        at_cols(ye, yb, aty);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD S5[8], Q[2];
            c.template hld<8>(T_G, GC_S, j, S5);
            c.template hld<2>(T_H, HC_Q, j, Q);
            // after at least one iteration the z of the equality rows is their right-hand side
            for (int r = 0; r < 3; ++r)
                acc_row(v, ax[j][r], c.hld1(T_H, HC_BE + r, j), c.ld(K_EEI + r, j));
            for (int e = 0; e < 5; ++e) {
                acc_row(v, S5[e] * x[j][e], zb[j][e], c.ld(K_EBI + e, j));
                VD q = (e >= 3) ? Q[e - 3] : VD(0.0);
                acc_col(v, q, c.ld(K_P + e, j) * x[j][e], aty[j][e], c.ld(K_DI + e, j));
            }
        }
        finish_norms(v, cinv, nq_unscaled, nq_scaled, N);
        if (c.cfg->check_dualgap) {
            VD a = VD(0.0), b = VD(0.0), s = VD(0.0);
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                VD Hc[16];
                c.tld(T_H, j, Hc);
                for (int r = 0; r < 3; ++r) s = s + Hc[HC_BE + r] * ye[j][r];   // l == u == b
                for (int e = 0; e < 5; ++e) {
                    a = a + c.ld(K_P + e, j) * x[j][e] * x[j][e];
                    if (e >= 3) b = b + Hc[HC_Q + e - 3] * x[j][e];
                    s = s + support_finite(yb[j][e], Hc[HC_LB + e], Hc[HC_UB + e]);
                }
            }
            N.xtPx = wsum(a), N.qtx = wsum(b), N.sc = wsum(s);
        }
    }

    AC_MEM int primal_infeasible(double eps)
    {
        VD nrm = VD(0.0), lhs = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD Hc[16];
            c.tld(T_H, j, Hc);
            for (int r = 0; r < 3; ++r) {   // equality rows: finite bounds, no projection
                VD b = Hc[HC_BE + r];
                nrm = vmax(nrm, vabs(dye[j][r] / c.ld(K_EEI + r, j)));
                lhs = lhs + support(dye[j][r], b, b);
            }
            for (int e = 0; e < 5; ++e) {
                VD lo = Hc[HC_LB + e], hi = Hc[HC_UB + e];
                dyb[j][e] = project_dy(dyb[j][e], lo, hi);
                nrm = vmax(nrm, vabs(dyb[j][e] / c.ld(K_EBI + e, j)));
                lhs = lhs + support(dyb[j][e], lo, hi);
            }
        }
        double nr = wmax(nrm), lh = wsum(lhs);
        if (uni(!(nr > eps) || !(lh < -eps * nr))) return 0;
        VD aty[C][5], m = VD(0.0);
        at_cols(dye, dyb, aty);
        AC_UNROLL
        for (int j = 0; j < C; ++j)
            for (int e = 0; e < 5; ++e) m = vmax(m, vabs(c.ld(K_DI + e, j) * aty[j][e]));
        return wmax(m) < eps * nr;
    }

    AC_MEM int dual_infeasible(double eps)
    {
        VD nrm = VD(0.0), qdx = VD(0.0), pm = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD Q[2];
            c.template hld<2>(T_H, HC_Q, j, Q);
            for (int e = 0; e < 5; ++e) {
                VD di = c.ld(K_DI + e, j);
                nrm = vmax(nrm, vabs(dx[j][e] / di));
                if (e >= 3) qdx = qdx + Q[e - 3] * dx[j][e];
                pm = vmax(pm, vabs(di * (c.ld(K_P + e, j) * dx[j][e])));
            }
        }
        double nr = wmax(nrm), qd = wsum(qdx), pmx = wmax(pm);
        if (uni(!(nr > eps) || !(qd < -cs * eps * nr) || !(pmx < cs * eps * nr))) return 0;
        VD adx[C][3];
        eq_rows(dx, adx);
        VB bad = vb_all(false);
        const VD lim = VD(eps * nr);
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VD Hc[16], S5[8];
            c.tld(T_H, j, Hc);
            c.template hld<8>(T_G, GC_S, j, S5);
            for (int r = 0; r < 3; ++r) {   // equality rows: both bounds finite
                VD a = adx[j][r] * c.ld(K_EEI + r, j);
                bad = bad | (a > lim) | (a < -lim);
            }
            for (int e = 0; e < 5; ++e) {
                VD a = c.ld(K_EBI + e, j) * (S5[e] * dx[j][e]);
                VD lo = Hc[HC_LB + e], hi = Hc[HC_UB + e];
                bad = bad | ((hi < VD(kBig)) & (a > lim)) | ((lo > VD(-kBig)) & (a < -lim));
            }
        }
        return !wany(bad);
    }

    AC_MEM int check(const Norms& N, int approximate)
    {
        const acmpc_config& g = *c.cfg;
        double k = approximate ? 10.0 : 1.0;
        if (uni(N.pri > kInfty || N.dua > kInfty)) return ACMPC_NON_CVX;
        double eps_p = k * g.eps_abs + k * g.eps_rel * fmax(N.nz, N.nAx);
        double eps_d = k * g.eps_abs + k * g.eps_rel * cinv * fmax(N.nq, fmax(N.nAty, N.nPx));
        int p_ok = uni(N.pri < eps_p), d_ok = uni(N.dua < eps_d);
        int p_inf = 0, d_inf = 0;
        if (!p_ok) p_inf = primal_infeasible(k * g.eps_prim_inf);
        if (!d_ok) d_inf = dual_infeasible(k * g.eps_dual_inf);
        const int g_ok = g.check_dualgap ? uni(dualgap_ok(N, cinv, k * g.eps_abs, k * g.eps_rel)) : 1;
        if (p_ok && d_ok && g_ok) return approximate ? ACMPC_SOLVED_INACCURATE : ACMPC_SOLVED;
        if (p_inf) return approximate ? ACMPC_PRIMAL_INFEASIBLE_INACCURATE : ACMPC_PRIMAL_INFEASIBLE;
        if (d_inf) return approximate ? ACMPC_DUAL_INFEASIBLE_INACCURATE : ACMPC_DUAL_INFEASIBLE;
        return 0;
    }

    AC_MEM void solve(SolveInfo& info, double* warm, bool use_warm)
    {
        const acmpc_config& g = *c.cfg;
        if (uni(bad_bounds)) {   // osqp_setup / osqp_update_bounds reject l > u: no solve, solver state untouched
            info.status = ACMPC_UNSOLVED, info.iter = 0, info.rho_updates = 0;
            info.pri_res = info.dua_res = info.obj_val = 0.0;
            AC_UNROLL
            for (int j = 0; j < C; ++j)
                for (int e = 0; e < 5; ++e) x[j][e] = VD(0.0);
            return;
        }
        double* wrow = warm ? warm + WR_HEADER + 2 * WR_SPEED_FIELDS * C * 32 : nullptr;
        const bool have = warm && use_warm && warm[WR_VALID + 2] != 0.0;
        R.set(have ? warm[WR_RHO + 2] : clampu(g.rho, kRhoMin, kRhoMax));
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            for (int e = 0; e < 5; ++e) x[j][e] = zb[j][e] = yb[j][e] = dx[j][e] = dyb[j][e] = VD(0.0);
            for (int r = 0; r < 3; ++r) ye[j][r] = ze[j][r] = dye[j][r] = VD(0.0);
            if (have) {
                for (int e = 0; e < 5; ++e) {
                    x[j][e] = ld_lane(wrow + ((0 + e) * C + j) * 32);
                    zb[j][e] = ld_lane(wrow + ((5 + e) * C + j) * 32);
                    yb[j][e] = ld_lane(wrow + ((10 + e) * C + j) * 32);
                }
                for (int r = 0; r < 3; ++r) {
                    ye[j][r] = ld_lane(wrow + ((15 + r) * C + j) * 32);
                    ze[j][r] = ld_lane(wrow + ((18 + r) * C + j) * 32);
                }
            }
        }
        warp_sync();
        factor();
        AC_PHASE(c, 2);   // first factorisation
        Norms N;
        int status = 0, iter = 0, updates = 0;
        // same structure as SpeedQP::solve
        bool checked = false, adapt = false, last = false;
        int to_check = g.check_termination > 0 ? g.check_termination : -1;
        int to_adapt = (g.adaptive_rho && g.adaptive_rho_interval > 0) ? g.adaptive_rho_interval : -1;
        for (;;) {
            ++iter;
            checked = (--to_check == 0);
            if (checked) to_check = g.check_termination;
            adapt = (--to_adapt == 0);
            if (adapt) to_adapt = g.adaptive_rho_interval;
            last = iter >= g.max_iter;
            const bool slow = checked || adapt || last;
            if (iter == 1 || slow) {
                if (iter == 1) iterate<true, true>();
                else iterate<false, true>();
                AC_PHASE(c, 3);   // iterations
                if (slow) {
                    compute_norms(N);
                    if (checked) status = check(N, 0);
                    if (uni(status == 0) && adapt) {
                        double rn = rho_estimate(N, R.rho);
                        if (uni(rn > R.rho * g.adaptive_rho_tolerance || rn < R.rho / g.adaptive_rho_tolerance)) {
                            R.set(rn);
                            ++updates;
                            factor();
                        }
                    }
                    if (uni(status == 0) && last) {
                        if (!checked) status = check(N, 0);
                        if (uni(status == 0)) status = check(N, 1);
                        if (uni(status == 0)) status = ACMPC_MAX_ITER_REACHED;
                    }
                    AC_PHASE(c, 4);   // termination checks (+ refactorisations)
                    if (uni(status != 0)) break;
                }
            } else {
                iterate<false, false>();
            }
        }
        info.status = status, info.iter = iter, info.rho_updates = updates;
        info.pri_res = N.pri, info.dua_res = N.dua;
        VD obj = VD(0.0);
        AC_UNROLL
        for (int j = 0; j < C; ++j)
            for (int e = 0; e < 5; ++e) {
                VD xe = x[j][e];
                obj = obj + VD(0.5) * c.ld(K_P + e, j) * xe * xe;
                if (e >= 3) obj = obj + c.hld1(T_H, HC_Q + e - 3, j) * xe;
            }
        info.obj_val = final_obj(status, wsum(obj) * cinv);
        if (warm) {
            AC_UNROLL
            for (int j = 0; j < C; ++j) {
                for (int e = 0; e < 5; ++e) {
                    st_lane(wrow + ((0 + e) * C + j) * 32, x[j][e]);
                    st_lane(wrow + ((5 + e) * C + j) * 32, zb[j][e]);
                    st_lane(wrow + ((10 + e) * C + j) * 32, yb[j][e]);
                }
                for (int r = 0; r < 3; ++r) {
                    st_lane(wrow + ((15 + r) * C + j) * 32, ye[j][r]);
                    st_lane(wrow + ((18 + r) * C + j) * 32, c.hld1(T_H, HC_BE + r, j));
                }
            }
            AC_LANE0
            {
                warm[WR_RHO + 2] = R.rho;
                warm[WR_VALID + 2] = 1.0;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// one full MPC step for one instance.  `raw_path` = (H,3) staged in shared memory (inside the scan
// region: dead until the first control factorisation).
// ------------------------------------------------------------------------------------------------
struct InstanceOut {
    double *controls, *prediction, *cum_time, *states, *v_ref, *cost, *pri_res, *dua_res;
    int32_t *status, *status_speed, *iters, *rho_updates;
    double* waypoints;
    double* derived;
};

// Phase 1 of the step: waypoints + speed-profile QP (spatial_mpc.py:176-184).  Uses registers and the scratch
// region only.  `vel_out` [n] always receives the profile (zeros unless the QP status is "solved",
// spatial_mpc.py:115-122); it is the hand-over to phase 2.
// `way` != nullptr: the stand-alone SpatialMPC.compute_speed_profile (spatial_mpc.py:89-123) on a ReferencePath the
// caller built -- kappas / distances are read from rows 3 / 4 of way[7,n] instead of a raw path, the velocities row
// is written only when the QP is "solved" (left untouched otherwise) and vel_out (may be NULL) receives dec.x as is.
template <int C>
AC_DEV int speed_instance(const Ctx<C>& c, const double* raw_path, double v_max_live, int localised,
                          double* vel_out, const InstanceOut& o, double* warm = nullptr, bool use_warm = false,
                          double* way = nullptr)
{
    const int n = c.n;
    PathRegs<C> path;
    if (way) {
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB ok = vi_lt(st, n);
            path.xs[j] = path.ys[j] = path.psi[j] = path.wid[j] = VD(0.0);
            path.kap[j] = ld_idx_if(ok, way + 3 * n, st, 0.0);
            path.dist[j] = ld_idx_if(ok, way + 4 * n, st, 0.0);
        }
    } else {
        build_waypoints<C>(c, raw_path, path);
    }
    warp_sync();
    SolveInfo si;
    VD vel[C];
    SpeedQP<C> sq(c);
    AC_PHASE(c, 8);    // speed kernel: staging + waypoints
    sq.assemble_and_scale(path, v_max_live, localised);
    AC_PHASE(c, 9);    // assembly + Ruiz
    sq.solve(si, vel, warm, localised ? 1 : 0, use_warm);
    AC_PHASE(c, 13);   // solve tail
    // Outputs leave as coalesced rows: lane l owns stages C*l .., a stride-C pattern that would write half-empty
    // sectors (and, when the outputs are peer-mapped, one small NVLink write each); every row goes through a
    // shared-memory tile and is written out lane-contiguous.  The scratch region is dead after the solve.
    double* T = c.W;
    auto row = [&](double* dst, const VD (&v)[C]) {
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) st_idx_if(vi_lt(c.stage(j), n), T, c.stage(j), v[j]);
        warp_sync();
        warp_copy(dst, T, n);
    };
    if (way) {
        if (si.status == ACMPC_SOLVED) row(way + 6 * n, vel);
        if (vel_out) row(vel_out, vel);
    } else {
        AC_UNROLL
        for (int j = 0; j < C; ++j)
            vel[j] = (si.status == ACMPC_SOLVED) ? vsel(vi_lt(c.stage(j), n), vel[j], VD(0.0)) : VD(0.0);
        row(vel_out, vel);
        if (o.waypoints) {
            row(o.waypoints + 0 * n, path.xs), row(o.waypoints + 1 * n, path.ys), row(o.waypoints + 2 * n, path.psi);
            row(o.waypoints + 3 * n, path.kap), row(o.waypoints + 4 * n, path.dist), row(o.waypoints + 5 * n, path.wid);
            row(o.waypoints + 6 * n, vel);
        }
        if (o.v_ref && o.v_ref != vel_out) row(o.v_ref, vel);
    }
    AC_PHASE(c, 14);   // outputs
    AC_LANE0
    {
        if (o.status_speed) *o.status_speed = si.status;
        if (o.iters) o.iters[0] = si.iter;
        if (o.rho_updates) o.rho_updates[0] = si.rho_updates;
    }
    return si.iter;   // ADMM iterations of the speed-profile QP (a scheduling hint for phase 2)
}

// Phase 2: control QP, unpack, rollout, cost (spatial_mpc.py:186-212).  The waypoints are rebuilt from the
// staged raw path (cheaper than handing six rows per stage over); `vel_in` [n] is phase 1's profile.
template <int C>
AC_DEV void control_instance(const Ctx<C>& c, const double* raw_path, const double* vel_in, double offset,
                             const InstanceOut& o, double* warm = nullptr, bool use_warm = false)
{
    const int n = c.n, H = c.H;
    PathRegs<C> path;
    build_waypoints<C>(c, raw_path, path);
    warp_sync();
    SolveInfo ci;
    VD vel[C];
    AC_UNROLL
    for (int j = 0; j < C; ++j) {
        vel[j] = ld_idx_if(vi_lt(c.stage(j), n), vel_in, c.stage(j), 0.0);
        c.st(F_XS, j, path.xs[j]), c.st(F_YS, j, path.ys[j]), c.st(F_PSI, j, path.psi[j]);
    }
    ControlQP<C> cq(c);
    AC_PHASE(c, 0);   // waypoints
    cq.setup(path, vel, offset);
    AC_PHASE(c, 1);   // setup (assembly + Ruiz)
    cq.solve(ci, warm, use_warm);
    AC_PHASE(c, 6);   // end of solve (obj, warm store)
    // unpack (spatial_mpc.py:193-212) and roll out (dynamics.py:42-63)
    const double L = c.cfg->wheelbase;
    VD un[C][3], uv[C], uk[C];
    AC_UNROLL
    for (int j = 0; j < C; ++j) {
        un[j][0] = cq.x[j][0] / c.ld(K_DI + 0, j);                  // e_y
        un[j][1] = cq.x[j][1] / c.ld(K_DI + 1, j);                  // e_psi
        un[j][2] = cq.x[j][2] / c.ld(K_DI + 2, j);                  // t
        uv[j] = cq.x[j][3] / c.ld(K_DI + 3, j);                     // v        (u_{s-1} lives with stage s)
        uk[j] = vatan((cq.x[j][4] / c.ld(K_DI + 4, j)) * VD(L));    // delta = atan(kappa_cmd L)
    }
    // every output array leaves as coalesced rows through a shared-memory tile (see speed_instance); the scan region
    // is dead after the solve
    double* T = c.W;
    if (o.states) {
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB isx = vi_lt(st, H);
            st_idx_if(isx, T, st * 3, un[j][0]), st_idx_if(isx, T, st * 3 + 1, un[j][1]), st_idx_if(isx, T, st * 3 + 2, un[j][2]);
        }
        warp_sync();
        warp_copy(o.states, T, 3 * H);
    }
    if (o.controls) {
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB hasU = vi_ge(st, 1) & vi_lt(st, H);
            st_idx_if(hasU, T, st + (-1), uv[j]), st_idx_if(hasU, T, st + (n - 1), uk[j]);
        }
        warp_sync();
        warp_copy(o.controls, T, 2 * n);
    }
    if (o.prediction) {
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB ok = vi_lt(st, n);
            VD ps = c.ld(F_PSI, j);
            st_idx_if(ok, T, st * 2, c.ld(F_XS, j) - un[j][0] * vsin(ps));
            st_idx_if(ok, T, st * 2 + 1, c.ld(F_YS, j) + un[j][0] * vcos(ps));
        }
        warp_sync();
        warp_copy(o.prediction, T, 2 * n);
    }
    if (o.cum_time) {
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) st_idx_if(vi_lt(c.stage(j), n), T, c.stage(j), un[j][2]);
        warp_sync();
        warp_copy(o.cum_time, T, n);
    }
    if (o.derived) {   // spatial_mpc.py:208-211 over x_0 .. x_{n-1}: times, accelerations (sic: e_y column), steer_rates
        VD nx[C][3];
        pull_next_k<C, 3>(un, nx);
        warp_sync();
        AC_UNROLL
        for (int j = 0; j < C; ++j) {
            VI st = c.stage(j);
            VB ok = vi_le(st, n - 2);
            VD dt = nx[j][2] - un[j][2];
            st_idx_if(ok, T, st, dt);
            st_idx_if(ok, T, st + (n - 1), (nx[j][0] - un[j][0]) / dt);
            st_idx_if(ok, T, st + 2 * (n - 1), (nx[j][1] - un[j][1]) / dt);
        }
        warp_sync();
        warp_copy(o.derived, T, 3 * (n - 1));
    }
    warp_sync();   // the tile is the next instance's staging area
    AC_PHASE(c, 7);   // outputs
    AC_LANE0
    {
        if (o.cost) *o.cost = ci.obj_val;
        if (o.pri_res) *o.pri_res = ci.pri_res;
        if (o.dua_res) *o.dua_res = ci.dua_res;
        if (o.status) *o.status = ci.status;
        if (o.iters) o.iters[1] = ci.iter;
        if (o.rho_updates) o.rho_updates[1] = ci.rho_updates;
    }
}

}  // namespace acmpc
