// mpc_body.cuh -- one warp solves one MPC step (SpatialMPC.get_control) entirely on chip.
//
// Reference path being replaced (file:line under /root/reference/src/acmpc/control/):
//   spatial_mpc.py:125-154   construct_waypoints          -> build_waypoints()
//   solvers/speed_profile.py:26-59,131-150 + osqp        -> SpeedQP
//   dynamics.py:23-40 (t2s), :65-103 (linearise)          -> ControlQP::assemble()
//   solvers/control.py:26-79,121-158 + osqp               -> ControlQP
//   spatial_mpc.py:193-212, dynamics.py:42-63 (s2t)       -> write_outputs()
//
// Design: all per-instance state lives in shared memory as structure-of-arrays
// indexed by horizon stage; lane l of the warp owns stages l, l+32, ...  The OSQP iteration is
// carried out in the *reduced* form  (P + sigma I + A' diag(rho) A) x~ = sigma x - q + A'(rho z - y),
// z~ = A x~,  exploiting the stage structure of A instead of a general sparse LDL':
//   * control QP: inputs u_k are eliminated analytically (their block is diagonal), leaving an SPD
//     block-tridiagonal system with 3x3 blocks over the states, factorised once per rho
//     (block LDL' with explicit 3x3 inverses) and solved by two short serial sweeps;
//   * speed QP: the reduced matrix is scalar tridiagonal; its explicit inverse is formed once per rho
//     (one lane per column) so every iteration is a conflict-free dense mat-vec.
// Everything else (Ruiz equilibration, rho classes, relaxation, projection, dual update, unscaled
// residuals, infeasibility certificates, adaptive rho) follows OSQP 0.6.x step for step, because
// the reference's answer is defined by OSQP's iterate at its termination check.
//
// The same source is compiled
//   (a) by nvcc for sm_100a into the product library (32 lanes per QP) -- the ONLY product path;
//   (b) by g++ with -DACMPC_EMULATE (ONE lane, plain loops) into tests/_emul/libacmpc_emul.so, a
//       test-only aid to debug the arithmetic in a GPU-less container.  It is never loaded by the
//       package and is not a fallback.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/acmpc_b200.h"

#ifdef ACMPC_EMULATE
#define ACMPC_DEV static inline
#define ACMPC_MEM inline
#define ACMPC_LANES 1
#define ACMPC_SYNC() ((void)0)
#else
#define ACMPC_DEV __device__ __forceinline__
#define ACMPC_MEM __device__ __forceinline__
#define ACMPC_LANES 32
#define ACMPC_SYNC() __syncwarp()
#endif

namespace acmpc {

constexpr double kInfty = 1e30;
constexpr double kMinScaling = 1e-4;
constexpr double kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6;
constexpr double kRhoMax = 1e6;
constexpr double kRhoEqOverIneq = 1e3;
constexpr double kRhoTol = 1e-4;
constexpr double kBig = kInfty * kMinScaling;  // "infinite bound" threshold in scaled space
constexpr double kPi = 3.14159265358979323846;
// stages owned by one lane (the one-lane emulation owns them all)
constexpr int kMaxStagesPerLane = (ACMPC_LANES == 1) ? ACMPC_MAX_HORIZON : (ACMPC_MAX_HORIZON + 31) / 32;

// ------------------------------------------------------------------------------------------------
// shared-memory field map: every field is an array of Hs doubles (one per horizon stage)
// ------------------------------------------------------------------------------------------------
enum : int {
    // ReferencePath rows (paths.py:4-72) -- live for the whole step
    F_XS = 0, F_YS, F_PSI, F_KAP, F_DIST, F_WID, F_VEL,
    F_PATH_END,
    // ---- control QP (aliases the speed-QP region below) ----
    C_M = F_PATH_END,       // 3: coefficient of x_k in its own equality block k (raw -1)
    C_A11 = C_M + 3, C_A12, C_A21, C_A22, C_A31, C_A33, C_B22, C_B31,  // block k+1, columns of stage k
    C_S,                    // 5: identity (bound) rows
    C_P = C_S + 5,          // 5: diagonal of P
    C_Q = C_P + 5,          // 2: q of (v, kappa_cmd); the state part of q is exactly zero
    C_DI = C_Q + 2,         // 5: 1/D
    C_EEI = C_DI + 5,       // 3: 1/E of equality block k
    C_EBI = C_EEI + 3,      // 5: 1/E of bound rows
    C_LB = C_EBI + 5,       // 5
    C_UB = C_LB + 5,        // 5
    C_BE = C_UB + 5,        // 3: scaled equality rhs (l == u)
    C_X = C_BE + 3,         // 5
    C_ZB = C_X + 5,         // 5
    C_YB = C_ZB + 5,        // 5
    C_YE = C_YB + 5,        // 3
    C_ZE = C_YE + 3,        // 3
    C_SI = C_ZE + 3,        // 6: Sigma_k^{-1} (symmetric 3x3: 00 10 11 20 21 22)
    C_G = C_SI + 6,         // 9: G_k = S_{k,k-1} Sigma_{k-1}^{-1} (row major)
    C_IV = C_G + 9,         // 1/K_uu(v)
    C_IK,                   // 1/K_uu(kappa_cmd)
    C_R,                    // 5: rhs -> x~
    C_T = C_R + 5,          // 13: scratch (neighbour exchange, delta_x / delta_y at checks)
    C_TYPE = C_T + 13,      // bound-row classes, 2 bits per row, stored as a double
    C_END,
    // ---- speed-profile QP ----
    S_AL = F_PATH_END,      // coefficient of v_i in acceleration row i
    S_AU,                   // coefficient of v_{i+1} in acceleration row i
    S_S, S_P, S_Q, S_DI, S_EAI, S_EBI, S_LA, S_UA, S_LB, S_UB,
    S_X, S_ZA, S_ZB, S_YA, S_YB, S_R, S_T0, S_T1, S_T2, S_DX, S_DYA, S_DYB, S_W, S_TYPE,
    S_KINV,                 // n*n doubles follow (explicit inverse of the reduced matrix)
    S_END_FIXED = S_KINV
};

constexpr int kFieldsPerStage = C_END;

ACMPC_DEV int smem_doubles(int H) { return kFieldsPerStage * H; }

// ------------------------------------------------------------------------------------------------
// lane helpers
// ------------------------------------------------------------------------------------------------
#ifdef ACMPC_EMULATE
ACMPC_DEV double warp_max(double v) { return v; }
ACMPC_DEV double warp_sum(double v) { return v; }
ACMPC_DEV int warp_any(int p) { return p; }
#else
ACMPC_DEV double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
ACMPC_DEV double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
ACMPC_DEV int warp_any(int p) { return __any_sync(0xffffffffu, p); }
#endif

ACMPC_DEV double np_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}
ACMPC_DEV double limit_scaling(double v)
{
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}
ACMPC_DEV double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// OSQP constraint classes (set_rho_vec): 0 inequality, 1 equality, 2 loose
ACMPC_DEV int row_class(double l, double u)
{
    if (l < -kBig && u > kBig) return 2;
    if (u - l < kRhoTol) return 1;
    return 0;
}

struct RhoSet {
    double rho, rho_eq, rinv, rinv_eq;
    ACMPC_MEM void set(double r)
    {
        rho = r;
        rho_eq = kRhoEqOverIneq * r;
        rinv = 1.0 / rho;
        rinv_eq = 1.0 / rho_eq;
    }
    ACMPC_MEM double of(int cls) const { return cls == 0 ? rho : (cls == 1 ? rho_eq : kRhoMin); }
    ACMPC_MEM double inv_of(int cls) const { return cls == 0 ? rinv : (cls == 1 ? rinv_eq : 1.0 / kRhoMin); }
};

// norms gathered at a termination check (update_info + compute_*_tol + compute_rho_estimate)
struct Norms {
    // unscaled (termination)
    double pri, dua, nz, nAx, nq, nAty, nPx;
    // scaled (rho estimate)
    double s_pri, s_dua, s_z, s_Ax, s_q, s_Aty, s_Px;
};

struct SolveInfo {
    int status, iter, rho_updates;
    double pri_res, dua_res, obj_val;
};

ACMPC_DEV double rho_estimate(const Norms& N, double rho)
{
    double p = N.s_pri / (fmax(N.s_z, N.s_Ax) + 1e-10);
    double d = N.s_dua / (fmax(N.s_q, fmax(N.s_Aty, N.s_Px)) + 1e-10);
    double r = rho * sqrt(p / (d + 1e-10));
    return clampd(r, kRhoMin, kRhoMax);
}

// OSQP overwrites info.obj_val when a certificate fires (check_termination): +-OSQP_INFTY, NaN if non-convex
ACMPC_DEV double final_obj(int status, double obj)
{
    if (status == ACMPC_PRIMAL_INFEASIBLE || status == ACMPC_PRIMAL_INFEASIBLE_INACCURATE) return kInfty;
    if (status == ACMPC_DUAL_INFEASIBLE || status == ACMPC_DUAL_INFEASIBLE_INACCURATE) return -kInfty;
    if (status == ACMPC_NON_CVX) return NAN;
    return obj;
}

// ------------------------------------------------------------------------------------------------
// per-QP context
// ------------------------------------------------------------------------------------------------
struct Ctx {
    double* S;   // base of this QP's shared-memory block
    int H, n, Hs, lane;
    const acmpc_config* cfg;
    ACMPC_MEM double* f(int field) const { return S + field * Hs; }
};

// spatial_mpc.py:125-154.  `W` = raw (H,3) path staged at S_RAW (aliases the QP region).
ACMPC_DEV void build_waypoints(const Ctx& c, const double* W)
{
    const int n = c.n, H = c.H;
    double *xs = c.f(F_XS), *ys = c.f(F_YS), *psi = c.f(F_PSI), *kap = c.f(F_KAP), *dist = c.f(F_DIST),
           *wid = c.f(F_WID), *vel = c.f(F_VEL);
    for (int i = c.lane; i < n; i += ACMPC_LANES) {
        const double* cur = W + 3 * i;
        const double* nxt = W + 3 * (i + 1);
        const double* prv = (i == 0) ? W + 3 * (H - 1) : W + 3 * (i - 1);
        double ax = nxt[0] - cur[0], ay = nxt[1] - cur[1];
        double bx = cur[0] - prv[0], by = cur[1] - prv[1];
        double p = atan2(ay, ax);
        double d = sqrt(ax * ax + ay * ay);
        double behind = atan2(by, bx);
        double dang = np_mod(p - behind + kPi, 2.0 * kPi) - kPi;
        xs[i] = cur[0];
        ys[i] = cur[1];
        wid[i] = nxt[2];
        psi[i] = p;
        dist[i] = d;
        kap[i] = dang / (d + 1e-12) + 1e-12;
        vel[i] = 0.0;
    }
    ACMPC_SYNC();
    if (c.lane == 0) kap[0] = kap[1];
    ACMPC_SYNC();
}

// ================================================================================================
// Speed-profile QP   min 1/2 v'v - vbar'v   s.t.  a_min <= (v_{i+1}-v_i)/(2 d_i) <= a_max,
//                                                  v_min <= v_i <= vbar_i          (speed_profile.py)
// rows: acceleration row i (i = 0..n-2) is owned by stage i, bound row i by stage i.
// ================================================================================================
struct SpeedQP {
    const Ctx& c;
    int n;
    bool use_inverse;
    double cs, cinv;   // cost scaling
    RhoSet R;
    double nq_unscaled, nq_scaled;

    ACMPC_MEM explicit SpeedQP(const Ctx& ctx) : c(ctx), n(ctx.n)
    {
        use_inverse = (S_KINV + 0) * c.Hs + n * n <= kFieldsPerStage * c.Hs;
    }

    ACMPC_MEM int cls_a(int i) const { return ((int)c.f(S_TYPE)[i]) & 3; }
    ACMPC_MEM int cls_b(int i) const { return (((int)c.f(S_TYPE)[i]) >> 2) & 3; }

    // speed_profile.py:26-59 / :131-150, then OSQP scale_data + set_rho_vec
    ACMPC_MEM void assemble_and_scale(double v_max_live, int localised)
    {
        const acmpc_config& g = *c.cfg;
        const double *kap = c.f(F_KAP), *dist = c.f(F_DIST);
        double *AL = c.f(S_AL), *AU = c.f(S_AU), *SS = c.f(S_S), *P = c.f(S_P), *Q = c.f(S_Q);
        double *DI = c.f(S_DI), *EAI = c.f(S_EAI), *EBI = c.f(S_EBI);
        double *LA = c.f(S_LA), *UA = c.f(S_UA), *LB = c.f(S_LB), *UB = c.f(S_UB), *W = c.f(S_W);
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            double vb;
            if (localised) {
                vb = v_max_live;
            } else {
                double ak = fabs(kap[i]);
                double vdyn = sqrt(g.ay_max / (ak + 1e-12));
                if (ak < g.ki_min) vdyn = v_max_live;
                double v = vdyn < v_max_live ? vdyn : v_max_live;
                v = g.v_min > v ? g.v_min : v;
                vb = v + 2.0;
                if (g.has_end_velocity && i == n - 1) vb = g.end_velocity;
            }
            bool has_row = i <= n - 2;
            AL[i] = has_row ? -1.0 / (2.0 * dist[i]) : 0.0;
            AU[i] = has_row ? 1.0 / (2.0 * dist[i]) : 0.0;
            SS[i] = 1.0;
            P[i] = 1.0;
            Q[i] = -1.0 * vb;
            DI[i] = 1.0;    // holds D during the Ruiz loop, inverted at the end
            EAI[i] = 1.0;   // E (acceleration rows)
            EBI[i] = 1.0;   // E (bound rows)
            LA[i] = g.a_min, UA[i] = g.a_max;
            LB[i] = g.v_min, UB[i] = vb;
        }
        ACMPC_SYNC();
        cs = 1.0;
        for (int pass = 0; pass < g.scaling; ++pass) {
            // column / row inf-norms of [P A'; A 0]
            for (int i = c.lane; i < n; i += ACMPC_LANES) {
                double dn = fmax(fabs(P[i]), fabs(SS[i]));
                if (i <= n - 2) dn = fmax(dn, fabs(AL[i]));
                if (i >= 1) dn = fmax(dn, fabs(AU[i - 1]));
                W[i] = 1.0 / sqrt(limit_scaling(dn));
            }
            ACMPC_SYNC();
            double psum = 0.0, qmax = 0.0;
            for (int i = c.lane; i < n; i += ACMPC_LANES) {
                double d = W[i];
                if (i <= n - 2) {
                    double ea = 1.0 / sqrt(limit_scaling(fmax(fabs(AL[i]), fabs(AU[i]))));
                    AL[i] = (AL[i] * ea) * d;
                    AU[i] = (AU[i] * ea) * W[i + 1];
                    EAI[i] *= ea;
                }
                double eb = 1.0 / sqrt(limit_scaling(fabs(SS[i])));
                SS[i] = (SS[i] * eb) * d;
                EBI[i] *= eb;
                P[i] = (P[i] * d) * d;
                Q[i] *= d;
                DI[i] *= d;
                psum += fabs(P[i]);
                qmax = fmax(qmax, fabs(Q[i]));
            }
            psum = warp_sum(psum);
            qmax = warp_max(qmax);
            double ct = fmax(psum / (double)n, limit_scaling(qmax));
            ct = 1.0 / limit_scaling(ct);
            for (int i = c.lane; i < n; i += ACMPC_LANES) {
                P[i] *= ct;
                Q[i] *= ct;
            }
            cs *= ct;
            ACMPC_SYNC();
        }
        cinv = 1.0 / cs;
        // scale bounds, classify rows, invert D/E, norm of q
        double nqu = 0.0, nqs = 0.0;
        double* TY = c.f(S_TYPE);
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            int ca = 0;
            if (i <= n - 2) {
                LA[i] *= EAI[i], UA[i] *= EAI[i];
                ca = row_class(LA[i], UA[i]);
            }
            LB[i] *= EBI[i], UB[i] *= EBI[i];
            int cb = row_class(LB[i], UB[i]);
            TY[i] = (double)(ca | (cb << 2));
            DI[i] = 1.0 / DI[i];
            EAI[i] = 1.0 / EAI[i];
            EBI[i] = 1.0 / EBI[i];
            nqu = fmax(nqu, fabs(DI[i] * Q[i]));
            nqs = fmax(nqs, fabs(Q[i]));
        }
        nq_unscaled = warp_max(nqu);
        nq_scaled = warp_max(nqs);
        ACMPC_SYNC();
    }

    // reduced matrix K = P + sigma + A' rho A (tridiagonal) -> LDL' -> explicit inverse
    ACMPC_MEM void factor()
    {
        const double sigma = c.cfg->sigma;
        const double *AL = c.f(S_AL), *AU = c.f(S_AU), *SS = c.f(S_S), *P = c.f(S_P);
        double *Ld = c.f(S_T0), *Dn = c.f(S_T1), *Od = c.f(S_T2), *Kinv = c.f(S_KINV);
        // diagonal into Dn, off-diagonal (i,i+1) into Od
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            double d = P[i] + sigma + R.of(cls_b(i)) * SS[i] * SS[i];
            if (i <= n - 2) d += R.of(cls_a(i)) * AL[i] * AL[i];
            if (i >= 1) d += R.of(cls_a(i - 1)) * AU[i - 1] * AU[i - 1];
            Dn[i] = d;
            Od[i] = (i <= n - 2) ? R.of(cls_a(i)) * AL[i] * AU[i] : 0.0;
        }
        ACMPC_SYNC();
        if (c.lane == 0) {
            // L has unit diagonal and sub-diagonal Ld[i] (i>=1); Dn becomes 1/pivot
            double piv = Dn[0];
            Dn[0] = 1.0 / piv;
            for (int i = 1; i < n; ++i) {
                double l = Od[i - 1] * Dn[i - 1];
                Ld[i] = l;
                piv = Dn[i] - l * Od[i - 1];
                Dn[i] = 1.0 / piv;
            }
        }
        ACMPC_SYNC();
        if (use_inverse) {
            for (int j = c.lane; j < n; j += ACMPC_LANES) {
                double* col = Kinv + j * n;
                double y = 1.0;
                for (int i = 0; i < j; ++i) col[i] = 0.0;
                col[j] = y;
                for (int i = j + 1; i < n; ++i) {
                    y = -Ld[i] * y;
                    col[i] = y;
                }
                double x = col[n - 1] * Dn[n - 1];
                col[n - 1] = x;
                for (int i = n - 2; i >= 0; --i) {
                    x = col[i] * Dn[i] - Ld[i + 1] * x;
                    col[i] = x;
                }
            }
            ACMPC_SYNC();
        }
    }

    // x~ = K^{-1} r  (r and x~ both in S_R)
    ACMPC_MEM void kkt_solve()
    {
        double* Rv = c.f(S_R);
        if (use_inverse) {
            const double* Kinv = c.f(S_KINV);
            double acc[kMaxStagesPerLane];
            int cnt = 0;
            for (int i = c.lane; i < n; i += ACMPC_LANES) {
                double a0 = 0.0, a1 = 0.0;
                int j = 0;
                for (; j + 1 < n; j += 2) {
                    a0 += Kinv[j * n + i] * Rv[j];
                    a1 += Kinv[(j + 1) * n + i] * Rv[j + 1];
                }
                if (j < n) a0 += Kinv[j * n + i] * Rv[j];
                acc[cnt++] = a0 + a1;
            }
            ACMPC_SYNC();
            cnt = 0;
            for (int i = c.lane; i < n; i += ACMPC_LANES) Rv[i] = acc[cnt++];
            ACMPC_SYNC();
        } else {
            const double *Ld = c.f(S_T0), *Dn = c.f(S_T1);
            if (c.lane == 0) {
                for (int i = 1; i < n; ++i) Rv[i] -= Ld[i] * Rv[i - 1];
                Rv[n - 1] *= Dn[n - 1];
                for (int i = n - 2; i >= 0; --i) Rv[i] = Rv[i] * Dn[i] - Ld[i + 1] * Rv[i + 1];
            }
            ACMPC_SYNC();
        }
    }

    ACMPC_MEM void compute_norms(Norms& N)
    {
        const double *AL = c.f(S_AL), *AU = c.f(S_AU), *SS = c.f(S_S), *P = c.f(S_P), *Q = c.f(S_Q);
        const double *DI = c.f(S_DI), *EAI = c.f(S_EAI), *EBI = c.f(S_EBI);
        const double *X = c.f(S_X), *ZA = c.f(S_ZA), *ZB = c.f(S_ZB), *YA = c.f(S_YA), *YB = c.f(S_YB);
        double v[12];
        for (int t = 0; t < 12; ++t) v[t] = 0.0;
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            double aty = SS[i] * YB[i];
            if (i <= n - 2) {
                double ax = AL[i] * X[i] + AU[i] * X[i + 1];
                double r = ax - ZA[i];
                v[0] = fmax(v[0], fabs(EAI[i] * r)), v[7] = fmax(v[7], fabs(r));
                v[1] = fmax(v[1], fabs(EAI[i] * ZA[i])), v[8] = fmax(v[8], fabs(ZA[i]));
                v[2] = fmax(v[2], fabs(EAI[i] * ax)), v[9] = fmax(v[9], fabs(ax));
                aty += AL[i] * YA[i];
            }
            if (i >= 1) aty += AU[i - 1] * YA[i - 1];
            double ax = SS[i] * X[i];
            double r = ax - ZB[i];
            v[0] = fmax(v[0], fabs(EBI[i] * r)), v[7] = fmax(v[7], fabs(r));
            v[1] = fmax(v[1], fabs(EBI[i] * ZB[i])), v[8] = fmax(v[8], fabs(ZB[i]));
            v[2] = fmax(v[2], fabs(EBI[i] * ax)), v[9] = fmax(v[9], fabs(ax));
            double px = P[i] * X[i];
            double dr = Q[i] + px + aty;
            v[3] = fmax(v[3], fabs(DI[i] * dr)), v[6] = fmax(v[6], fabs(dr));
            v[4] = fmax(v[4], fabs(DI[i] * aty)), v[10] = fmax(v[10], fabs(aty));
            v[5] = fmax(v[5], fabs(DI[i] * px)), v[11] = fmax(v[11], fabs(px));
        }
        for (int t = 0; t < 12; ++t) v[t] = warp_max(v[t]);
        N.pri = v[0], N.nz = v[1], N.nAx = v[2];
        N.dua = cinv * v[3], N.nAty = v[4], N.nPx = v[5], N.nq = nq_unscaled;
        N.s_dua = v[6], N.s_pri = v[7], N.s_z = v[8], N.s_Ax = v[9], N.s_Aty = v[10], N.s_Px = v[11];
        N.s_q = nq_scaled;
    }

    // is_primal_infeasible: delta_y in S_DYA (accel rows) / S_DYB (bound rows)
    ACMPC_MEM int primal_infeasible(double eps)
    {
        const double *AL = c.f(S_AL), *AU = c.f(S_AU), *SS = c.f(S_S), *DI = c.f(S_DI);
        const double *EAI = c.f(S_EAI), *EBI = c.f(S_EBI);
        const double *LA = c.f(S_LA), *UA = c.f(S_UA), *LB = c.f(S_LB), *UB = c.f(S_UB);
        double *DYA = c.f(S_DYA), *DYB = c.f(S_DYB);
        double nrm = 0.0, lhs = 0.0;
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            if (i <= n - 2) {
                double dy = DYA[i];
                if (UA[i] > kBig) dy = (LA[i] < -kBig) ? 0.0 : fmin(dy, 0.0);
                else if (LA[i] < -kBig) dy = fmax(dy, 0.0);
                DYA[i] = dy;
                nrm = fmax(nrm, fabs(dy / EAI[i]));
                lhs += UA[i] * fmax(dy, 0.0) + LA[i] * fmin(dy, 0.0);
            }
            double dy = DYB[i];
            if (UB[i] > kBig) dy = (LB[i] < -kBig) ? 0.0 : fmin(dy, 0.0);
            else if (LB[i] < -kBig) dy = fmax(dy, 0.0);
            DYB[i] = dy;
            nrm = fmax(nrm, fabs(dy / EBI[i]));
            lhs += UB[i] * fmax(dy, 0.0) + LB[i] * fmin(dy, 0.0);
        }
        nrm = warp_max(nrm);
        lhs = warp_sum(lhs);
        ACMPC_SYNC();
        if (!(nrm > eps) || !(lhs < -eps * nrm)) return 0;
        double m = 0.0;
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            double a = SS[i] * DYB[i];
            if (i <= n - 2) a += AL[i] * DYA[i];
            if (i >= 1) a += AU[i - 1] * DYA[i - 1];
            m = fmax(m, fabs(DI[i] * a));
        }
        m = warp_max(m);
        return m < eps * nrm;
    }

    // is_dual_infeasible: delta_x in S_DX
    ACMPC_MEM int dual_infeasible(double eps)
    {
        const double *AL = c.f(S_AL), *AU = c.f(S_AU), *SS = c.f(S_S), *P = c.f(S_P), *Q = c.f(S_Q);
        const double *DI = c.f(S_DI), *EAI = c.f(S_EAI), *EBI = c.f(S_EBI);
        const double *LA = c.f(S_LA), *UA = c.f(S_UA), *LB = c.f(S_LB), *UB = c.f(S_UB);
        const double* DX = c.f(S_DX);
        double nrm = 0.0, qdx = 0.0, pm = 0.0;
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            nrm = fmax(nrm, fabs(DX[i] / DI[i]));
            qdx += Q[i] * DX[i];
            pm = fmax(pm, fabs(DI[i] * (P[i] * DX[i])));
        }
        nrm = warp_max(nrm), qdx = warp_sum(qdx), pm = warp_max(pm);
        if (!(nrm > eps) || !(qdx < -cs * eps * nrm) || !(pm < cs * eps * nrm)) return 0;
        int bad = 0;
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            if (i <= n - 2) {
                double a = EAI[i] * (AL[i] * DX[i] + AU[i] * DX[i + 1]);
                if ((UA[i] < kBig && a > eps * nrm) || (LA[i] > -kBig && a < -eps * nrm)) bad = 1;
            }
            double a = EBI[i] * (SS[i] * DX[i]);
            if ((UB[i] < kBig && a > eps * nrm) || (LB[i] > -kBig && a < -eps * nrm)) bad = 1;
        }
        return !warp_any(bad);
    }

    // check_termination(work, approximate): returns status or 0 (continue)
    ACMPC_MEM int check(const Norms& N, int approximate)
    {
        const acmpc_config& g = *c.cfg;
        double k = approximate ? 10.0 : 1.0;
        if (N.pri > kInfty || N.dua > kInfty) return ACMPC_NON_CVX;
        double eps_p = k * g.eps_abs + k * g.eps_rel * fmax(N.nz, N.nAx);
        double eps_d = k * g.eps_abs + k * g.eps_rel * cinv * fmax(N.nq, fmax(N.nAty, N.nPx));
        int p_ok = N.pri < eps_p, d_ok = N.dua < eps_d;
        int p_inf = 0, d_inf = 0;
        if (!p_ok) p_inf = primal_infeasible(k * g.eps_prim_inf);
        if (!d_ok) d_inf = dual_infeasible(k * g.eps_dual_inf);
        if (p_ok && d_ok) return approximate ? ACMPC_SOLVED_INACCURATE : ACMPC_SOLVED;
        if (p_inf) return approximate ? ACMPC_PRIMAL_INFEASIBLE_INACCURATE : ACMPC_PRIMAL_INFEASIBLE;
        if (d_inf) return approximate ? ACMPC_DUAL_INFEASIBLE_INACCURATE : ACMPC_DUAL_INFEASIBLE;
        return 0;
    }

    // osqp_solve, cold start.  Result (unscaled v) is left in S_R.
    ACMPC_MEM void solve(SolveInfo& info)
    {
        const acmpc_config& g = *c.cfg;
        const double alpha = g.alpha, sigma = g.sigma;
        const double *AL = c.f(S_AL), *AU = c.f(S_AU), *SS = c.f(S_S), *Q = c.f(S_Q);
        const double *LA = c.f(S_LA), *UA = c.f(S_UA), *LB = c.f(S_LB), *UB = c.f(S_UB);
        double *X = c.f(S_X), *ZA = c.f(S_ZA), *ZB = c.f(S_ZB), *YA = c.f(S_YA), *YB = c.f(S_YB);
        double *Rv = c.f(S_R), *W = c.f(S_W);
        double *DX = c.f(S_DX), *DYA = c.f(S_DYA), *DYB = c.f(S_DYB);
        R.set(clampd(g.rho, kRhoMin, kRhoMax));
        for (int i = c.lane; i < n; i += ACMPC_LANES) X[i] = ZA[i] = ZB[i] = YA[i] = YB[i] = 0.0;
        ACMPC_SYNC();
        factor();
        Norms N;
        int status = 0, iter = 0, updates = 0, checked = 0;
        for (iter = 1; iter <= g.max_iter; ++iter) {
            // w_a = rho z - y on acceleration rows (needed by the neighbour column)
            for (int i = c.lane; i < n; i += ACMPC_LANES)
                W[i] = (i <= n - 2) ? R.of(cls_a(i)) * ZA[i] - YA[i] : 0.0;
            ACMPC_SYNC();
            for (int i = c.lane; i < n; i += ACMPC_LANES) {
                double r = sigma * X[i] - Q[i] + SS[i] * (R.of(cls_b(i)) * ZB[i] - YB[i]) + AL[i] * W[i];
                if (i >= 1) r += AU[i - 1] * W[i - 1];
                Rv[i] = r;
            }
            ACMPC_SYNC();
            kkt_solve();
            checked = (g.check_termination > 0 && iter % g.check_termination == 0);
            const bool keep_delta = checked || iter == g.max_iter;
            for (int i = c.lane; i < n; i += ACMPC_LANES) {
                double xt = Rv[i];
                if (i <= n - 2) {
                    int cl = cls_a(i);
                    double zt = AL[i] * xt + AU[i] * Rv[i + 1];
                    double zh = alpha * zt + (1.0 - alpha) * ZA[i];
                    double zn = clampd(zh + R.inv_of(cl) * YA[i], LA[i], UA[i]);
                    double dy = R.of(cl) * (zh - zn);
                    ZA[i] = zn;
                    YA[i] += dy;
                    if (keep_delta) DYA[i] = dy;
                }
                {
                    int cl = cls_b(i);
                    double zt = SS[i] * xt;
                    double zh = alpha * zt + (1.0 - alpha) * ZB[i];
                    double zn = clampd(zh + R.inv_of(cl) * YB[i], LB[i], UB[i]);
                    double dy = R.of(cl) * (zh - zn);
                    ZB[i] = zn;
                    YB[i] += dy;
                    if (keep_delta) DYB[i] = dy;
                }
                double xn = alpha * xt + (1.0 - alpha) * X[i];
                if (keep_delta) DX[i] = xn - X[i];
                X[i] = xn;
            }
            ACMPC_SYNC();
            if (checked) {
                compute_norms(N);
                status = check(N, 0);
                if (status) break;
            }
            if (g.adaptive_rho && g.adaptive_rho_interval > 0 && iter % g.adaptive_rho_interval == 0) {
                if (!checked) compute_norms(N);
                double rn = rho_estimate(N, R.rho);
                if (rn > R.rho * g.adaptive_rho_tolerance || rn < R.rho / g.adaptive_rho_tolerance) {
                    R.set(rn);
                    ++updates;
                    ACMPC_SYNC();
                    factor();
                }
            }
        }
        if (iter > g.max_iter) iter = g.max_iter;
        if (!checked) {
            compute_norms(N);
            status = check(N, 0);
        }
        if (!status) {
            status = check(N, 1);
            if (!status) status = ACMPC_MAX_ITER_REACHED;
        }
        info.status = status, info.iter = iter, info.rho_updates = updates;
        info.pri_res = N.pri, info.dua_res = N.dua;
        // unscale: x = D x_s
        const double* DI = c.f(S_DI);
        double obj = 0.0;
        const double* P = c.f(S_P);
        for (int i = c.lane; i < n; i += ACMPC_LANES) {
            obj += 0.5 * P[i] * X[i] * X[i] + Q[i] * X[i];
            Rv[i] = X[i] / DI[i];
        }
        info.obj_val = final_obj(status, warp_sum(obj) * cinv);
        ACMPC_SYNC();
    }
};

// ================================================================================================
// Control QP (solvers/control.py + dynamics.py:65-103), decision vector per stage k:
//   x_k = (e_y, e_psi, t), u_k = (v, kappa_cmd); stage n = H-1 has x_n only.
// rows: equality block k (3 rows, "-x_k + A_{k-1} x_{k-1} + B_{k-1} u_{k-1}"; block 0 is -x_0 = -x0)
//       is owned by stage k; its entries on stage k-1's variables are stored with stage k-1.
//       bound rows: one per variable, owned by the variable's stage.
// ================================================================================================
struct ControlQP {
    const Ctx& c;
    int n, H;
    double cs, cinv;
    RhoSet R;
    double nq_unscaled, nq_scaled;

    ACMPC_MEM explicit ControlQP(const Ctx& ctx) : c(ctx), n(ctx.n), H(ctx.H) {}

    ACMPC_MEM int nvar(int k) const { return k < n ? 5 : 3; }
    ACMPC_MEM int cls_b(int k, int j) const { return (((int)c.f(C_TYPE)[k]) >> (2 * j)) & 3; }

    // linearise + stack (dynamics.py:65-103, solvers/control.py:26-79), raw (unscaled) data
    ACMPC_MEM void assemble(double offset)
    {
        const acmpc_config& g = *c.cfg;
        const double *xs = c.f(F_XS), *ys = c.f(F_YS), *psi = c.f(F_PSI), *kap = c.f(F_KAP);
        const double *dist = c.f(F_DIST), *wid = c.f(F_WID), *vel = c.f(F_VEL);
        const double eps = 1e-12;
        // t2s of the state (offset, 0, pi/2) on waypoint 0 (spatial_mpc.py:186-189, dynamics.py:23-40)
        double psi0 = psi[0];
        double x0[3];
        x0[0] = cos(psi0) * (0.0 - ys[0]) - sin(psi0) * (offset - xs[0]);
        x0[1] = np_mod((kPi / 2.0 - psi0) + kPi, 2.0 * kPi) - kPi;
        x0[2] = 0.0;
        const double margin = g.width / 2.0;
        const double kmax = tan(g.delta_max) / g.wheelbase;
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            for (int r = 0; r < 3; ++r) c.f(C_M + r)[k] = -1.0;
            double d = 0, ka = 0, v = 0;
            if (k < n) {
                d = dist[k], ka = kap[k], v = vel[k];
                c.f(C_A11)[k] = 1.0;
                c.f(C_A12)[k] = d;
                c.f(C_A21)[k] = -(ka * ka) * d;
                c.f(C_A22)[k] = 1.0;
                c.f(C_A31)[k] = -ka / (v * d + eps);
                c.f(C_A33)[k] = 1.0;
                c.f(C_B22)[k] = d;
                c.f(C_B31)[k] = -1.0 / (v * v * d + eps);
            } else {
                for (int e = C_A11; e <= C_B31; ++e) c.f(e)[k] = 0.0;
            }
            for (int j = 0; j < 5; ++j) {
                bool real = j < nvar(k);
                c.f(C_S + j)[k] = real ? 1.0 : 0.0;
                c.f(C_DI + j)[k] = 1.0;
                c.f(C_EBI + j)[k] = 1.0;
            }
            for (int r = 0; r < 3; ++r) c.f(C_EEI + r)[k] = 1.0;
            // cost
            const double* Qx = (k < n) ? g.step_cost : g.final_cost;
            for (int j = 0; j < 3; ++j) c.f(C_P + j)[k] = Qx[j];
            c.f(C_P + 3)[k] = (k < n) ? g.r_term[0] : 0.0;
            c.f(C_P + 4)[k] = (k < n) ? g.r_term[1] : 0.0;
            c.f(C_Q + 0)[k] = (k < n) ? -g.r_term[0] * v : 0.0;
            c.f(C_Q + 1)[k] = (k < n) ? -g.r_term[1] * ka : 0.0;
            // equality rhs of block k
            if (k == 0) {
                for (int r = 0; r < 3; ++r) c.f(C_BE + r)[k] = -x0[r];
            } else {
                double dp = dist[k - 1], kp = kap[k - 1], vp = vel[k - 1];
                double b31 = -1.0 / (vp * vp * dp + eps), f3 = 1.0 / (vp * dp + eps);
                c.f(C_BE + 0)[k] = 0.0 * vp + 0.0 * kp - 0.0;
                c.f(C_BE + 1)[k] = 0.0 * vp + dp * kp - 0.0;
                c.f(C_BE + 2)[k] = b31 * vp + 0.0 * kp - f3;
            }
            // bounds (solvers/control.py:47-70,121-149), clipped to +-OSQP_INFTY
            double lo, hi;
            if (k == 0) lo = hi = x0[0];
            else lo = (-wid[k - 1] / 2.0) + margin, hi = (wid[k - 1] / 2.0) - margin;
            c.f(C_LB + 0)[k] = lo, c.f(C_UB + 0)[k] = hi;
            c.f(C_LB + 1)[k] = -kInfty, c.f(C_UB + 1)[k] = kInfty;
            c.f(C_LB + 2)[k] = 0.01, c.f(C_UB + 2)[k] = kInfty;
            c.f(C_LB + 3)[k] = g.input_v_min - 0.1, c.f(C_UB + 3)[k] = g.input_v_max + 0.1;
            c.f(C_LB + 4)[k] = -kmax, c.f(C_UB + 4)[k] = kmax;
        }
        ACMPC_SYNC();
    }

    // OSQP scale_data (Ruiz, `scaling` passes) + bounds scaling + set_rho_vec classes
    ACMPC_MEM void scale()
    {
        const acmpc_config& g = *c.cfg;
        double *T = c.f(C_T);   // T+0..2: partial row norms of block k+1 published by stage k
                                // T+3..5: e of equality block k published by stage k
        cs = 1.0;
        const int nv_total = 5 * H - 2;
        for (int pass = 0; pass < g.scaling; ++pass) {
            // (1) stage k publishes the partial norms of rows of block k+1 coming from its columns
            for (int k = c.lane; k < H; k += ACMPC_LANES) {
                double a11 = fabs(c.f(C_A11)[k]), a12 = fabs(c.f(C_A12)[k]), a21 = fabs(c.f(C_A21)[k]);
                double a22 = fabs(c.f(C_A22)[k]), a31 = fabs(c.f(C_A31)[k]), a33 = fabs(c.f(C_A33)[k]);
                double b22 = fabs(c.f(C_B22)[k]), b31 = fabs(c.f(C_B31)[k]);
                (T + 0 * c.Hs)[k] = fmax(a11, a12);
                (T + 1 * c.Hs)[k] = fmax(fmax(a21, a22), b22);
                (T + 2 * c.Hs)[k] = fmax(fmax(a31, a33), b31);
            }
            ACMPC_SYNC();
            // (2) every stage finishes the norms of its own rows / columns; publishes e of block k
            //     (T+3..5) and keeps its column scalings in T+6..10
            for (int k = c.lane; k < H; k += ACMPC_LANES) {
                double m[3], s[5], p[5];
                for (int r = 0; r < 3; ++r) m[r] = fabs(c.f(C_M + r)[k]);
                for (int j = 0; j < 5; ++j) s[j] = fabs(c.f(C_S + j)[k]), p[j] = fabs(c.f(C_P + j)[k]);
                double a11 = fabs(c.f(C_A11)[k]), a12 = fabs(c.f(C_A12)[k]), a21 = fabs(c.f(C_A21)[k]);
                double a22 = fabs(c.f(C_A22)[k]), a31 = fabs(c.f(C_A31)[k]), a33 = fabs(c.f(C_A33)[k]);
                double b22 = fabs(c.f(C_B22)[k]), b31 = fabs(c.f(C_B31)[k]);
                double cn[5];
                cn[0] = fmax(fmax(fmax(m[0], a11), fmax(a21, a31)), fmax(s[0], p[0]));
                cn[1] = fmax(fmax(m[1], a12), fmax(a22, fmax(s[1], p[1])));
                cn[2] = fmax(fmax(m[2], a33), fmax(s[2], p[2]));
                cn[3] = fmax(b31, fmax(s[3], p[3]));
                cn[4] = fmax(b22, fmax(s[4], p[4]));
                for (int j = 0; j < 5; ++j) (T + (6 + j) * c.Hs)[k] = 1.0 / sqrt(limit_scaling(cn[j]));
                for (int r = 0; r < 3; ++r) {
                    double rn = m[r];
                    if (k >= 1) rn = fmax(rn, (T + r * c.Hs)[k - 1]);
                    (T + (3 + r) * c.Hs)[k] = 1.0 / sqrt(limit_scaling(rn));
                }
            }
            ACMPC_SYNC();
            // (3) apply: A <- E A D, P <- D P D, q <- D q ; accumulate D, E ; cost normalisation
            double psum = 0.0, qmax = 0.0;
            for (int k = c.lane; k < H; k += ACMPC_LANES) {
                double d[5], ee[3];
                for (int j = 0; j < 5; ++j) d[j] = (T + (6 + j) * c.Hs)[k];
                for (int r = 0; r < 3; ++r) ee[r] = (T + (3 + r) * c.Hs)[k];
                for (int r = 0; r < 3; ++r) {
                    c.f(C_M + r)[k] = (c.f(C_M + r)[k] * ee[r]) * d[r];
                    c.f(C_EEI + r)[k] *= ee[r];
                }
                if (k < n) {
                    double e0 = (T + 3 * c.Hs)[k + 1], e1 = (T + 4 * c.Hs)[k + 1], e2 = (T + 5 * c.Hs)[k + 1];
                    c.f(C_A11)[k] = (c.f(C_A11)[k] * e0) * d[0];
                    c.f(C_A12)[k] = (c.f(C_A12)[k] * e0) * d[1];
                    c.f(C_A21)[k] = (c.f(C_A21)[k] * e1) * d[0];
                    c.f(C_A22)[k] = (c.f(C_A22)[k] * e1) * d[1];
                    c.f(C_B22)[k] = (c.f(C_B22)[k] * e1) * d[4];
                    c.f(C_A31)[k] = (c.f(C_A31)[k] * e2) * d[0];
                    c.f(C_A33)[k] = (c.f(C_A33)[k] * e2) * d[2];
                    c.f(C_B31)[k] = (c.f(C_B31)[k] * e2) * d[3];
                }
                const int nv = nvar(k);
                for (int j = 0; j < 5; ++j) {
                    double sj = c.f(C_S + j)[k];
                    double eb = 1.0 / sqrt(limit_scaling(fabs(sj)));
                    c.f(C_S + j)[k] = (sj * eb) * d[j];
                    c.f(C_EBI + j)[k] *= eb;
                    c.f(C_DI + j)[k] *= d[j];
                    double pj = (c.f(C_P + j)[k] * d[j]) * d[j];
                    c.f(C_P + j)[k] = pj;
                    if (j < nv) psum += fabs(pj);
                }
                for (int j = 0; j < 2; ++j) {
                    double qj = c.f(C_Q + j)[k] * d[3 + j];
                    c.f(C_Q + j)[k] = qj;
                    qmax = fmax(qmax, fabs(qj));
                }
            }
            psum = warp_sum(psum);
            qmax = warp_max(qmax);
            double ct = fmax(psum / (double)nv_total, limit_scaling(qmax));
            ct = 1.0 / limit_scaling(ct);
            for (int k = c.lane; k < H; k += ACMPC_LANES) {
                for (int j = 0; j < 5; ++j) c.f(C_P + j)[k] *= ct;
                for (int j = 0; j < 2; ++j) c.f(C_Q + j)[k] *= ct;
            }
            cs *= ct;
            ACMPC_SYNC();
        }
        cinv = 1.0 / cs;
        double nqu = 0.0, nqs = 0.0;
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            int bits = 0;
            for (int j = 0; j < 5; ++j) {
                double e = c.f(C_EBI + j)[k];
                double lo = c.f(C_LB + j)[k] * e, hi = c.f(C_UB + j)[k] * e;
                c.f(C_LB + j)[k] = lo, c.f(C_UB + j)[k] = hi;
                bits |= row_class(lo, hi) << (2 * j);
                c.f(C_EBI + j)[k] = 1.0 / e;
                c.f(C_DI + j)[k] = 1.0 / c.f(C_DI + j)[k];
            }
            c.f(C_TYPE)[k] = (double)bits;
            for (int r = 0; r < 3; ++r) {
                double e = c.f(C_EEI + r)[k];
                c.f(C_BE + r)[k] *= e;
                c.f(C_EEI + r)[k] = 1.0 / e;
            }
            if (k < n)
                for (int j = 0; j < 2; ++j) {
                    double qj = c.f(C_Q + j)[k];
                    nqu = fmax(nqu, fabs(c.f(C_DI + 3 + j)[k] * qj));
                    nqs = fmax(nqs, fabs(qj));
                }
        }
        nq_unscaled = warp_max(nqu);
        nq_scaled = warp_max(nqs);
        ACMPC_SYNC();
    }

    // Reduced matrix, analytic elimination of the inputs, block LDL' over the states.
    ACMPC_MEM void factor()
    {
        const double sigma = c.cfg->sigma, re = R.rho_eq;
        double* T = c.f(C_T);   // T+0 : re*b31^2-type terms for the next stage (see below)
        // (1) per stage: K_uu^{-1}, and what stage k+1 needs from the elimination of u_k
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            if (k < n) {
                double b31 = c.f(C_B31)[k], b22 = c.f(C_B22)[k];
                double s3 = c.f(C_S + 3)[k], s4 = c.f(C_S + 4)[k];
                double kv = c.f(C_P + 3)[k] + sigma + R.of(cls_b(k, 3)) * s3 * s3 + re * b31 * b31;
                double kk = c.f(C_P + 4)[k] + sigma + R.of(cls_b(k, 4)) * s4 * s4 + re * b22 * b22;
                double iv = 1.0 / kv, ik = 1.0 / kk;
                c.f(C_IV)[k] = iv, c.f(C_IK)[k] = ik;
                // coupling of u_k with x_{k+1} is (re*b31*m'_2) on t and (re*b22*m'_1) on e_psi;
                // publish iv*(re*b31)^2 and ik*(re*b22)^2 so stage k+1 can finish S_{k+1,k+1}
                (T + 0 * c.Hs)[k] = iv * (re * b31) * (re * b31);
                (T + 1 * c.Hs)[k] = ik * (re * b22) * (re * b22);
            } else {
                c.f(C_IV)[k] = 0.0, c.f(C_IK)[k] = 0.0;
                (T + 0 * c.Hs)[k] = 0.0, (T + 1 * c.Hs)[k] = 0.0;
            }
        }
        ACMPC_SYNC();
        // (2) per stage: S_kk (into C_SI, raw) and S_{k+1,k} (into C_G of stage k+1, raw)
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            double m0 = c.f(C_M + 0)[k], m1 = c.f(C_M + 1)[k], m2 = c.f(C_M + 2)[k];
            double s0 = c.f(C_S + 0)[k], s1 = c.f(C_S + 1)[k], s2 = c.f(C_S + 2)[k];
            double k00 = c.f(C_P + 0)[k] + sigma + R.of(cls_b(k, 0)) * s0 * s0 + re * m0 * m0;
            double k11 = c.f(C_P + 1)[k] + sigma + R.of(cls_b(k, 1)) * s1 * s1 + re * m1 * m1;
            double k22 = c.f(C_P + 2)[k] + sigma + R.of(cls_b(k, 2)) * s2 * s2 + re * m2 * m2;
            double k10 = 0.0, k20 = 0.0, k21 = 0.0;
            if (k >= 1) {   // elimination of u_{k-1}
                k11 -= (T + 1 * c.Hs)[k - 1] * m1 * m1;
                k22 -= (T + 0 * c.Hs)[k - 1] * m2 * m2;
            }
            if (k < n) {
                double a11 = c.f(C_A11)[k], a12 = c.f(C_A12)[k], a21 = c.f(C_A21)[k], a22 = c.f(C_A22)[k];
                double a31 = c.f(C_A31)[k], a33 = c.f(C_A33)[k], b22 = c.f(C_B22)[k], b31 = c.f(C_B31)[k];
                double iv = c.f(C_IV)[k], ik = c.f(C_IK)[k];
                k00 += re * (a11 * a11 + a21 * a21 + a31 * a31);
                k10 += re * (a11 * a12 + a21 * a22);
                k11 += re * (a12 * a12 + a22 * a22);
                k20 += re * (a31 * a33);
                k22 += re * (a33 * a33);
                // minus (1/K_uu) c c' with c_v = re*b31*(a31,0,a33), c_k = re*b22*(a21,a22,0)
                double gv = iv * (re * b31) * (re * b31), gk = ik * (re * b22) * (re * b22);
                k00 -= gv * a31 * a31 + gk * a21 * a21;
                k10 -= gk * a21 * a22;
                k11 -= gk * a22 * a22;
                k20 -= gv * a31 * a33;
                k22 -= gv * a33 * a33;
                // S_{k+1,k}: row r = m'_r * (re * row_r(x_k) - elimination), m' = m_{k+1}
                double n0 = c.f(C_M + 0)[k + 1], n1 = c.f(C_M + 1)[k + 1], n2 = c.f(C_M + 2)[k + 1];
                double* G = c.f(C_G);
                (G + 0 * c.Hs)[k + 1] = n0 * re * a11;
                (G + 1 * c.Hs)[k + 1] = n0 * re * a12;
                (G + 2 * c.Hs)[k + 1] = 0.0;
                (G + 3 * c.Hs)[k + 1] = n1 * (re * a21 - gk * a21);
                (G + 4 * c.Hs)[k + 1] = n1 * (re * a22 - gk * a22);
                (G + 5 * c.Hs)[k + 1] = 0.0;
                (G + 6 * c.Hs)[k + 1] = n2 * (re * a31 - gv * a31);
                (G + 7 * c.Hs)[k + 1] = 0.0;
                (G + 8 * c.Hs)[k + 1] = n2 * (re * a33 - gv * a33);
            }
            double* SI = c.f(C_SI);
            (SI + 0 * c.Hs)[k] = k00, (SI + 1 * c.Hs)[k] = k10, (SI + 2 * c.Hs)[k] = k11;
            (SI + 3 * c.Hs)[k] = k20, (SI + 4 * c.Hs)[k] = k21, (SI + 5 * c.Hs)[k] = k22;
        }
        ACMPC_SYNC();
        // (3) serial block LDL': Sigma_k = S_kk - G_k S_{k,k-1}', G_k = S_{k,k-1} Sigma_{k-1}^{-1}
        if (c.lane == 0) {
            double* SI = c.f(C_SI);
            double* G = c.f(C_G);
            double i00 = 0, i10 = 0, i11 = 0, i20 = 0, i21 = 0, i22 = 0;
            for (int k = 0; k < H; ++k) {
                double a00 = (SI + 0 * c.Hs)[k], a10 = (SI + 1 * c.Hs)[k], a11 = (SI + 2 * c.Hs)[k];
                double a20 = (SI + 3 * c.Hs)[k], a21 = (SI + 4 * c.Hs)[k], a22 = (SI + 5 * c.Hs)[k];
                if (k >= 1) {
                    double s[9], gg[9];
                    for (int e = 0; e < 9; ++e) s[e] = (G + e * c.Hs)[k];
                    // gg = s * inv  (inv symmetric)
                    for (int r = 0; r < 3; ++r) {
                        gg[3 * r + 0] = s[3 * r] * i00 + s[3 * r + 1] * i10 + s[3 * r + 2] * i20;
                        gg[3 * r + 1] = s[3 * r] * i10 + s[3 * r + 1] * i11 + s[3 * r + 2] * i21;
                        gg[3 * r + 2] = s[3 * r] * i20 + s[3 * r + 1] * i21 + s[3 * r + 2] * i22;
                    }
                    a00 -= gg[0] * s[0] + gg[1] * s[1] + gg[2] * s[2];
                    a10 -= gg[3] * s[0] + gg[4] * s[1] + gg[5] * s[2];
                    a11 -= gg[3] * s[3] + gg[4] * s[4] + gg[5] * s[5];
                    a20 -= gg[6] * s[0] + gg[7] * s[1] + gg[8] * s[2];
                    a21 -= gg[6] * s[3] + gg[7] * s[4] + gg[8] * s[5];
                    a22 -= gg[6] * s[6] + gg[7] * s[7] + gg[8] * s[8];
                    for (int e = 0; e < 9; ++e) (G + e * c.Hs)[k] = gg[e];
                }
                // inverse of the SPD 3x3 through LDL'
                double d0i = 1.0 / a00;
                double l10 = a10 * d0i, l20 = a20 * d0i;
                double d1i = 1.0 / (a11 - l10 * a10);
                double l21 = (a21 - l20 * a10) * d1i;
                double d2i = 1.0 / (a22 - l20 * a20 - l21 * (a21 - l20 * a10));
                double w20 = l10 * l21 - l20;   // (L^{-1})_{20}
                i22 = d2i;
                i21 = -l21 * d2i;
                i20 = w20 * d2i;
                i11 = d1i + l21 * l21 * d2i;
                i10 = -l10 * d1i - l21 * w20 * d2i;
                i00 = d0i + l10 * l10 * d1i + w20 * w20 * d2i;
                (SI + 0 * c.Hs)[k] = i00, (SI + 1 * c.Hs)[k] = i10, (SI + 2 * c.Hs)[k] = i11;
                (SI + 3 * c.Hs)[k] = i20, (SI + 4 * c.Hs)[k] = i21, (SI + 5 * c.Hs)[k] = i22;
            }
        }
        ACMPC_SYNC();
    }

    // One ADMM iteration's linear algebra: rhs -> x~ (in C_R), given w = rho z - y.
    // On exit C_R holds x~ (5 per stage).
    ACMPC_MEM void kkt_solve()
    {
        const double sigma = c.cfg->sigma, re = R.rho_eq;
        double* T = c.f(C_T);   // T+0..2: w of equality block k ; T+3,4: gv,gk published by stage k
        double* Rr = c.f(C_R);
        // (1) w_eq of own block
        for (int k = c.lane; k < H; k += ACMPC_LANES)
            for (int r = 0; r < 3; ++r)
                (T + r * c.Hs)[k] = re * c.f(C_ZE + r)[k] - c.f(C_YE + r)[k];
        ACMPC_SYNC();
        // (2) rhs of every variable, partial elimination terms
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            double w0 = (T + 0 * c.Hs)[k], w1 = (T + 1 * c.Hs)[k], w2 = (T + 2 * c.Hs)[k];
            double r[5];
            for (int j = 0; j < 5; ++j) {
                double wb = R.of(cls_b(k, j)) * c.f(C_ZB + j)[k] - c.f(C_YB + j)[k];
                r[j] = sigma * c.f(C_X + j)[k] + c.f(C_S + j)[k] * wb;
            }
            r[0] += c.f(C_M + 0)[k] * w0;
            r[1] += c.f(C_M + 1)[k] * w1;
            r[2] += c.f(C_M + 2)[k] * w2;
            double gv = 0.0, gk = 0.0;
            if (k < n) {
                double n0 = (T + 0 * c.Hs)[k + 1], n1 = (T + 1 * c.Hs)[k + 1], n2 = (T + 2 * c.Hs)[k + 1];
                double a11 = c.f(C_A11)[k], a12 = c.f(C_A12)[k], a21 = c.f(C_A21)[k], a22 = c.f(C_A22)[k];
                double a31 = c.f(C_A31)[k], a33 = c.f(C_A33)[k], b22 = c.f(C_B22)[k], b31 = c.f(C_B31)[k];
                r[0] += a11 * n0 + a21 * n1 + a31 * n2;
                r[1] += a12 * n0 + a22 * n1;
                r[2] += a33 * n2;
                r[3] += b31 * n2 - c.f(C_Q + 0)[k];
                r[4] += b22 * n1 - c.f(C_Q + 1)[k];
                // p = K_uu^{-1} r_u ; g = re*b*p
                double pv = c.f(C_IV)[k] * r[3], pk = c.f(C_IK)[k] * r[4];
                gv = re * b31 * pv, gk = re * b22 * pk;
                r[0] -= a31 * gv + a21 * gk;
                r[1] -= a22 * gk;
                r[2] -= a33 * gv;
                r[3] = pv, r[4] = pk;
            }
            for (int j = 0; j < 5; ++j) (Rr + j * c.Hs)[k] = r[j];
            (T + 3 * c.Hs)[k] = gv, (T + 4 * c.Hs)[k] = gk;
        }
        ACMPC_SYNC();
        for (int k = c.lane; k < H; k += ACMPC_LANES)
            if (k >= 1) {
                (Rr + 1 * c.Hs)[k] -= c.f(C_M + 1)[k] * (T + 4 * c.Hs)[k - 1];
                (Rr + 2 * c.Hs)[k] -= c.f(C_M + 2)[k] * (T + 3 * c.Hs)[k - 1];
            }
        ACMPC_SYNC();
        // (3) forward sweep  y_k = r_k - G_k y_{k-1}
        const double* G = c.f(C_G);
        if (c.lane == 0) {
            double y0 = Rr[0], y1 = (Rr + c.Hs)[0], y2 = (Rr + 2 * c.Hs)[0];
            for (int k = 1; k < H; ++k) {
                double t0 = (Rr + 0 * c.Hs)[k] - ((G + 0 * c.Hs)[k] * y0 + (G + 1 * c.Hs)[k] * y1 + (G + 2 * c.Hs)[k] * y2);
                double t1 = (Rr + 1 * c.Hs)[k] - ((G + 3 * c.Hs)[k] * y0 + (G + 4 * c.Hs)[k] * y1 + (G + 5 * c.Hs)[k] * y2);
                double t2 = (Rr + 2 * c.Hs)[k] - ((G + 6 * c.Hs)[k] * y0 + (G + 7 * c.Hs)[k] * y1 + (G + 8 * c.Hs)[k] * y2);
                y0 = t0, y1 = t1, y2 = t2;
                (Rr + 0 * c.Hs)[k] = y0, (Rr + 1 * c.Hs)[k] = y1, (Rr + 2 * c.Hs)[k] = y2;
            }
        }
        ACMPC_SYNC();
        // (4) w_k = Sigma_k^{-1} y_k
        const double* SI = c.f(C_SI);
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            double y0 = (Rr + 0 * c.Hs)[k], y1 = (Rr + 1 * c.Hs)[k], y2 = (Rr + 2 * c.Hs)[k];
            double i00 = (SI + 0 * c.Hs)[k], i10 = (SI + 1 * c.Hs)[k], i11 = (SI + 2 * c.Hs)[k];
            double i20 = (SI + 3 * c.Hs)[k], i21 = (SI + 4 * c.Hs)[k], i22 = (SI + 5 * c.Hs)[k];
            (Rr + 0 * c.Hs)[k] = i00 * y0 + i10 * y1 + i20 * y2;
            (Rr + 1 * c.Hs)[k] = i10 * y0 + i11 * y1 + i21 * y2;
            (Rr + 2 * c.Hs)[k] = i20 * y0 + i21 * y1 + i22 * y2;
        }
        ACMPC_SYNC();
        // (5) backward sweep  x_k = w_k - G_{k+1}' x_{k+1}
        if (c.lane == 0) {
            double x0 = (Rr + 0 * c.Hs)[H - 1], x1 = (Rr + 1 * c.Hs)[H - 1], x2 = (Rr + 2 * c.Hs)[H - 1];
            for (int k = H - 2; k >= 0; --k) {
                int q = k + 1;
                double t0 = (Rr + 0 * c.Hs)[k] - ((G + 0 * c.Hs)[q] * x0 + (G + 3 * c.Hs)[q] * x1 + (G + 6 * c.Hs)[q] * x2);
                double t1 = (Rr + 1 * c.Hs)[k] - ((G + 1 * c.Hs)[q] * x0 + (G + 4 * c.Hs)[q] * x1 + (G + 7 * c.Hs)[q] * x2);
                double t2 = (Rr + 2 * c.Hs)[k] - ((G + 2 * c.Hs)[q] * x0 + (G + 5 * c.Hs)[q] * x1 + (G + 8 * c.Hs)[q] * x2);
                x0 = t0, x1 = t1, x2 = t2;
                (Rr + 0 * c.Hs)[k] = x0, (Rr + 1 * c.Hs)[k] = x1, (Rr + 2 * c.Hs)[k] = x2;
            }
        }
        ACMPC_SYNC();
        // (6) recover the inputs  u_k = p_k - K_uu^{-1} (c' x_k + c'' x_{k+1})
        for (int k = c.lane; k < n; k += ACMPC_LANES) {
            double ey = (Rr + 0 * c.Hs)[k], ep = (Rr + 1 * c.Hs)[k], tt = (Rr + 2 * c.Hs)[k];
            double ep1 = (Rr + 1 * c.Hs)[k + 1], t1 = (Rr + 2 * c.Hs)[k + 1];
            double b31 = c.f(C_B31)[k], b22 = c.f(C_B22)[k];
            double sv = c.f(C_A31)[k] * ey + c.f(C_A33)[k] * tt + c.f(C_M + 2)[k + 1] * t1;
            double sk = c.f(C_A21)[k] * ey + c.f(C_A22)[k] * ep + c.f(C_M + 1)[k + 1] * ep1;
            (Rr + 3 * c.Hs)[k] -= c.f(C_IV)[k] * (re * b31) * sv;
            (Rr + 4 * c.Hs)[k] -= c.f(C_IK)[k] * (re * b22) * sk;
        }
        ACMPC_SYNC();
    }

    // (A v) on the rows of block k+1 contributed by stage k's variables -> T+0..2 of stage k
    ACMPC_MEM void publish_dyn_products(const double* V)
    {
        double* T = c.f(C_T);
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            double p0 = 0, p1 = 0, p2 = 0;
            if (k < n) {
                double ey = (V + 0 * c.Hs)[k], ep = (V + 1 * c.Hs)[k], tt = (V + 2 * c.Hs)[k];
                double vv = (V + 3 * c.Hs)[k], kc = (V + 4 * c.Hs)[k];
                p0 = c.f(C_A11)[k] * ey + c.f(C_A12)[k] * ep;
                p1 = c.f(C_A21)[k] * ey + c.f(C_A22)[k] * ep + c.f(C_B22)[k] * kc;
                p2 = c.f(C_A31)[k] * ey + c.f(C_A33)[k] * tt + c.f(C_B31)[k] * vv;
            }
            (T + 0 * c.Hs)[k] = p0, (T + 1 * c.Hs)[k] = p1, (T + 2 * c.Hs)[k] = p2;
        }
        ACMPC_SYNC();
    }

    ACMPC_MEM void compute_norms(Norms& N)
    {
        double* T = c.f(C_T);
        const double* X = c.f(C_X);
        publish_dyn_products(X);
        double v[12];
        for (int t = 0; t < 12; ++t) v[t] = 0.0;
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            double aty[5];
            const int nv = nvar(k);
            for (int j = 0; j < 5; ++j) aty[j] = c.f(C_S + j)[k] * c.f(C_YB + j)[k];
            // equality rows of block k
            for (int r = 0; r < 3; ++r) {
                double ax = c.f(C_M + r)[k] * (X + r * c.Hs)[k];
                if (k >= 1) ax += (T + r * c.Hs)[k - 1];
                double z = c.f(C_ZE + r)[k], ei = c.f(C_EEI + r)[k];
                double res = ax - z;
                v[0] = fmax(v[0], fabs(ei * res)), v[7] = fmax(v[7], fabs(res));
                v[1] = fmax(v[1], fabs(ei * z)), v[8] = fmax(v[8], fabs(z));
                v[2] = fmax(v[2], fabs(ei * ax)), v[9] = fmax(v[9], fabs(ax));
                aty[r] += c.f(C_M + r)[k] * c.f(C_YE + r)[k];
            }
            if (k < n) {
                double y0 = c.f(C_YE + 0)[k + 1], y1 = c.f(C_YE + 1)[k + 1], y2 = c.f(C_YE + 2)[k + 1];
                aty[0] += c.f(C_A11)[k] * y0 + c.f(C_A21)[k] * y1 + c.f(C_A31)[k] * y2;
                aty[1] += c.f(C_A12)[k] * y0 + c.f(C_A22)[k] * y1;
                aty[2] += c.f(C_A33)[k] * y2;
                aty[3] += c.f(C_B31)[k] * y2;
                aty[4] += c.f(C_B22)[k] * y1;
            }
            for (int j = 0; j < nv; ++j) {
                double x = (X + j * c.Hs)[k];
                double ax = c.f(C_S + j)[k] * x, z = c.f(C_ZB + j)[k], ei = c.f(C_EBI + j)[k];
                double res = ax - z;
                v[0] = fmax(v[0], fabs(ei * res)), v[7] = fmax(v[7], fabs(res));
                v[1] = fmax(v[1], fabs(ei * z)), v[8] = fmax(v[8], fabs(z));
                v[2] = fmax(v[2], fabs(ei * ax)), v[9] = fmax(v[9], fabs(ax));
                double px = c.f(C_P + j)[k] * x;
                double q = (j >= 3) ? c.f(C_Q + j - 3)[k] : 0.0;
                double dr = q + px + aty[j];
                double di = c.f(C_DI + j)[k];
                v[3] = fmax(v[3], fabs(di * dr)), v[6] = fmax(v[6], fabs(dr));
                v[4] = fmax(v[4], fabs(di * aty[j])), v[10] = fmax(v[10], fabs(aty[j]));
                v[5] = fmax(v[5], fabs(di * px)), v[11] = fmax(v[11], fabs(px));
            }
        }
        for (int t = 0; t < 12; ++t) v[t] = warp_max(v[t]);
        N.pri = v[0], N.nz = v[1], N.nAx = v[2];
        N.dua = cinv * v[3], N.nAty = v[4], N.nPx = v[5], N.nq = nq_unscaled;
        N.s_dua = v[6], N.s_pri = v[7], N.s_z = v[8], N.s_Ax = v[9], N.s_Aty = v[10], N.s_Px = v[11];
        N.s_q = nq_scaled;
        ACMPC_SYNC();
    }

    // scratch map at check iterations: T+0..2 exchange, T+3..7 delta_x, T+8..12: delta_y bound rows,
    // delta_y of the equality rows lives in C_R+0..2 (x~ is dead after the update step)
    ACMPC_MEM int primal_infeasible(double eps)
    {
        double* T = c.f(C_T);
        double* DYB = T + 8 * c.Hs;
        double* DYE = c.f(C_R);
        double nrm = 0.0, lhs = 0.0;
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            for (int r = 0; r < 3; ++r) {   // equality rows: finite bounds, no projection
                double dy = (DYE + r * c.Hs)[k], b = c.f(C_BE + r)[k];
                nrm = fmax(nrm, fabs(dy / c.f(C_EEI + r)[k]));
                lhs += b * fmax(dy, 0.0) + b * fmin(dy, 0.0);
            }
            for (int j = 0; j < nvar(k); ++j) {
                double dy = (DYB + j * c.Hs)[k], lo = c.f(C_LB + j)[k], hi = c.f(C_UB + j)[k];
                if (hi > kBig) dy = (lo < -kBig) ? 0.0 : fmin(dy, 0.0);
                else if (lo < -kBig) dy = fmax(dy, 0.0);
                (DYB + j * c.Hs)[k] = dy;
                nrm = fmax(nrm, fabs(dy / c.f(C_EBI + j)[k]));
                lhs += hi * fmax(dy, 0.0) + lo * fmin(dy, 0.0);
            }
        }
        nrm = warp_max(nrm);
        lhs = warp_sum(lhs);
        ACMPC_SYNC();
        if (!(nrm > eps) || !(lhs < -eps * nrm)) return 0;
        double m = 0.0;
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            double a[5];
            for (int j = 0; j < 5; ++j) a[j] = c.f(C_S + j)[k] * (DYB + j * c.Hs)[k];
            for (int r = 0; r < 3; ++r) a[r] += c.f(C_M + r)[k] * (DYE + r * c.Hs)[k];
            if (k < n) {
                double y0 = (DYE + 0 * c.Hs)[k + 1], y1 = (DYE + 1 * c.Hs)[k + 1], y2 = (DYE + 2 * c.Hs)[k + 1];
                a[0] += c.f(C_A11)[k] * y0 + c.f(C_A21)[k] * y1 + c.f(C_A31)[k] * y2;
                a[1] += c.f(C_A12)[k] * y0 + c.f(C_A22)[k] * y1;
                a[2] += c.f(C_A33)[k] * y2;
                a[3] += c.f(C_B31)[k] * y2;
                a[4] += c.f(C_B22)[k] * y1;
            }
            for (int j = 0; j < nvar(k); ++j) m = fmax(m, fabs(c.f(C_DI + j)[k] * a[j]));
        }
        m = warp_max(m);
        return m < eps * nrm;
    }

    ACMPC_MEM int dual_infeasible(double eps)
    {
        double* T = c.f(C_T);
        const double* DX = T + 3 * c.Hs;
        double nrm = 0.0, qdx = 0.0, pm = 0.0;
        for (int k = c.lane; k < H; k += ACMPC_LANES)
            for (int j = 0; j < nvar(k); ++j) {
                double dx = (DX + j * c.Hs)[k], di = c.f(C_DI + j)[k];
                nrm = fmax(nrm, fabs(dx / di));
                if (j >= 3) qdx += c.f(C_Q + j - 3)[k] * dx;
                pm = fmax(pm, fabs(di * (c.f(C_P + j)[k] * dx)));
            }
        nrm = warp_max(nrm), qdx = warp_sum(qdx), pm = warp_max(pm);
        if (!(nrm > eps) || !(qdx < -cs * eps * nrm) || !(pm < cs * eps * nrm)) return 0;
        publish_dyn_products(DX);
        int bad = 0;
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            for (int r = 0; r < 3; ++r) {   // equality rows: both bounds finite
                double a = c.f(C_M + r)[k] * (DX + r * c.Hs)[k];
                if (k >= 1) a += (T + r * c.Hs)[k - 1];
                a *= c.f(C_EEI + r)[k];
                if (a > eps * nrm || a < -eps * nrm) bad = 1;
            }
            for (int j = 0; j < nvar(k); ++j) {
                double a = c.f(C_EBI + j)[k] * (c.f(C_S + j)[k] * (DX + j * c.Hs)[k]);
                double lo = c.f(C_LB + j)[k], hi = c.f(C_UB + j)[k];
                if ((hi < kBig && a > eps * nrm) || (lo > -kBig && a < -eps * nrm)) bad = 1;
            }
        }
        int r = !warp_any(bad);
        ACMPC_SYNC();
        return r;
    }

    ACMPC_MEM int check(const Norms& N, int approximate)
    {
        const acmpc_config& g = *c.cfg;
        double k = approximate ? 10.0 : 1.0;
        if (N.pri > kInfty || N.dua > kInfty) return ACMPC_NON_CVX;
        double eps_p = k * g.eps_abs + k * g.eps_rel * fmax(N.nz, N.nAx);
        double eps_d = k * g.eps_abs + k * g.eps_rel * cinv * fmax(N.nq, fmax(N.nAty, N.nPx));
        int p_ok = N.pri < eps_p, d_ok = N.dua < eps_d;
        int p_inf = 0, d_inf = 0;
        if (!p_ok) p_inf = primal_infeasible(k * g.eps_prim_inf);
        if (!d_ok) d_inf = dual_infeasible(k * g.eps_dual_inf);
        if (p_ok && d_ok) return approximate ? ACMPC_SOLVED_INACCURATE : ACMPC_SOLVED;
        if (p_inf) return approximate ? ACMPC_PRIMAL_INFEASIBLE_INACCURATE : ACMPC_PRIMAL_INFEASIBLE;
        if (d_inf) return approximate ? ACMPC_DUAL_INFEASIBLE_INACCURATE : ACMPC_DUAL_INFEASIBLE;
        return 0;
    }

    ACMPC_MEM void solve(SolveInfo& info)
    {
        const acmpc_config& g = *c.cfg;
        const double alpha = g.alpha;
        double* T = c.f(C_T);
        double* Rr = c.f(C_R);
        R.set(clampd(g.rho, kRhoMin, kRhoMax));
        for (int k = c.lane; k < H; k += ACMPC_LANES) {
            for (int j = 0; j < 5; ++j) c.f(C_X + j)[k] = c.f(C_ZB + j)[k] = c.f(C_YB + j)[k] = 0.0;
            for (int r = 0; r < 3; ++r) c.f(C_YE + r)[k] = c.f(C_ZE + r)[k] = 0.0;
        }
        ACMPC_SYNC();
        factor();
        Norms N;
        int status = 0, iter = 0, updates = 0, checked = 0;
        for (iter = 1; iter <= g.max_iter; ++iter) {
            kkt_solve();
            publish_dyn_products(Rr);   // z~ contributions of stage k to block k+1
            checked = (g.check_termination > 0 && iter % g.check_termination == 0);
            const bool keep_delta = checked || iter == g.max_iter;
            const double re = R.rho_eq;
            for (int k = c.lane; k < H; k += ACMPC_LANES) {
                double xt[5];
                for (int j = 0; j < 5; ++j) xt[j] = (Rr + j * c.Hs)[k];
                // equality rows of block k: l == u == b.  x~ of this stage is in registers now, so
                // delta_y of these rows is parked in C_R+0..2 (nobody else reads C_R[k] here).
                for (int r = 0; r < 3; ++r) {
                    double zt = c.f(C_M + r)[k] * xt[r];
                    if (k >= 1) zt += (T + r * c.Hs)[k - 1];
                    double b = c.f(C_BE + r)[k];
                    double zh = alpha * zt + (1.0 - alpha) * c.f(C_ZE + r)[k];
                    double zn = clampd(zh + R.rinv_eq * c.f(C_YE + r)[k], b, b);
                    double dy = re * (zh - zn);
                    c.f(C_ZE + r)[k] = zn;
                    c.f(C_YE + r)[k] += dy;
                    if (keep_delta) (Rr + r * c.Hs)[k] = dy;
                }
                const int nv = nvar(k);
                for (int j = 0; j < nv; ++j) {
                    int cl = cls_b(k, j);
                    double zt = c.f(C_S + j)[k] * xt[j];
                    double zh = alpha * zt + (1.0 - alpha) * c.f(C_ZB + j)[k];
                    double zn = clampd(zh + R.inv_of(cl) * c.f(C_YB + j)[k], c.f(C_LB + j)[k], c.f(C_UB + j)[k]);
                    double dy = R.of(cl) * (zh - zn);
                    c.f(C_ZB + j)[k] = zn;
                    c.f(C_YB + j)[k] += dy;
                    double xo = c.f(C_X + j)[k];
                    double xn = alpha * xt[j] + (1.0 - alpha) * xo;
                    c.f(C_X + j)[k] = xn;
                    if (keep_delta) {
                        (T + (3 + j) * c.Hs)[k] = xn - xo;
                        (T + (8 + j) * c.Hs)[k] = dy;
                    }
                }
            }
            ACMPC_SYNC();
            if (checked) {
                compute_norms(N);
                status = check(N, 0);
                if (status) break;
            }
            if (g.adaptive_rho && g.adaptive_rho_interval > 0 && iter % g.adaptive_rho_interval == 0) {
                if (!checked) compute_norms(N);
                double rn = rho_estimate(N, R.rho);
                if (rn > R.rho * g.adaptive_rho_tolerance || rn < R.rho / g.adaptive_rho_tolerance) {
                    R.set(rn);
                    ++updates;
                    ACMPC_SYNC();
                    factor();
                }
            }
        }
        if (iter > g.max_iter) iter = g.max_iter;
        if (!checked) {
            compute_norms(N);
            status = check(N, 0);
        }
        if (!status) {
            status = check(N, 1);
            if (!status) status = ACMPC_MAX_ITER_REACHED;
        }
        info.status = status, info.iter = iter, info.rho_updates = updates;
        info.pri_res = N.pri, info.dua_res = N.dua;
        double obj = 0.0;
        for (int k = c.lane; k < H; k += ACMPC_LANES)
            for (int j = 0; j < nvar(k); ++j) {
                double x = c.f(C_X + j)[k];
                obj += 0.5 * c.f(C_P + j)[k] * x * x;
                if (j >= 3) obj += c.f(C_Q + j - 3)[k] * x;
            }
        info.obj_val = final_obj(status, warp_sum(obj) * cinv);
    }
};

// ------------------------------------------------------------------------------------------------
// one full MPC step for one instance.  `raw_path` = (H,3) staged in shared memory (it may alias the
// QP region: it is dead once build_waypoints() returns).
// ------------------------------------------------------------------------------------------------
struct InstanceOut {
    double *controls, *prediction, *cum_time, *states, *v_ref, *cost, *pri_res, *dua_res;
    int32_t *status, *status_speed, *iters, *rho_updates;
    double* waypoints;
};

ACMPC_DEV void solve_instance(const Ctx& c, const double* raw_path, double offset, double v_max_live,
                              int localised, const InstanceOut& o)
{
    const int n = c.n, H = c.H;
    build_waypoints(c, raw_path);
    SolveInfo si, ci;
    {
        SpeedQP sq(c);
        sq.assemble_and_scale(v_max_live, localised);
        sq.solve(si);
        // spatial_mpc.py:115-122: velocities are assigned only when the status is "solved"
        double* vel = c.f(F_VEL);
        const double* vs = c.f(S_R);
        for (int i = c.lane; i < n; i += ACMPC_LANES) vel[i] = (si.status == ACMPC_SOLVED) ? vs[i] : 0.0;
        ACMPC_SYNC();
    }
    ControlQP cq(c);
    cq.assemble(offset);
    cq.scale();
    cq.solve(ci);
    // unpack (spatial_mpc.py:193-212) and roll out (dynamics.py:42-63)
    const double L = c.cfg->wheelbase;
    const double *xs = c.f(F_XS), *ys = c.f(F_YS), *psi = c.f(F_PSI), *vel = c.f(F_VEL);
    for (int k = c.lane; k < H; k += ACMPC_LANES) {
        double ey = c.f(C_X + 0)[k] / c.f(C_DI + 0)[k];
        double ep = c.f(C_X + 1)[k] / c.f(C_DI + 1)[k];
        double tt = c.f(C_X + 2)[k] / c.f(C_DI + 2)[k];
        if (o.states) o.states[3 * k] = ey, o.states[3 * k + 1] = ep, o.states[3 * k + 2] = tt;
        if (k < n) {
            double v = c.f(C_X + 3)[k] / c.f(C_DI + 3)[k];
            double kc = c.f(C_X + 4)[k] / c.f(C_DI + 4)[k];
            if (o.controls) o.controls[k] = v, o.controls[n + k] = atan(kc * L);
            if (o.prediction) {
                o.prediction[2 * k] = xs[k] - ey * sin(psi[k]);
                o.prediction[2 * k + 1] = ys[k] + ey * cos(psi[k]);
            }
            if (o.cum_time) o.cum_time[k] = tt;
            if (o.v_ref) o.v_ref[k] = vel[k];
            if (o.waypoints)
                for (int f = 0; f < 7; ++f) o.waypoints[f * n + k] = c.f(F_XS + f)[k];
        }
    }
    if (c.lane == 0) {
        if (o.cost) *o.cost = ci.obj_val;
        if (o.pri_res) *o.pri_res = ci.pri_res;
        if (o.dua_res) *o.dua_res = ci.dua_res;
        if (o.status) *o.status = ci.status;
        if (o.status_speed) *o.status_speed = si.status;
        if (o.iters) o.iters[0] = si.iter, o.iters[1] = ci.iter;
        if (o.rho_updates) o.rho_updates[0] = si.rho_updates, o.rho_updates[1] = ci.rho_updates;
    }
}

}  // namespace acmpc
