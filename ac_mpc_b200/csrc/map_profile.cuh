// map_profile.cuh -- the whole-track speed profile: ONE speed-profile QP with
// n ~ 10^4 .. 4*10^4 waypoints, solved once at start-up.
//
// Reference path being replaced (file:line under /root/reference/src/acmpc/):
//   control/controller.py:49-57      Controller.compute_track_speed_profile
//   control/spatial_mpc.py:125-154   construct_waypoints over the whole centre line
//   control/spatial_mpc.py:60-87     compute_map_speed_profile: a fresh SpeedProfileSolver with
//                                    control_horizon = len(path), max_iterations = 40000, a_min / ay_max of
//                                    the map_speed_profile_constraints block
//   control/solvers/speed_profile.py:26-86  the QP and its osqp setup / solve
//   agent.py:287-302, :137-143       savgol_filter(v, 21, 3) and the [-25, +75) window mean (reference_speeds)
//
// It is the same QP as SpeedQP in mpc_warp.cuh, four hundred times longer, so the layout is different:
//   * a cooperative grid, ONE STAGE PER THREAD, 512 threads per CTA (23 CTAs for Monza's 11.6 k waypoints, 82 for
//     the Nordschleife's 41.7 k); iterates, problem data and the scan coefficients of the thread's stage stay in
//     registers for the whole solve.
//   * the reduced KKT system (P + sigma I + A' rho A) x~ = r is tridiagonal; its LDL' is computed once per rho by
//     one thread of CTA 0 over shared-memory tiles (a continued fraction: serial by nature, 0.4-1.5 ms).  Every
//     solve is two affine scans y_s = r_s + N_s y_{s-1},  x_s = y_s/d_s + N_{s+1} x_{s+1}, run on three levels:
//     Kogge-Stone over the 32 lanes with precomputed prefix products (5 FMAs + 5 shuffles), an affine scan over
//     the CTA's 16 warp aggregates by warp 0, and a chain over the CTA aggregates through L2: every CTA publishes
//     ONE (alpha, beta) pair as flag-in-data words and folds the pairs of the CTAs before it as they arrive.
//     => two L2 round trips per ADMM iteration, no barrier and no other global traffic in the iteration.
//   * neighbour values ride on the scans: the term a_{s-1}(rho z - y)_{s-1} that stage s-1 contributes to r_s is
//     folded into the forward chain ("g form": the carry entering a warp is g = t_prev + N_first y_prev), and the
//     x~_{s+1} that row s needs IS the backward carry.
//   * termination checks (every 25 iterations), equilibration and the factorisation use generic grid-wide helpers
//     (halo exchange through global memory, deterministic two-level reductions), one barrier each; every CTA
//     reduces the same partials in the same order, so all control flow is grid-uniform.
// The ADMM arithmetic follows SpeedQP statement by statement (OSQP 0.6 semantics, see mpc_warp.cuh).
#pragma once

#include "mpc_warp.cuh"

namespace acmpc {
namespace mapqp {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCtas = 148;
constexpr int kRed = 20;   // values of one grid reduction

struct MapParams {
    const double* track;   // [M,3] (x, y, width), or nullptr: kappas / distances are read from `waypoints`
    int M, n;              // n = M - 1 waypoints
    int do_solve;          // 0: construct_waypoints only
    acmpc_config cfg;      // speed_profile_constraints (a_min, ay_max, v_max already replaced) + OSQP settings
    double* waypoints;     // [7,n] ReferencePath rows
    double* solution;      // [n] dec.x (unscaled iterate at termination) or nullptr
    double* info;          // [8] status iter rho_updates pri_res dua_res obj_val rho -
    unsigned* counter;     // grid barrier arrivals (zero at launch)
    ulonglong4* slots;     // [2][ctas][ctas] inboxes of the chain: [chain parity][consumer position][producer position],
                           // one CTA aggregate as four flag-in-data words each
    double* halo;          // [2][2][kMaxCtas * kWarps]
    double* red;           // [2][kMaxCtas][kRed]
    double* kd;            // [grid * kThreads] K diagonal -> 1/pivot
    double* ko;            // [grid * kThreads] K off-diagonal -> N
};

struct Shared {
    double fa[kWarps + 1], fb[kWarps + 1], ba[kWarps + 1], bb[kWarps + 1];
    double gcar[kWarps], ccar[kWarps];
    double redw[kWarps][kRed];
    double redres[kRed];
    double tile_d[kThreads], tile_o[kThreads];
};

AC_DEV double ld_cg(const double* p) { return __ldcg(p); }
AC_DEV double2 ld_cg2(const double2* p) { return __ldcg(p); }

// Grid barrier on a monotonically increasing arrival counter (ctr[0]); ctr[1] is an abort flag read by the same
// 64-bit load.  A waiter that spins for seconds (a legitimate wait is at most the few milliseconds of the serial
// factorisation) raises it and every barrier falls through from then on, so a scheduling accident ends in an
// error code instead of a hung device.
AC_DEV void arrive_and_wait(unsigned* ctr, unsigned target)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned long long v;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        if ((unsigned)v >= target || (v >> 32) != 0ull) break;
        if (++spins > (1u << 22)) {
            *(volatile unsigned*)(ctr + 1) = 1u;
            break;
        }
    }
}

struct Grid {
    const MapParams& p;
    Shared& sh;
    int lane, warp, cta, ncta, s, n;
    unsigned epoch;    // grid barriers passed (arrival counter target)
    unsigned chains;   // chain steps taken (tag of the aggregates; odd = forward, even = backward)
    int hbuf, rbuf;

    AC_MEM Grid(const MapParams& pp, Shared& ss)
        : p(pp), sh(ss), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), cta(blockIdx.x), ncta(gridDim.x),
          s(blockIdx.x * kThreads + threadIdx.x), n(pp.n), epoch(0), chains(0), hbuf(0), rbuf(0)
    {
    }

    AC_MEM void sync()
    {
        __syncthreads();
        ++epoch;
        if (threadIdx.x == 0) arrive_and_wait(p.counter, epoch * (unsigned)ncta);
        __syncthreads();
    }

    // values of the neighbouring stages: to_next is what stage s+1 receives from s, to_prev what s-1 receives
    AC_MEM void exchange(double to_next, double to_prev, double& from_prev, double& from_next)
    {
        const int nw = ncta * kWarps, gw = cta * kWarps + warp;
        double* H = p.halo + (size_t)hbuf * 2 * kMaxCtas * kWarps;
        hbuf ^= 1;
        if (lane == 31) H[gw] = to_next;
        if (lane == 0) H[kMaxCtas * kWarps + gw] = to_prev;
        sync();
        double fp = __shfl_up_sync(kFull, to_next, 1), fn = __shfl_down_sync(kFull, to_prev, 1);
        if (lane == 0) fp = gw > 0 ? ld_cg(H + gw - 1) : 0.0;
        if (lane == 31) fn = gw + 1 < nw ? ld_cg(H + kMaxCtas * kWarps + gw + 1) : 0.0;
        from_prev = fp, from_next = fn;
    }

    // v[0..NM) -> max over the grid, v[NM..NM+NS) -> sum over the grid (same bits in every CTA)
    template <int NM, int NS>
    AC_MEM void reduce(double (&v)[NM + NS])
    {
        static_assert(NM + NS <= kRed, "reduction too wide");
        AC_UNROLL
        for (int k = 0; k < NM + NS; ++k) {
            double a = v[k];
            AC_UNROLL
            for (int o = 16; o > 0; o >>= 1) {
                const double b = __shfl_xor_sync(kFull, a, o);
                a = k < NM ? (a > b ? a : b) : a + b;
            }
            if (lane == 0) sh.redw[warp][k] = a;
        }
        __syncthreads();
        double* R = p.red + (size_t)rbuf * kMaxCtas * kRed;
        rbuf ^= 1;
        const int t = threadIdx.x;
        if (t < NM + NS) {
            double a = sh.redw[0][t];
            for (int w = 1; w < kWarps; ++w) {
                const double b = sh.redw[w][t];
                a = t < NM ? (a > b ? a : b) : a + b;
            }
            R[cta * kRed + t] = a;
        }
        sync();
        if (t < NM + NS) {
            double a = ld_cg(R + t);
            for (int c = 1; c < ncta; ++c) {
                const double b = ld_cg(R + c * kRed + t);
                a = t < NM ? (a > b ? a : b) : a + b;
            }
            sh.redres[t] = a;
        }
        __syncthreads();
        AC_UNROLL
        for (int k = 0; k < NM + NS; ++k) v[k] = sh.redres[k];
        __syncthreads();
    }

    // Warp 0 only, between two __syncthreads: affine scan over the CTA's 16 (alpha, beta) pairs (slot 0 = the
    // identity = the carry entering the CTA), publish the CTA aggregate at chain position `pos`, fold the aggregates
    // of positions < pos as they arrive, and leave the carry entering each warp in carry[0..16).
    // There is no barrier: an aggregate travels as four 8-byte words {32 data bits | 32-bit tag = chain step}
    // (8-byte accesses are single-copy atomic, so a word whose tag matches carries valid data -- NCCL's LL
    // protocol), and a CTA waits only for the CTAs before it in the chain.  Two buffers by chain parity are enough
    // because forward and backward chains alternate: a CTA can publish step E+2 only after it has folded step
    // E+1, i.e. after every consumer of its step-E aggregate has published E+1, which it does after reading E.
    AC_MEM void chain(const double* a17, const double* b17, double* carry, int pos)
    {
        const int l = lane;
        const bool in = l >= 1 && l <= kWarps;
        double A = in ? a17[l] : 1.0, B = in ? b17[l] : 0.0;
        AC_UNROLL
        for (int d = 1; d < 32; d <<= 1) {
            const double Ap = __shfl_up_sync(kFull, A, d), Bp = __shfl_up_sync(kFull, B, d);
            if (l >= d) B = fma(A, Bp, B), A = A * Ap;
        }
        // push: the aggregate goes into a private inbox slot of EVERY consumer (the CTAs after this one in the
        // chain), so each inbox line has exactly one polling CTA.  (With one shared slot per producer, polled by
        // all its consumers, an update took 1-3 us to reach some of the pollers; a line with a single poller
        // sees it in 0.25-0.45 us: tools/flag_latency_probe.cu, tools/chain_probe.cu.)
        const unsigned long long tag = (unsigned long long)chains << 32;
        ulonglong4* inbox = p.slots + (size_t)(chains & 1u) * ncta * ncta;   // [consumer position][producer position]
        const int per = (ncta + 31) >> 5;
        {
            const unsigned long long a = (unsigned long long)__double_as_longlong(__shfl_sync(kFull, A, kWarps)),
                                     b = (unsigned long long)__double_as_longlong(__shfl_sync(kFull, B, kWarps));
            for (int k = 0; k < per; ++k) {
                const int cpos = pos + 1 + l + 32 * k;
                if (cpos < ncta) {
                    unsigned long long* w = (unsigned long long*)(inbox + (size_t)cpos * ncta + pos);
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"((a & 0xffffffffull) | tag),
                                 "l"((a >> 32) | tag)
                                 : "memory");
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w + 2),
                                 "l"((b & 0xffffffffull) | tag), "l"((b >> 32) | tag)
                                 : "memory");
                }
            }
        }
        __syncwarp();
        // The poll loop is WARP-UNIFORM: every lane stays in it until all lanes have their aggregate (vote).  With a
        // per-lane exit the warp leaves the loop diverged, and every later __shfl_*_sync of the warp takes the
        // compiler's BRA.DIV slow path (a WARPSYNC.COLLECTIVE per SHFL): tools/chain_probe2.cu measured 5 k cycles
        // for the 22 shuffles of the fold below instead of 600 -- two thirds of the whole ADMM iteration.
        double FA = 1.0, FB = 0.0;
        for (int k = 0; k < per; ++k) {
            const int j = l * per + k;   // producer position
            bool ready = j >= pos;
            const unsigned long long* w = (const unsigned long long*)(inbox + (size_t)pos * ncta + (ready ? 0 : j));
            unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            unsigned spins = 0;
            do {
                if (!ready) {
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(w) : "memory");
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(w + 2) : "memory");
                    ready = (unsigned)(w0 >> 32) == chains && (unsigned)(w1 >> 32) == chains &&
                            (unsigned)(w2 >> 32) == chains && (unsigned)(w3 >> 32) == chains;
                }
                if ((++spins & 1023u) == 0u) {   // see arrive_and_wait
                    if (*(volatile unsigned*)(p.counter + 1) != 0u) ready = true;
                    if (spins > (1u << 22)) {
                        *(volatile unsigned*)(p.counter + 1) = 1u;
                        ready = true;
                    }
                }
            } while (!__all_sync(kFull, ready));
            if (j < pos) {
                const double va = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                const double vb = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
                FB = fma(va, FB, vb), FA = va * FA;
            }
        }
        AC_UNROLL
        for (int d = 1; d < 32; d <<= 1) {
            const double Ap = __shfl_up_sync(kFull, FA, d), Bp = __shfl_up_sync(kFull, FB, d);
            if (l >= d) FB = fma(FA, Bp, FB), FA = FA * Ap;
        }
        const double Gin = __shfl_sync(kFull, FB, 31);
        if (l < kWarps) carry[l] = fma(A, Gin, B);
    }
};

struct MapQP {
    Grid& g;
    const acmpc_config& cfg;
    double al, au, ss, p, q, la, ua, lb, ub, di, eai, ebi;
    int cls;
    double rho_a, rinv_a, rho_b, rinv_b;
    double x, za, zb, ya, yb, dx, dya, dyb;
    double dinv, M, phi[kLevels], psi[kLevels], Qw, Pb;
    double cs, cinv, nq_unscaled, nq_scaled;
    RhoSet R;
    // first-stage quantities of the infeasibility certificates, reduced together with the norms
    double c_nrp, c_lhs, c_nrd, c_qdx, c_pmx;

    AC_MEM MapQP(Grid& gg) : g(gg), cfg(gg.p.cfg) {}

    // speed_profile.py:26-59, then OSQP scale_data + set_rho_vec (SpeedQP::assemble_and_scale)
    AC_MEM void assemble_and_scale(double kap, double dist)
    {
        const int s = g.s, n = g.n;
        const bool ok = s < n, row = s <= n - 2;
        {
            const double ak = fabs(kap);
            double vdyn = sqrt(cfg.ay_max / (ak + 1e-12));
            vdyn = ak < cfg.ki_min ? cfg.v_max : vdyn;
            double v = vdyn < cfg.v_max ? vdyn : cfg.v_max;
            v = cfg.v_min > v ? cfg.v_min : v;
            double vb = v + 2.0;
            if (cfg.has_end_velocity && s == n - 1) vb = cfg.end_velocity;
            const double h = 1.0 / (2.0 * dist);
            al = row ? -h : 0.0, au = row ? h : 0.0;
            ss = ok ? 1.0 : 0.0, p = ok ? 1.0 : 0.0;
            q = ok ? -1.0 * vb : 0.0;
            di = eai = ebi = 1.0;
            la = row ? cfg.a_min : 0.0, ua = row ? cfg.a_max : 0.0;
            lb = ok ? cfg.v_min : 0.0, ub = ok ? vb : 0.0;
        }
        cs = 1.0;
        for (int pass = 0; pass < cfg.scaling; ++pass) {
            double aup, dn, unused;
            g.exchange(fabs(au), 0.0, aup, unused);
            const double m = vmax(vmax(fabs(p), fabs(ss)), vmax(fabs(al), aup));
            const double d = inv_sqrt(limit_scaling(m));
            g.exchange(0.0, d, unused, dn);
            const double ea = inv_sqrt(limit_scaling(vmax(fabs(al), fabs(au))));
            al = (al * ea) * d;
            au = (au * ea) * dn;
            eai = eai * ea;
            const double eb = inv_sqrt(limit_scaling(fabs(ss)));
            ss = (ss * eb) * d;
            ebi = ebi * eb;
            p = (p * d) * d;
            q = q * d;
            di = di * d;
            double v[2] = {fabs(q), fabs(p)};
            g.reduce<1, 1>(v);
            double ct = fmax(v[1] / (double)n, limit_scaling_u(v[0]));
            ct = 1.0 / limit_scaling_u(ct);
            p = p * ct, q = q * ct;
            cs *= ct;
        }
        cinv = 1.0 / cs;
        la = la * eai, ua = ua * eai;
        lb = lb * ebi, ub = ub * ebi;
        cls = row_class(la, ua) | (row_class(lb, ub) << 2);
        di = 1.0 / di;
        eai = 1.0 / eai;
        ebi = 1.0 / ebi;
        double v[2] = {fabs(di * q), fabs(q)};
        g.reduce<2, 0>(v);
        nq_unscaled = v[0], nq_scaled = v[1];
    }

    // K = P + sigma + A' rho A -> LDL' (serial, CTA 0) -> scan coefficients of this thread's stage
    AC_MEM void factor()
    {
        const int s = g.s, n = g.n, lane = g.lane;
        const double sigma = cfg.sigma;
        rho_a = R.of(cls, 0), rinv_a = R.inv_of(cls, 0);
        rho_b = R.of(cls, 2), rinv_b = R.inv_of(cls, 2);
        double tp, unused;
        g.exchange(rho_a * au * au, 0.0, tp, unused);
        g.p.kd[s] = p + sigma + rho_b * ss * ss + rho_a * al * al + tp;
        g.p.ko[s] = rho_a * al * au;
        g.sync();
        if (g.cta == 0) {
            double oprev = 0.0, dprev = 0.0;
            for (int base = 0; base < n; base += kThreads) {
                const int cnt = n - base < kThreads ? n - base : kThreads;
                if ((int)threadIdx.x < cnt) {
                    g.sh.tile_d[threadIdx.x] = ld_cg(g.p.kd + base + threadIdx.x);
                    g.sh.tile_o[threadIdx.x] = ld_cg(g.p.ko + base + threadIdx.x);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
#pragma unroll 4
                    for (int i = 0; i < cnt; ++i) {
                        const double dd = g.sh.tile_d[i], oo = g.sh.tile_o[i];
                        const double lw = oprev * dprev;   // L_{s,s-1}
                        const double piv = dd - lw * oprev;
                        dprev = urecip(piv);
                        oprev = oo;
                        g.sh.tile_d[i] = dprev;
                        g.sh.tile_o[i] = -lw;
                    }
                }
                __syncthreads();
                if ((int)threadIdx.x < cnt) {
                    g.p.kd[base + threadIdx.x] = g.sh.tile_d[threadIdx.x];
                    g.p.ko[base + threadIdx.x] = g.sh.tile_o[threadIdx.x];
                }
                // (tile_d / tile_o are rewritten only after the next __syncthreads of the loop)
                __syncthreads();
            }
        }
        g.sync();
        dinv = s < n ? ld_cg(g.p.kd + s) : 0.0;
        const double N = s < n ? ld_cg(g.p.ko + s) : 0.0;
        M = s + 1 < n ? ld_cg(g.p.ko + s + 1) : 0.0;
        // forward Kogge-Stone coefficients (zero carry into the warp) and Qw = prod_{first+1..s} N
        {
            double f = lane == 0 ? 0.0 : N, qq = lane == 0 ? 1.0 : N;
            AC_UNROLL
            for (int L = 0; L < kLevels; ++L) {
                const int d = 1 << L;
                phi[L] = f;
                const double fu = __shfl_up_sync(kFull, f, d), qu = __shfl_up_sync(kFull, qq, d);
                f = lane >= d ? f * fu : 0.0;
                qq = lane >= d ? qq * qu : qq;
            }
            Qw = qq;
        }
        // backward coefficients and Pb = prod_{s..last} M
        {
            double f = lane == 31 ? 0.0 : M, pp = M;
            AC_UNROLL
            for (int L = 0; L < kLevels; ++L) {
                const int d = 1 << L;
                psi[L] = f;
                const double fd = __shfl_down_sync(kFull, f, d), pd = __shfl_down_sync(kFull, pp, d);
                f = lane + d < 32 ? f * fd : 0.0;
                pp = lane + d < 32 ? pp * pd : pp;
            }
            Pb = pp;
        }
        if (lane == 31) g.sh.fa[g.warp + 1] = M * Qw;
        if (lane == 0) g.sh.ba[kWarps - g.warp] = Pb;
        __syncthreads();
    }

    // x~ = K^{-1} (r + shift(t)),  xn = x~ of the next stage
    AC_MEM void kkt_solve(double r, double t, double& xt, double& xn)
    {
        const int lane = g.lane, warp = g.warp;
        const double tp = __shfl_up_sync(kFull, t, 1);
        double Y = lane > 0 ? r + tp : r;
        AC_UNROLL
        for (int L = 0; L < kLevels; ++L) Y = fma(phi[L], __shfl_up_sync(kFull, Y, 1 << L), Y);
        if (lane == 31) g.sh.fb[warp + 1] = fma(M, Y, t);
        __syncthreads();
        ++g.chains;
        if (warp == 0) g.chain(g.sh.fa, g.sh.fb, g.sh.gcar, g.cta);
        __syncthreads();
        const double y = fma(Qw, g.sh.gcar[warp], Y);
        double Z = y * dinv;
        AC_UNROLL
        for (int L = 0; L < kLevels; ++L) Z = fma(psi[L], __shfl_down_sync(kFull, Z, 1 << L), Z);
        if (lane == 0) g.sh.bb[kWarps - warp] = Z;
        __syncthreads();
        ++g.chains;
        if (warp == 0) g.chain(g.sh.ba, g.sh.bb, g.sh.ccar, g.ncta - 1 - g.cta);
        __syncthreads();
        const double c = g.sh.ccar[kWarps - 1 - warp];
        xt = fma(Pb, c, Z);
        const double nx = __shfl_down_sync(kFull, xt, 1);
        xn = lane == 31 ? c : nx;
    }

    // one ADMM iteration (SpeedQP::iterate<true>)
    AC_MEM void iterate(const double alpha, const double sigma)
    {
        const double wa = rho_a * za - ya;
        const double t = au * wa;
        const double r = sigma * x - q + ss * (rho_b * zb - yb) + al * wa;
        double xt, xn;
        kkt_solve(r, t, xt, xn);
        {
            const double zt = al * xt + au * xn;
            const double zh = alpha * zt + (1.0 - alpha) * za;
            const double zn = vclamp(zh + rinv_a * ya, la, ua);
            const double dy = rho_a * (zh - zn);
            za = zn;
            ya = ya + dy;
            dya = dy;
        }
        {
            const double zt = ss * xt;
            const double zh = alpha * zt + (1.0 - alpha) * zb;
            const double zn = vclamp(zh + rinv_b * yb, lb, ub);
            const double dy = rho_b * (zh - zn);
            zb = zn;
            yb = yb + dy;
            dyb = dy;
        }
        const double xnew = alpha * xt + (1.0 - alpha) * x;
        dx = xnew - x;
        x = xnew;
    }

    // SpeedQP::compute_norms + the eps-independent halves of the two certificates, one exchange + one reduction
    AC_MEM void compute_norms(Norms& N)
    {
        double xn, tp;
        g.exchange(au * ya, x, tp, xn);
        double v[12];
        for (int k = 0; k < 12; ++k) v[k] = 0.0;
        const double aty = ss * yb + al * ya + tp;
        acc_row(v, al * x + au * xn, za, eai);
        acc_row(v, ss * x, zb, ebi);
        acc_col(v, q, p * x, aty, di);
        const double pa = project_dy(dya, la, ua), pb = project_dy(dyb, lb, ub);
        double w[17];
        for (int k = 0; k < 12; ++k) w[k] = v[k];
        w[12] = vmax(fabs(pa / eai), fabs(pb / ebi));
        w[13] = fabs(dx / di);
        w[14] = fabs(di * (p * dx));
        w[15] = support(pa, la, ua) + support(pb, lb, ub);
        w[16] = q * dx;
        g.reduce<15, 2>(w);
        N.pri = w[0], N.nz = w[1], N.nAx = w[2];
        N.dua = cinv * w[3], N.nAty = w[4], N.nPx = w[5], N.nq = nq_unscaled;
        N.s_dua = w[6], N.s_pri = w[7], N.s_z = w[8], N.s_Ax = w[9], N.s_Aty = w[10], N.s_Px = w[11];
        N.s_q = nq_scaled;
        c_nrp = w[12], c_nrd = w[13], c_pmx = w[14], c_lhs = w[15], c_qdx = w[16];
    }

    AC_MEM int primal_infeasible(double eps)
    {
        if (!(c_nrp > eps) || !(c_lhs < -eps * c_nrp)) return 0;
        const double pa = project_dy(dya, la, ua), pb = project_dy(dyb, lb, ub);
        double tp, unused;
        g.exchange(au * pa, 0.0, tp, unused);
        double v[1] = {fabs(di * (ss * pb + al * pa + tp))};
        g.reduce<1, 0>(v);
        return v[0] < eps * c_nrp;
    }

    AC_MEM int dual_infeasible(double eps)
    {
        if (!(c_nrd > eps) || !(c_qdx < -cs * eps * c_nrd) || !(c_pmx < cs * eps * c_nrd)) return 0;
        double dn, unused;
        g.exchange(0.0, dx, unused, dn);
        const double lim = eps * c_nrd;
        const double a = eai * (al * dx + au * dn), b = ebi * (ss * dx);
        const bool bad = ((ua < kBig) && (a > lim)) || ((la > -kBig) && (a < -lim)) || ((ub < kBig) && (b > lim)) ||
                         ((lb > -kBig) && (b < -lim));
        double v[1] = {bad ? 1.0 : 0.0};
        g.reduce<1, 0>(v);
        return v[0] == 0.0;
    }

    AC_MEM int check(const Norms& N, int approximate)
    {
        const double k = approximate ? 10.0 : 1.0;
        if (N.pri > kInfty || N.dua > kInfty) return ACMPC_NON_CVX;
        const double eps_p = k * cfg.eps_abs + k * cfg.eps_rel * fmax(N.nz, N.nAx);
        const double eps_d = k * cfg.eps_abs + k * cfg.eps_rel * cinv * fmax(N.nq, fmax(N.nAty, N.nPx));
        const int p_ok = N.pri < eps_p, d_ok = N.dua < eps_d;
        int p_inf = 0, d_inf = 0;
        if (!p_ok) p_inf = primal_infeasible(k * cfg.eps_prim_inf);
        if (!d_ok) d_inf = dual_infeasible(k * cfg.eps_dual_inf);
        if (p_ok && d_ok) return approximate ? ACMPC_SOLVED_INACCURATE : ACMPC_SOLVED;
        if (p_inf) return approximate ? ACMPC_PRIMAL_INFEASIBLE_INACCURATE : ACMPC_PRIMAL_INFEASIBLE;
        if (d_inf) return approximate ? ACMPC_DUAL_INFEASIBLE_INACCURATE : ACMPC_DUAL_INFEASIBLE;
        return 0;
    }

    // osqp_solve from a cold start (SpeedQP::solve)
    AC_MEM void solve(SolveInfo& info, double& vout)
    {
        const double alpha = cfg.alpha, sigma = cfg.sigma;
        R.set(clampu(cfg.rho, kRhoMin, kRhoMax));
        x = za = zb = ya = yb = 0.0;
        dx = dya = dyb = 0.0;
        factor();
        Norms N;
        int status = 0, iter = 0, updates = 0;
        int to_check = cfg.check_termination > 0 ? cfg.check_termination : -1;
        int to_adapt = (cfg.adaptive_rho && cfg.adaptive_rho_interval > 0) ? cfg.adaptive_rho_interval : -1;
        for (;;) {
            ++iter;
            const bool checked = (--to_check == 0);
            if (checked) to_check = cfg.check_termination;
            const bool adapt = (--to_adapt == 0);
            if (adapt) to_adapt = cfg.adaptive_rho_interval;
            const bool last = iter >= cfg.max_iter;
            iterate(alpha, sigma);
            if (checked || adapt || last) {
                compute_norms(N);
                if (checked) status = check(N, 0);
                if (status == 0 && adapt) {
                    const double rn = rho_estimate(N, R.rho);
                    if (rn > R.rho * cfg.adaptive_rho_tolerance || rn < R.rho / cfg.adaptive_rho_tolerance) {
                        R.set(rn);
                        ++updates;
                        factor();
                    }
                }
                if (status == 0 && last) {
                    if (!checked) status = check(N, 0);
                    if (status == 0) status = check(N, 1);
                    if (status == 0) status = ACMPC_MAX_ITER_REACHED;
                }
                if (status != 0) break;
            }
        }
        info.status = status, info.iter = iter, info.rho_updates = updates;
        info.pri_res = N.pri, info.dua_res = N.dua;
        double v[1] = {0.5 * p * x * x + q * x};
        g.reduce<0, 1>(v);
        info.obj_val = final_obj(status, v[0] * cinv);
        vout = x / di;
    }
};

__global__ void __launch_bounds__(kThreads, 1) acmpc_map_profile_kernel(const __grid_constant__ MapParams p)
{
    __shared__ Shared sh;
    Grid g(p, sh);
    const int s = g.s, n = p.n, M = p.M;
    const bool ok = s < n;
    // spatial_mpc.py:125-154: the "previous" point of waypoint 0 is the LAST point of the track
    double cx = 0, cy = 0, nw = 0, ps = 0, d = 0, kap = 0;
    if (ok && !p.track) {   // compute_map_speed_profile on a ReferencePath built earlier
        kap = p.waypoints[3 * n + s], d = p.waypoints[4 * n + s];
    } else if (ok) {
        const double* W = p.track;
        const int ic = 3 * s, in = ic + 3, ip = s == 0 ? 3 * (M - 1) : ic - 3;
        cx = W[ic], cy = W[ic + 1];
        const double nx = W[in], ny = W[in + 1];
        nw = W[in + 2];
        const double px = W[ip], py = W[ip + 1];
        const double ax = nx - cx, ay = ny - cy, bx = cx - px, by = cy - py;
        ps = atan2(ay, ax);
        d = sqrt(ax * ax + ay * ay);
        const double behind = atan2(by, bx);
        const double dang = np_mod(ps - behind + kPi, 2.0 * kPi) - kPi;
        kap = dang / (d + 1e-12) + 1e-12;
    }
    {   // kappa_0 := kappa_1
        const double k1 = __shfl_down_sync(kFull, kap, 1);
        if (s == 0 && p.track) kap = k1;
    }
    if (!p.do_solve) {
        if (ok) {
            double* o = p.waypoints;
            o[s] = cx, o[n + s] = cy, o[2 * n + s] = ps, o[3 * n + s] = kap, o[4 * n + s] = d, o[5 * n + s] = nw;
            o[6 * n + s] = 0.0;
        }
        return;
    }
    MapQP Q(g);
    Q.assemble_and_scale(kap, d);
    SolveInfo info;
    double v;
    Q.solve(info, v);
    if (ok) {
        double* o = p.waypoints;
        if (p.track)
            o[s] = cx, o[n + s] = cy, o[2 * n + s] = ps, o[3 * n + s] = kap, o[4 * n + s] = d, o[5 * n + s] = nw;
        // spatial_mpc.py:115-117: velocities assigned only when OSQP reports "solved" (else left as they were)
        if (info.status == ACMPC_SOLVED) o[6 * n + s] = v;
        else if (p.track) o[6 * n + s] = 0.0;
        if (p.solution) p.solution[s] = v;   // dec.x whatever the status
    }
    if (s == 0) {
        double* I = p.info;
        I[0] = (double)info.status, I[1] = (double)info.iter, I[2] = (double)info.rho_updates;
        I[3] = info.pri_res, I[4] = info.dua_res, I[5] = info.obj_val, I[6] = Q.R.rho;
        I[7] = (double)*(volatile unsigned*)(p.counter + 1);   // barrier abort flag
    }
}

// ---- agent.py:287-302 / :137-143 -------------------------------------------------------------------------------
// reference_speeds = savgol_filter(velocities, 21, 3)  (scipy default mode "interp": interior = 21-tap
// least-squares cubic smoother, the first / last 10 samples = the cubic fitted to the first / last 21 samples),
// then for every map index c the mean of reference_speeds over [c - 25, c + 75) with wrap-around.
constexpr int kSgWindow = 21, kSgHalf = 10;

struct SavgolCoeffs {
    double interior[kSgWindow];        // weights of v[i-10 .. i+10]
    double edge[kSgHalf][kSgWindow];   // row i: weights of v[0..21) for output i (mirrored for the tail)
};

__global__ void acmpc_savgol_kernel(const double* __restrict__ v, int n, const __grid_constant__ SavgolCoeffs c,
                                    double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0;
    if (i < kSgHalf) {
        for (int k = 0; k < kSgWindow; ++k) a = fma(c.edge[i][k], v[k], a);
    } else if (i >= n - kSgHalf) {
        const int r = n - 1 - i;
        for (int k = 0; k < kSgWindow; ++k) a = fma(c.edge[r][k], v[n - 1 - k], a);
    } else {
        for (int k = 0; k < kSgWindow; ++k) a = fma(c.interior[k], v[i - kSgHalf + k], a);
    }
    out[i] = a;
}

__global__ void acmpc_window_mean_kernel(const double* __restrict__ v, int n, int behind, int ahead,
                                         double* __restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double a = 0.0;
    int j = (c - behind) % n;
    if (j < 0) j += n;
    for (int k = 0; k < behind + ahead; ++k) {
        a += v[j];
        j = j + 1 == n ? 0 : j + 1;
    }
    out[c] = a / (double)(behind + ahead);
}

}   // namespace mapqp
}   // namespace acmpc
