"""Independent dense numpy statement of the OSQP iteration (Stellato et al. 2020, Alg. 1 + sec. 5).

Written separately from oracle/osqp_port.c (dense LU on the quasi-definite KKT matrix instead of a
sparse LDL', vectorised numpy instead of loops) so that the two can pin each other.  Test helper.
"""
import numpy as np

INF = 1e30
MIN_SC, MAX_SC = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6


def _limit(v):
    v = np.where(v < MIN_SC, 1.0, v)
    return np.minimum(v, MAX_SC)


def ruiz(P, q, A, l, u, passes=10):
    n, m = P.shape[0], A.shape[0]
    P, q, A = P.copy(), q.copy(), A.copy()
    D, E, c = np.ones(n), np.ones(m), 1.0
    for _ in range(passes):
        dcol = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0.0)
        erow = np.abs(A).max(axis=1) if m else np.zeros(0)
        d = 1.0 / np.sqrt(_limit(dcol))
        e = 1.0 / np.sqrt(_limit(erow))
        P = d[:, None] * P * d[None, :]
        A = e[:, None] * A * d[None, :]
        q = d * q
        D, E = D * d, E * e
        cm = np.abs(P).max(axis=0).mean()
        nq = _limit(np.array([np.abs(q).max()]))[0]
        ct = 1.0 / _limit(np.array([max(cm, nq)]))[0]
        P, q, c = P * ct, q * ct, c * ct
    return P, q, A, E * l, E * u, D, E, c


def solve(P, q, A, l, u, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
          eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000, check=25, scaling=10,
          adaptive_rho=True, adaptive_rho_interval=50, adaptive_rho_tolerance=5.0,
          x0=None, y0=None):
    P = np.asarray(P, float)
    A = np.asarray(A, float)
    n, m = P.shape[0], A.shape[0]
    l = np.maximum(np.asarray(l, float), -INF)
    u = np.minimum(np.asarray(u, float), INF)
    Ps, qs, As, ls, us, D, E, c = ruiz(P, np.asarray(q, float), A, l, u, scaling)

    loose = (ls < -INF * MIN_SC) & (us > INF * MIN_SC)
    eq = ~loose & (us - ls < 1e-4)

    def rho_vector(r):
        return np.where(loose, RHO_MIN, np.where(eq, 1e3 * r, r))

    def kkt(rv):
        return np.block([[Ps + sigma * np.eye(n), As.T], [As, -np.diag(1.0 / rv)]])

    rv = rho_vector(rho)
    K = kkt(rv)
    x = np.zeros(n) if x0 is None else x0 / D
    y = np.zeros(m) if y0 is None else y0 / E * c
    z = As @ x if x0 is not None else np.zeros(m)
    status, updates = "unsolved", 0
    it = 0
    for it in range(1, max_iter + 1):
        xp, zp = x, z
        rhs = np.concatenate([sigma * xp - qs, zp - y / rv])
        sol = np.linalg.solve(K, rhs)
        xt = sol[:n]
        zt = zp + (sol[n:] - y) / rv
        x = alpha * xt + (1 - alpha) * xp
        zhat = alpha * zt + (1 - alpha) * zp
        z = np.clip(zhat + y / rv, ls, us)
        dy = rv * (zhat - z)
        y = y + dy
        dx = x - xp
        checked = it % check == 0
        if checked or (adaptive_rho and it % adaptive_rho_interval == 0):
            Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
            rp_vec, rd_vec = Ax - z, Px + qs + Aty
            pri = np.abs(rp_vec / E).max() if m else 0.0
            dua = np.abs(rd_vec / D).max() / c
        if checked:
            ep = eps_abs + eps_rel * max(np.abs(z / E).max(), np.abs(Ax / E).max())
            ed = eps_abs + eps_rel * max(np.abs(qs / D).max(), np.abs(Aty / D).max(), np.abs(Px / D).max()) / c
            p_ok, d_ok = pri < ep, dua < ed
            if p_ok and d_ok:
                status = "solved"
                break
            if not p_ok:
                dyp = dy.copy()
                up_inf, lo_inf = us > INF * MIN_SC, ls < -INF * MIN_SC
                dyp[up_inf & lo_inf] = 0.0
                only_up = up_inf & ~lo_inf
                dyp[only_up] = np.minimum(dyp[only_up], 0.0)
                only_lo = lo_inf & ~up_inf
                dyp[only_lo] = np.maximum(dyp[only_lo], 0.0)
                nrm = np.abs(E * dyp).max()
                if nrm > eps_prim_inf:
                    lhs = us @ np.maximum(dyp, 0) + ls @ np.minimum(dyp, 0)
                    if lhs < -eps_prim_inf * nrm and np.abs((As.T @ dyp) / D).max() < eps_prim_inf * nrm:
                        status = "primal infeasible"
                        break
            if not d_ok:
                nrm = np.abs(D * dx).max()
                if nrm > eps_dual_inf and qs @ dx < -c * eps_dual_inf * nrm:
                    if np.abs((Ps @ dx) / D).max() < c * eps_dual_inf * nrm:
                        Adx = (As @ dx) / E
                        bad = ((us < INF * MIN_SC) & (Adx > eps_dual_inf * nrm)) | \
                              ((ls > -INF * MIN_SC) & (Adx < -eps_dual_inf * nrm))
                        if not bad.any():
                            status = "dual infeasible"
                            break
        if adaptive_rho and it % adaptive_rho_interval == 0:
            pn = np.abs(rp_vec).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
            dn = np.abs(rd_vec).max() / (max(np.abs(qs).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
            rn = float(np.clip(rho * np.sqrt(pn / (dn + 1e-10)), RHO_MIN, RHO_MAX))
            if rn > rho * adaptive_rho_tolerance or rn < rho / adaptive_rho_tolerance:
                rho = rn
                rv = rho_vector(rho)
                K = kkt(rv)
                updates += 1
    else:
        status = "maximum iterations reached"
    return dict(x=D * x, y=E * y / c, status=status, iter=it, rho_updates=updates, rho=rho,
                xs=x, zs=z, ys=y, D=D, E=E, c=c)
