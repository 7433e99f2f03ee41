"""Pins the C restatement of OSQP (oracle/osqp_port.c) -- CPU only.

PARITY UNPINNED against the real library (absent, see oracle/osqp_port.h).  What can be pinned here:
  * iterate-level agreement with an independent dense numpy statement of the same paper algorithm;
  * convergence to the exact optimum reported by HiGHS (scipy's vendored QP solver);
  * the infeasibility certificates, update() and warm-start semantics.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import _dense_admm as da
from oracle import port


def random_qp(rng, n=15, m=24):
    M = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.3)
    P = M.T @ M * 0.5
    A = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.4)
    A[:n] += np.eye(n)
    q = rng.standard_normal(n) * 3
    xs = rng.standard_normal(n)
    l = A @ xs - rng.random(m) * 2
    u = A @ xs + rng.random(m) * 2
    l[:3] = u[:3] = (A @ xs)[:3]
    l[3:5], u[3:5] = -np.inf, np.inf
    u[5:8] = np.inf
    l[8:10] = -np.inf
    return P, q, A, l, u


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("eps", [1e-3, 1e-7])
def test_port_matches_dense_numpy_admm(seed, eps):
    P, q, A, l, u = random_qp(np.random.default_rng(seed))
    s = port.PortOSQP(sp.csc_matrix(P), q, sp.csc_matrix(A), l, u, eps_abs=eps, eps_rel=eps)
    x, y, info = s.solve()
    r = da.solve(P, q, A, l, u, eps_abs=eps, eps_rel=eps)
    assert port.STATUS_STRINGS[info.status] == r["status"] == "solved"
    assert info.iter == r["iter"] and info.rho_updates == r["rho_updates"]
    assert np.abs(x - r["x"]).max() < 1e-9
    assert np.abs(y - r["y"]).max() < 1e-9
    assert np.abs(s.vec("D") - r["D"]).max() < 1e-12 and abs(s.scalar("c") - r["c"]) < 1e-12


def _highs_qp(P, q, A, l, u):
    from scipy.optimize._highspy import _core as hs

    n, m = len(q), len(l)
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    inf = hs.kHighsInf
    lp = hs.HighsLp()
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = q
    lp.col_lower_ = np.full(n, -inf)
    lp.col_upper_ = np.full(n, inf)
    lp.row_lower_ = np.where(np.isinf(l), -inf, l)
    lp.row_upper_ = np.where(np.isinf(u), inf, u)
    Ac = sp.csc_matrix(A)
    lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
    lp.a_matrix_.start_, lp.a_matrix_.index_, lp.a_matrix_.value_ = Ac.indptr, Ac.indices, Ac.data
    h.passModel(lp)
    Pl = sp.csc_matrix(sp.tril(P))
    hess = hs.HighsHessian()
    hess.dim_ = n
    hess.format_ = hs.HessianFormat.kTriangular
    hess.start_, hess.index_, hess.value_ = Pl.indptr, Pl.indices, Pl.data
    h.passHessian(hess)
    h.run()
    assert h.getModelStatus() == hs.HighsModelStatus.kOptimal
    return np.array(h.getSolution().col_value)


@pytest.mark.parametrize("seed", range(3))
def test_port_converges_to_highs_optimum(seed):
    P, q, A, l, u = random_qp(np.random.default_rng(100 + seed))
    P = P + 1e-3 * np.eye(len(q))          # strictly convex -> unique optimum
    try:
        x_exact = _highs_qp(P, q, A, l, u)
    except Exception as e:                  # pragma: no cover - private scipy API moved
        pytest.skip(f"HiGHS QP interface unavailable: {e}")
    s = port.PortOSQP(sp.csc_matrix(P), q, sp.csc_matrix(A), l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=20000)
    x, _, info = s.solve()
    assert info.status == 1
    assert np.abs(x - x_exact).max() < 1e-5


def test_infeasibility_certificates():
    P, q = np.eye(2), np.ones(2)
    A = np.array([[1.0, 0], [1, 0], [0, 1]])
    l, u = np.array([1.0, -np.inf, -1]), np.array([np.inf, 0.0, 1])
    _, _, info = port.PortOSQP(sp.csc_matrix(P), q, sp.csc_matrix(A), l, u).solve()
    assert port.STATUS_STRINGS[info.status] == "primal infeasible"
    assert da.solve(P, q, A, l, u)["status"] == "primal infeasible"
    P, q = np.zeros((2, 2)), np.array([1.0, 0])
    A = np.eye(2)
    l, u = np.array([-np.inf, -1]), np.array([1.0, 1])
    _, _, info = port.PortOSQP(sp.csc_matrix(P), q, sp.csc_matrix(A), l, u).solve()
    assert port.STATUS_STRINGS[info.status] == "dual infeasible"


def test_update_equals_fresh_setup_and_warm_start_carries_over():
    rng = np.random.default_rng(7)
    P, q, A, l, u = random_qp(rng)
    A2 = A * (1.0 + 0.05 * rng.standard_normal(A.shape))
    q2 = q + 0.1
    Ac, A2c = sp.csc_matrix(A), sp.csc_matrix(A2)
    assert (Ac.indices == A2c.indices).all()
    s = port.PortOSQP(sp.csc_matrix(P), q, Ac, l, u)
    x1, _, i1 = s.solve()
    # cold start + update  ==  fresh setup
    s.cold_start()
    s.update(q=q2, l=l, u=u, Ax=A2c.data)
    xa, _, ia = s.solve()
    xb, _, ib = port.PortOSQP(sp.csc_matrix(P), q2, A2c, l, u).solve()
    assert ia.iter == ib.iter and np.array_equal(xa, xb)
    # warm start (OSQP default): the next solve starts from the previous iterates
    s.update(q=q2, l=l, u=u, Ax=A2c.data)
    xw, _, iw = s.solve()
    assert iw.status == 1 and iw.iter <= ia.iter
    assert np.abs(xw - xa).max() < 5e-2


def test_max_iter_status():
    P, q, A, l, u = random_qp(np.random.default_rng(3))
    _, _, info = port.PortOSQP(sp.csc_matrix(P), q, sp.csc_matrix(A), l, u, eps_abs=1e-12, eps_rel=1e-12,
                               max_iter=30).solve()
    assert info.iter == 30
    assert port.STATUS_STRINGS[info.status] in ("maximum iterations reached", "solved inaccurate")
