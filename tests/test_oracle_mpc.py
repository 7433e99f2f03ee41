"""Pins the C restatement of the MPC step (oracle/acmpc_port.c) against golden vectors produced by
the reference's own Python (tests/golden/make_golden.py) -- CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

import _golden
from oracle import port

G = _golden.load()


@pytest.mark.parametrize("group,cfgkw", _golden.groups(), ids=[g for g, _ in _golden.groups()])
def test_port_step_matches_reference_python(group, cfgkw):
    g = G[group]
    cfg = port.default_config(**cfgkw)
    B = g["paths"].shape[0]
    vmax = g.get("vmax", np.full(B, cfg.v_max))
    offs = g.get("offsets", np.zeros(B))
    loc = g.get("localised", np.zeros(B, dtype=int))
    for b in range(B):
        o = port.PortMPC(cfg).step(g["paths"][b], offs[b], vmax[b], bool(loc[b]))
        assert o["status"] == g["status"][b] and o["status_speed"] == g["status_speed"][b]
        assert o["iters"].tolist() == g["iters"][b].tolist()
        assert o["rho_updates"].tolist() == g["rho_updates"][b].tolist()
        H = cfg.horizon
        np.testing.assert_allclose(o["states"].ravel(), g["dec_x"][b][: 3 * H], rtol=0, atol=1e-9)
        np.testing.assert_allclose(o["v_ref"], g["v_ref"][b], rtol=0, atol=1e-9)
        np.testing.assert_allclose(o["controls"], g["controls"][b], rtol=0, atol=1e-9)
        np.testing.assert_allclose(o["prediction"], g["prediction"][b], rtol=0, atol=1e-9)
        np.testing.assert_allclose(o["cum_time"], g["cum_time"][b], rtol=0, atol=1e-9)
        np.testing.assert_allclose(o["cost"], g["cost"][b], rtol=1e-9, atol=1e-9)


def test_port_warm_sequence_matches_reference_object_reuse():
    """test_spatial_mpc.py reuses ONE SpatialMPC for all 28 paths: OSQP objects are warm-started."""
    g, paths = G["fixture_warm"], G["fixture_cold"]["paths"]
    mpc = port.PortMPC(port.default_config(**_golden.FIXTURE_CONFIG))
    for b in range(paths.shape[0]):
        o = mpc.step(paths[b], 0.0, None, False, warm=True)
        assert o["iters"].tolist() == g["iters"][b].tolist(), b
        assert o["status"] == g["status"][b]
        np.testing.assert_allclose(o["controls"], g["controls"][b], rtol=0, atol=1e-8)


def test_port_assembly_matches_reference_qp_data():
    """The (P, q, A, l, u) the reference hands to osqp.setup, entry by entry."""
    cfg = port.default_config(**_golden.racing_kwargs("monza"))
    g = G["racing_monza"]
    mpc = port.PortMPC(cfg)
    mpc.step(g["paths"][0], 0.0, g["vmax"][0], False)
    for which in ("speed", "control"):
        ref, mine = G[f"qp_{which}"], mpc.qp(which)
        A_ref = sp.coo_matrix((ref["A_val"], (ref["A_row"], ref["A_col"])), shape=tuple(ref["shape"])).toarray()
        assert mine["A"].shape == A_ref.shape
        np.testing.assert_allclose(mine["A"].toarray(), A_ref, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(mine["Pdiag"], ref["P_diag"], rtol=0, atol=0)
        np.testing.assert_allclose(mine["q"], ref["q"], rtol=1e-12, atol=1e-15)
        for k in ("l", "u"):
            np.testing.assert_allclose(np.clip(mine[k], -1e30, 1e30), np.clip(ref[k], -1e30, 1e30), rtol=1e-12, atol=1e-15)
    # sizes quoted in SURVEY.md fact 5
    assert mpc.qp("control")["A"].shape == (398, 248) and mpc.qp("speed")["A"].shape == (97, 49)


def test_control_qp_is_infeasible_as_posed_but_osqp_says_solved():
    """t_0 = 0 (equality) against t_0 >= 0.01 (bound)."""
    cfg = port.default_config(**_golden.racing_kwargs("monza"))
    mpc = port.PortMPC(cfg)
    o = mpc.step(G["racing_monza"]["paths"][0], 0.0, G["racing_monza"]["vmax"][0], False)
    qp = mpc.qp("control")
    H = cfg.horizon
    assert qp["l"][2] == qp["u"][2] == 0.0 and qp["l"][3 * H + 2] == 0.01
    assert o["status"] == 1


def test_batch_threads_are_deterministic():
    from ac_mpc_b200 import tracks

    paths, vmax = tracks.perturbed_batch("monza", 64, seed=5)
    cfg = port.default_config()
    a = port.solve_batch(cfg, paths, None, vmax, nthreads=1)
    b = port.solve_batch(cfg, paths, None, vmax, nthreads=4)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
