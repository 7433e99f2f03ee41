"""CPU-only: the arithmetic of the CUDA kernel body (ac_mpc_b200/csrc/mpc_warp.cuh compiled by g++ with
32 lock-step lanes, tests/_emul) against the golden vectors of the reference Python.  This does not exercise the
32-lane execution -- tests/test_gpu_parity.py does, on the B200."""
import numpy as np
import pytest

import _cases
import _emul
from oracle import port

CASES = list(_cases.golden_batches())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_one_lane_body_matches_golden(case):
    _, kw, paths, offs, vmax, loc, want = case
    cfg = port.default_config(**kw)
    got = _emul.solve_batch(cfg, paths, offs, vmax, loc)
    _cases.assert_matches_golden(got, want, cfg.horizon, atol=1e-8)


def test_one_lane_body_matches_oracle_on_a_perturbed_batch():
    from ac_mpc_b200 import tracks

    paths, vmax = tracks.perturbed_batch("monza", 96, seed=11)
    cfg = port.default_config()
    a = _emul.solve_batch(cfg, paths, None, vmax, False)
    b = port.solve_batch(cfg, paths, None, vmax, False, nthreads=4)
    assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["rho_updates"], b["rho_updates"])
    assert np.abs(a["controls"] - b["controls"]).max() < 1e-8


def test_warm_sequence_matches_reference_object_reuse():
    """The reference test reuses ONE SpatialMPC for its 28 fixture paths, so its OSQP objects are warm-started
    (x, z, y and the adapted rho carry over, the data is re-equilibrated by update()): golden group fixture_warm."""
    import _golden

    G = _golden.load()
    g, paths = G["fixture_warm"], G["fixture_cold"]["paths"]
    cfg = port.default_config(**_golden.FIXTURE_CONFIG)
    warm = _emul.warm_buffer(cfg, 1)
    for b in range(paths.shape[0]):
        o = _emul.solve_batch(cfg, paths[b:b + 1], None, None, False, warm=warm)
        assert o["iters"][0].tolist() == g["iters"][b].tolist(), b
        assert o["status"][0] == g["status"][b] and o["status_speed"][0] == g["status_speed"][b]
        assert o["rho_updates"][0].tolist() == g["rho_updates"][b].tolist()
        if g["status"][b] == 1:
            np.testing.assert_allclose(o["controls"][0], g["controls"][b], rtol=0, atol=1e-8)
            np.testing.assert_allclose(o["cost"][0], g["cost"][b], rtol=1e-8, atol=1e-8)


def test_warm_batch_matches_oracle_sequences_with_mixed_localisation():
    """Several instances advance along the track for a few steps, alternating is_localised (two separate
    speed-solver objects in the reference, spatial_mpc.py:43-58): each emulated record against its own
    oracle object."""
    from ac_mpc_b200 import tracks

    cfg = port.default_config()
    cl = tracks.synthetic_centreline("monza")
    B, steps = 6, 5
    start = np.random.default_rng(3).integers(0, cl.shape[0], B)
    objs = [port.PortMPC(port.default_config()) for _ in range(B)]
    warm = _emul.warm_buffer(cfg, B)
    for t in range(steps):
        paths = tracks.make_instances(cl, (start + 40 * t) % cl.shape[0], 50)
        loc = bool(t % 2)
        got = _emul.solve_batch(cfg, paths, None, None, loc, warm=warm)
        for b in range(B):
            want = objs[b].step(paths[b], 0.0, None, loc, warm=True)
            assert got["iters"][b].tolist() == want["iters"].tolist(), (t, b)
            assert got["status"][b] == want["status"]
            np.testing.assert_allclose(got["controls"][b], want["controls"], rtol=0, atol=1e-8)


SETTINGS = [dict(), dict(scaling=0), dict(scaling=3), dict(check_termination=1), dict(check_termination=10),
            dict(check_termination=7, adaptive_rho_interval=20), dict(adaptive_rho=0), dict(adaptive_rho_interval=25),
            dict(alpha=1.0), dict(rho=1.0), dict(eps_abs=1e-5, eps_rel=1e-5), dict(max_iter=60), dict(max_iter=1),
            dict(check_termination=0, max_iter=80),
            # OSQP 1.x termination semantics (duality-gap test on top of the residual tests), alone and combined
            dict(check_dualgap=1), dict(check_dualgap=1, eps_abs=1e-5, eps_rel=1e-5), dict(check_dualgap=1, scaling=0),
            dict(check_dualgap=1, adaptive_rho_interval=25, check_termination=5)]


@pytest.mark.parametrize("kw", SETTINGS, ids=[",".join(f"{k}={v}" for k, v in s.items()) or "default" for s in SETTINGS])
def test_osqp_settings_sweep_matches_oracle(kw):
    """Every OSQP setting the config exposes, away from its default: check cadence (incl. every iteration and
    never), adaptive-rho cadence not a multiple of the check cadence, no equilibration, tight tolerances (the
    QP is infeasible as posed: primal-infeasibility certificates), iteration caps."""
    from ac_mpc_b200 import tracks

    paths, vmax = tracks.perturbed_batch("monza", 24, seed=7)
    cfg = port.default_config(**kw)
    a = _emul.solve_batch(cfg, paths, None, vmax, False)
    b = port.solve_batch(cfg, paths, None, vmax, False, nthreads=4)
    for k in ("status", "status_speed", "iters", "rho_updates"):
        assert np.array_equal(a[k], b[k]), k
    ok = (b["status"] == 1) & (b["status_speed"] == 1)
    if ok.any():
        assert np.abs(a["controls"] - b["controls"])[ok].max() < 1e-8
    np.testing.assert_allclose(a["cost"], b["cost"], rtol=1e-7, atol=1e-7)


def test_track_narrower_than_the_car_is_rejected_like_osqp_does():
    """width/2 < vehicle margin makes l > u on every e_y row: osqp.setup / update refuse such data (the
    reference would see a ValueError); oracle and kernel report status 'unsolved' (-10) and iterate nothing.
    A width that leaves less than 1e-4 m of room turns those rows into OSQP's equality class instead."""
    from ac_mpc_b200 import tracks

    paths, vmax = tracks.perturbed_batch("monza", 8, seed=7)
    cfg = port.default_config()
    narrow = paths.copy()
    narrow[:, :, 2] = 1.9
    a = _emul.solve_batch(cfg, narrow, None, vmax, False)
    b = port.solve_batch(cfg, narrow, None, vmax, False, nthreads=2)
    assert np.all(a["status"] == -10) and np.all(b["status"] == -10)
    assert np.all(a["iters"][:, 1] == 0) and np.array_equal(a["iters"], b["iters"])
    tight = paths.copy()
    tight[:, :, 2] = 1.99 + 5e-5
    a = _emul.solve_batch(cfg, tight, None, vmax, False)
    b = port.solve_batch(cfg, tight, None, vmax, False, nthreads=2)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iters"], b["iters"])
    assert np.abs(a["controls"] - b["controls"]).max() < 1e-8


@pytest.mark.parametrize("H", [33, 64, 65, 80, 96, 97, 128])
def test_horizon_classes_match_oracle(H):
    """Every stages-per-lane class and its boundaries: C = 2 (33..64), C = 3 (65..96: the split layout -- cold fields outside
    shared memory, chunks H2 + NS in shared memory, 32-lane scan matrices), C = 4 (97..128)."""
    import _golden
    from ac_mpc_b200 import tracks

    kw = _golden.racing_kwargs("spa", H)
    paths, vmax = tracks.perturbed_batch("spa", 6, horizon=H, seed=H)
    cfg = port.default_config(**kw)
    a = _emul.solve_batch(cfg, paths, None, vmax, False)
    b = port.solve_batch(cfg, paths, None, vmax, False, nthreads=4)
    for k in ("status", "status_speed", "iters", "rho_updates"):
        assert np.array_equal(a[k], b[k]), k
    ok = b["status"] == 1
    for k in ("controls", "states", "v_ref", "prediction", "derived"):
        np.testing.assert_allclose(a[k][ok], b[k][ok], rtol=0, atol=1e-8, err_msg=k)
    np.testing.assert_allclose(a["cost"], b["cost"], rtol=1e-7, atol=1e-7)
