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
