"""Caller side of the MPC step: ControlProcess._reference_path / _update_shared_memory
(control/controller.py:257-280) and the command lookups of control/commands.py.

Golden vectors: tests/golden/publish_golden.npz, produced by the UNMODIFIED reference classes
(tests/golden/make_publish_golden.py) plus the vectors of the reference's own tests/test_commands.py:15-58.
Bar: bit-exact (these are selections, casts and two-term interpolations in the arrays' own precision)."""
import os

import numpy as np
import pytest

from ac_mpc_b200 import _capi
from oracle import port

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "publish_golden.npz")


def _group(g):
    with np.load(_PATH) as z:
        return {k.split("/")[1]: z[k] for k in z.files if k.startswith(g + "/")}


# ---- CPU: the oracle restatement against the reference's output --------------------------------------------
@pytest.mark.parametrize("name", ["f32", "f64"])
def test_oracle_lookups_match_reference(name):
    d = _group(f"lookup_{name}")
    for b in range(d["elapsed"].shape[0]):
        t = float(d["elapsed"][b])
        np.testing.assert_array_equal(port.select_command(d["cum_time"][b], d["commands"][b], t), d["selected"][b])
        np.testing.assert_array_equal(port.interpolate_command(d["cum_time"][b], d["commands"][b], t), d["interpolated"][b])


def test_oracle_matches_the_reference_test_vectors():
    d = _group("ref_test_interp")
    got = np.array([port.interpolate_command(d["cum_time"], d["commands"], float(t)) for t in d["elapsed"]])
    np.testing.assert_allclose(got, d["expected"], rtol=0, atol=5e-8)     # assertAlmostEqual, 7 places
    d = _group("reference_path")
    for c, p in zip(d["centrelines"], d["paths"]):
        np.testing.assert_array_equal(port.reference_path(c, 50), p)


def test_publish_entry_points_exported():
    L = _capi.load()
    assert L.acmpc_reference_paths_host(None, 1, 500, None, None) == 1
    assert L.acmpc_publish_host(None, 1, None, None, None, None, None, None) == 1
    assert L.acmpc_select_commands_f32_host(None, 1, 49, None, None, None, 0, None, None) == 1
    assert L.acmpc_select_commands_f64_host(None, 1, 49, None, None, None, 0, None, None) == 1


# ---- GPU --------------------------------------------------------------------------------------------------------
def _solver(**kw):
    from ac_mpc_b200 import BatchedMPC

    return BatchedMPC(_capi.default_config(**kw), device=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["f32", "f64"])
def test_lookups_match_reference_golden(name):
    d = _group(f"lookup_{name}")
    mpc = _solver()
    sel = mpc.select_commands(d["cum_time"], d["commands"], d["elapsed"])
    assert sel.dtype == d["selected"].dtype and mpc.launch_info()["launches"] == 1
    np.testing.assert_array_equal(sel, d["selected"])
    itp = mpc.select_commands(d["cum_time"], d["commands"], d["elapsed"], interpolate=True)
    np.testing.assert_array_equal(itp, d["interpolated"])
    # elapsed time before the first stamp: the selector returns the LAST command (Python's index -1), kept as is
    early = d["elapsed"] < 0
    assert early.any() and np.array_equal(sel[early], d["commands"][early, -1])


@pytest.mark.gpu
def test_reference_test_vectors():
    """tests/test_commands.py:15-58 of the reference."""
    mpc = _solver()
    d = _group("ref_test_index")
    B = len(d["elapsed"])
    ct = np.tile(d["cum_time"], (B, 1))
    _, idx = mpc.select_commands(ct, np.zeros((B, ct.shape[1], 2)), d["elapsed"], interpolate=True, return_indices=True)
    np.testing.assert_array_equal(idx[:, 0], d["expected_index"])
    np.testing.assert_allclose(ct[np.arange(B), idx[:, 0]] - d["elapsed"], d["expected_distance"], rtol=0, atol=5e-8)
    d = _group("ref_test_interp")
    B = len(d["elapsed"])
    got = mpc.select_commands(np.tile(d["cum_time"], (B, 1)), np.tile(d["commands"], (B, 1, 1)), d["elapsed"], interpolate=True)
    np.testing.assert_allclose(got, d["expected"], rtol=0, atol=5e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype", [(1, np.float32), (2, np.float64), (19, np.float32), (127, np.float64)])
def test_lookups_match_oracle_at_other_sizes(n, dtype):
    rng = np.random.default_rng(n)
    B = 300
    cum = np.cumsum(rng.uniform(0.0, 0.08, (B, n)), axis=1).astype(dtype)     # zero steps: repeated stamps (ties)
    cmd = rng.normal(0, 20, (B, n, 2)).astype(dtype)
    el = rng.uniform(-0.1, 1.1, B) * (cum[:, -1] + 0.1)
    mpc = _solver()
    for interp, fn in ((False, port.select_command), (True, port.interpolate_command)):
        got = mpc.select_commands(cum, cmd, el, interpolate=interp)
        with np.errstate(invalid="ignore", divide="ignore"):
            want = np.array([fn(cum[b], cmd[b], float(el[b])) for b in range(B)])
        np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_reference_path_and_publish_match_reference():
    d = _group("reference_path")
    mpc = _solver(horizon=50)
    np.testing.assert_array_equal(mpc.reference_paths(d["centrelines"]), d["paths"])
    # controller.py:260-266 only works when the stride yields exactly H rows (H = 40: 42 rows -> np.stack raises)
    with pytest.raises(RuntimeError, match="horizon"):
        _solver(horizon=40).reference_paths(d["centrelines"])
    for H in (20, 25, 100):
        want = np.array([port.reference_path(c, H) for c in d["centrelines"]])
        np.testing.assert_array_equal(_solver(horizon=H).reference_paths(d["centrelines"]), want)
    p = _group("publish")
    ci, ct, pl = mpc.publish(p["controls"], p["cum_time"], p["prediction"])
    np.testing.assert_array_equal(ci, p["control_inputs"])
    np.testing.assert_array_equal(ct, p["control_cumtime"])
    np.testing.assert_array_equal(pl, p["predicted_locations"])


@pytest.mark.gpu
def test_step_publish_lookup_chain():
    """The production sequence (controller.py:233-239, :110-112): perceived centre line -> reference path ->
    get_control -> float32 publish -> command selected 30 ms later, against the same chain on the oracle."""
    from ac_mpc_b200 import tracks
    import _golden

    kw = _golden.racing_kwargs("monza")
    mpc = _solver(**kw)
    cl = tracks.synthetic_centreline("monza")
    B = 64
    starts = np.random.default_rng(3).integers(0, len(cl) - 600, B)
    # 500 perceived points, 0.2 m apart, in the ego frame (x right, y forward)
    paths50 = tracks.make_instances(cl, starts, 500, None, None)[:, :, :2]
    centre = paths50.astype(np.float32)
    paths = mpc.reference_paths(centre)
    out = mpc.solve_host(paths, None, None, False)
    want = port.solve_batch(port.default_config(**kw), paths, None, None, False, nthreads=4)
    assert np.array_equal(out["status"], want["status"])
    ok = out["status"] == 1
    assert ok.mean() > 0.5
    ci, ct, pl = mpc.publish(out["controls"], out["cum_time"], out["prediction"])
    sel = mpc.select_commands(ct, ci, np.full(B, 0.03))
    for b in np.flatnonzero(ok):
        w_ci = want["controls"][b].T.astype(np.float32)
        w_ct = want["cum_time"][b].astype(np.float32)
        np.testing.assert_allclose(sel[b], port.select_command(w_ct, w_ci, 0.03), rtol=0, atol=1e-5)
