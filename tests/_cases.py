"""Shared iteration over the golden groups: the batch entry points take ONE is_localised flag per
launch, so a group is split into its unlocalised / localised instances."""
import numpy as np

import _golden


def golden_batches():
    """Yields (id, config kwargs, paths, offsets, vmax, is_localised, dict of golden outputs)."""
    G = _golden.load()
    for group, kw in _golden.groups():
        d = G[group]
        B = d["paths"].shape[0]
        vmax = d.get("vmax", np.full(B, float(kw["v_max"])))
        offs = d.get("offsets", np.zeros(B))
        loc = d.get("localised", np.zeros(B, dtype=int))
        for flag in (0, 1):
            m = loc == flag
            if not m.any():
                continue
            want = {k: d[k][m] for k in ("status", "status_speed", "iters", "rho_updates", "controls", "prediction",
                                         "cum_time", "v_ref", "cost", "dec_x", "waypoints", "pri_res", "dua_res",
                                         "times", "accelerations", "steer_rates")}
            yield f"{group}-loc{flag}", kw, d["paths"][m], offs[m], vmax[m], bool(flag), want


def assert_matches_golden(got, want, H, atol):
    """`got`: dict of output arrays of a batch solver; `want`: golden slices.  Integer fields bit-exact,
    floating point within `atol` (controls: the north-star bar is 1e-3; here far tighter)."""
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["status_speed"], want["status_speed"])
    assert np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["rho_updates"], want["rho_updates"])
    ok = want["status"] == 1
    B = want["status"].shape[0]
    np.testing.assert_allclose(got["states"].reshape(B, -1), want["dec_x"][:, : 3 * H], rtol=0, atol=atol)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=atol, atol=atol)
    for k in ("controls", "prediction", "cum_time", "v_ref", "waypoints"):
        np.testing.assert_allclose(got[k][ok], want[k][ok], rtol=0, atol=atol, err_msg=k)
    # spatial_mpc.py:208-211 by VALUE: times = diff(t), accelerations = diff(e_y) / times (sic), steer_rates =
    # diff(e_psi) / times.  Quotients of differences of quantities held to `atol`: the divisor is ~0.02-0.1 s
    # and the differences lose up to 3 digits, so the bar is 1e3 * atol relative-or-absolute (measured: ~1e-9)
    if "derived" in got:
        for row, k in enumerate(("times", "accelerations", "steer_rates")):
            np.testing.assert_allclose(got["derived"][ok][:, row], want[k][ok], rtol=1e3 * atol, atol=1e3 * atol, err_msg=k)
    np.testing.assert_allclose(got["pri_res"], want["pri_res"], rtol=1e-6, atol=atol)
    np.testing.assert_allclose(got["dua_res"], want["dua_res"], rtol=1e-6, atol=atol)
