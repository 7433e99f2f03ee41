"""Cut A of the drop-in boundary called stand-alone: SpatialMPC.compute_speed_profile /
update_prediction (spatial_mpc.py:89-123,156-168) and SpatialBicycleModel.t2s / s2t / linearise (dynamics.py:23-103).

Golden group `cut_a` of tests/golden/mpc_golden.npz = the reference's OWN classes driven through a sequence that mixes
compute_speed_profile and get_control on ONE object (they share the two speed-solver OSQP objects), plus the model
methods on the resulting paths.  CPU: the oracle restatement and the warp emulation of the kernel body against it;
-m gpu: the CUDA entry points and the mirror classes through the same sequence."""
import numpy as np
import pytest

import _golden
from oracle import port

TOL = 1e-7


def _seq():
    g = _golden.load()["cut_a"]
    for i in range(g["kind"].shape[0]):
        ev = None if np.isnan(g["end_vel"][i]) else float(g["end_vel"][i])
        yield i, int(g["kind"][i]), int(g["path_index"][i]), bool(g["localised"][i]), ev, float(g["vmax"][i]), g


def test_oracle_matches_reference_cut_a_sequence():
    G = _golden.load()
    P = G["racing_monza"]["paths"]
    obj = port.PortMPC(port.default_config(**_golden.racing_kwargs("monza")))
    for i, kind, pi, loc, ev, vm, g in _seq():
        if kind == 1:
            o = obj.step(P[pi], 0.0, vm, loc, warm=True)
            assert o["iters"][0] == g["iters"][i] and o["status_speed"] == g["status"][i]
            continue
        way = g["way_in"][i].copy()
        r = obj.speed_profile(way, vm, loc, ev, warm=True)
        assert (r["status"], r["iters"], r["rho_updates"]) == (g["status"][i], g["iters"][i], g["rho_updates"][i]), i
        np.testing.assert_allclose(r["x"], g["x"][i], rtol=0, atol=1e-9)
        np.testing.assert_allclose(way, g["way_out"][i], rtol=0, atol=1e-9)
        np.testing.assert_allclose(port.t2s(way[:3, 0], g["t2s_state"][i]), g["t2s_out"][i], rtol=0, atol=1e-13)
        np.testing.assert_allclose(port.s2t(way, g["s2t_states"][i]), g["s2t_out"][i], rtol=0, atol=1e-13)
        np.testing.assert_allclose(port.update_prediction(way, g["s2t_states"][i]), g["pred_out"][i], rtol=0, atol=1e-13)
        f, A, B = port.linearise(way)
        assert np.array_equal(f, g["lin_f"][i]) and np.array_equal(A, g["lin_A"][i]) and np.array_equal(B, g["lin_B"][i])


def test_kernel_body_emulation_matches_reference_cut_a_sequence():
    """The warp emulation of speed_instance's stand-alone mode, sharing its warm-start record with full steps."""
    import _emul

    G = _golden.load()
    P = G["racing_monza"]["paths"]
    cfg = port.default_config(**_golden.racing_kwargs("monza"))
    warm = _emul.warm_buffer(cfg, 1)
    for i, kind, pi, loc, ev, vm, g in _seq():
        if kind == 1:
            o = _emul.solve_batch(cfg, P[pi:pi + 1], None, np.array([vm]), loc, warm=warm)
            assert o["iters"][0, 0] == g["iters"][i] and o["status_speed"][0] == g["status"][i]
            continue
        way = g["way_in"][i][None].copy()
        r = _emul.speed_profile(cfg, way, np.array([vm]), loc, ev, warm=warm)
        assert (r["status"][0], r["iters"][0], r["rho_updates"][0]) == (g["status"][i], g["iters"][i], g["rho_updates"][i]), i
        np.testing.assert_allclose(r["x"][0], g["x"][i], rtol=0, atol=1e-8)
        np.testing.assert_allclose(way[0], g["way_out"][i], rtol=0, atol=1e-8)


def test_failed_speed_profile_leaves_velocities_untouched_in_the_emulation():
    import _emul

    g = _golden.load()["cut_a"]
    cfg = port.default_config(**dict(_golden.racing_kwargs("monza"), max_iter=10))
    way = g["way_in"][:2].copy()
    r = _emul.speed_profile(cfg, way, None, False, 14.0)
    assert np.all(r["status"] != 1)
    assert np.array_equal(way, g["way_in"][:2])            # spatial_mpc.py:115-122
    obj = port.PortMPC(cfg)
    w0 = g["way_in"][0].copy()
    ro = obj.speed_profile(w0, None, False, 14.0)
    assert ro["status"] == r["status"][0] and np.array_equal(w0, g["way_in"][0])
    np.testing.assert_allclose(r["x"][0], ro["x"], rtol=0, atol=1e-8)


# ---- the CUDA path --------------------------------------------------------------------------------------------------
VEH = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                     "max_steering_angle": lambda self: 0.3})()


@pytest.mark.gpu
def test_mirror_objects_match_reference_cut_a_sequence():
    """build_mpc -> SpatialMPC / SpatialBicycleModel of THIS package through the reference's call sequence."""
    from ac_mpc_b200 import tracks
    from ac_mpc_b200.control import ReferencePath, build_mpc

    G = _golden.load()
    P = G["racing_monza"]["paths"]
    mpc = build_mpc(tracks.racing_config("monza", 50), VEH)
    for i, kind, pi, loc, ev, vm, g in _seq():
        mpc.speed_profile_constraints["v_max"] = vm
        if kind == 1:
            mpc.get_control(P[pi], loc, 0.0)
            assert mpc.last_info["iters"][0] == g["iters"][i] and mpc.last_info["status_speed"] == "solved"
            continue
        rp = mpc.construct_waypoints(P[pi])
        rp.velocities = np.full(49, 5.0 + pi)
        np.testing.assert_allclose(rp.as_array(), g["way_in"][i], rtol=0, atol=TOL)
        out = mpc.compute_speed_profile(rp, loc, end_vel=ev)
        assert out is rp
        assert (mpc.last_speed_info["iters"], mpc.last_speed_info["rho_updates"]) == (g["iters"][i], g["rho_updates"][i]), i
        np.testing.assert_allclose(rp.as_array(), g["way_out"][i], rtol=0, atol=TOL)
        np.testing.assert_allclose(mpc.speed_profile, g["x"][i], rtol=0, atol=TOL)
        way = ReferencePath(49, g["way_out"][i].copy())
        np.testing.assert_allclose(mpc.model.t2s(way.get_state(0), g["t2s_state"][i]), g["t2s_out"][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(mpc.model.s2t(way, g["s2t_states"][i]), g["s2t_out"][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(mpc.update_prediction(g["s2t_states"][i], way), g["pred_out"][i], rtol=0, atol=1e-12)
        f, A, B = mpc.model.linearise(way)
        # plain IEEE arithmetic without contraction: bit-exact against numpy
        assert np.array_equal(f, g["lin_f"][i]) and np.array_equal(A, g["lin_A"][i]) and np.array_equal(B, g["lin_B"][i])


@pytest.mark.gpu
def test_speed_profile_batch_entry_points_match_oracle():
    """acmpc_speed_profile_batch_{host,device} on 3000 ReferencePaths (the kernel's own waypoints output), both
    localisation modes, with and without an end velocity; a capped max_iter leaves the failed rows untouched."""
    import torch

    from ac_mpc_b200 import BatchedMPC, _capi, tracks

    kw = _golden.racing_kwargs("spa")
    B = 3000
    paths, vmax = tracks.perturbed_batch("spa", B, seed=17)
    mpc = BatchedMPC(_capi.default_config(**kw), device=0)
    way0 = mpc.solve_host(paths, None, vmax, fields=["waypoints"])["waypoints"]
    way0[:, 6, :] = -1.0
    sub = np.random.default_rng(0).choice(B, 48, replace=False)
    for loc, ev in ((False, 14.0), (False, None), (True, 7.0)):
        way = way0.copy()
        r = mpc.speed_profile_host(way, vmax, loc, ev)
        d_way = torch.from_numpy(way0).cuda()
        rd = mpc.speed_profile_device(d_way, torch.from_numpy(vmax).cuda(), loc, ev)
        torch.cuda.synchronize()
        assert np.array_equal(rd["x"].cpu().numpy(), r["x"]) and np.array_equal(d_way.cpu().numpy(), way)
        assert np.array_equal(rd["iters"][:, 0].cpu().numpy(), r["iters"])
        assert np.array_equal(way[:, :6], way0[:, :6])
        for b in sub:
            obj = port.PortMPC(port.default_config(**kw))
            w = way0[b].copy()
            want = obj.speed_profile(w, vmax[b], loc, ev, warm=False)
            assert (r["status"][b], r["iters"][b], r["rho_updates"][b]) == (want["status"], want["iters"], want["rho_updates"])
            np.testing.assert_allclose(r["x"][b], want["x"], rtol=0, atol=TOL)
            np.testing.assert_allclose(way[b], w, rtol=0, atol=TOL)
    capped = BatchedMPC(_capi.default_config(**dict(kw, max_iter=10)), device=0)
    way = way0[:64].copy()
    r = capped.speed_profile_host(way, vmax[:64])
    assert np.all(r["status"] != 1) and np.array_equal(way, way0[:64])


@pytest.mark.gpu
def test_model_entry_points_batched_match_oracle():
    from ac_mpc_b200 import BatchedMPC, _capi, tracks

    mpc = BatchedMPC(_capi.default_config(), device=0)
    paths, vmax = tracks.perturbed_batch("monza", 513, seed=5)
    way = mpc.solve_host(paths, None, vmax, fields=["waypoints"])["waypoints"]
    rng = np.random.default_rng(1)
    states = np.stack([rng.uniform(-2, 2, (513, 49)), rng.uniform(-0.3, 0.3, (513, 49)), rng.uniform(0, 5, (513, 49))], axis=2)
    st3 = np.column_stack([rng.uniform(-1, 1, 513), rng.uniform(-1, 1, 513), rng.uniform(-7, 7, 513)])
    got = mpc.t2s(np.ascontiguousarray(way[:, :3, 0]), st3)
    for b in range(0, 513, 37):
        np.testing.assert_allclose(got[b], port.t2s(way[b, :3, 0], st3[b]), rtol=0, atol=1e-12)
    assert np.all(got[:, 1] >= -np.pi) and np.all(got[:, 1] < np.pi) and np.all(got[:, 2] == 0.0)
    s2t = mpc.s2t(way, states)
    pred = mpc.s2t(way, states, prediction=True)
    f, A, Bm = mpc.linearise(way)
    for b in range(0, 513, 37):
        np.testing.assert_allclose(s2t[b], port.s2t(way[b], states[b]), rtol=0, atol=1e-12)
        np.testing.assert_allclose(pred[b], port.update_prediction(way[b], states[b]), rtol=0, atol=1e-12)
        fo, Ao, Bo = port.linearise(way[b])
        assert np.array_equal(f[b], fo) and np.array_equal(A[b], Ao) and np.array_equal(Bm[b], Bo)
