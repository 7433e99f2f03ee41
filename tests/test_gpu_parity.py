"""-m gpu: the CUDA path, called through the C ABI (ctypes -> libacmpc_b200.so), against
  (1) the golden vectors produced by the reference's own Python (tests/golden/make_golden.py),
  (2) the CPU oracle on seeded batches at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's full batch sizes.
Tolerance: the north star asks for controls within 1e-3 absolute (steer rad, speed m/s) of the
reference OSQP solve; these tests hold the kernel to 1e-7 (same algorithm, different summation order),
and to bit-exact statuses / iteration counts / rho updates."""
import numpy as np
import pytest

import _cases
from ac_mpc_b200 import BatchedMPC, _capi, tracks
from oracle import port

pytestmark = pytest.mark.gpu

TOL = 1e-7
CASES = list(_cases.golden_batches())


def _solver(**kw):
    return BatchedMPC(_capi.default_config(**kw), device=0)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_matches_reference_golden(case):
    _, kw, paths, offs, vmax, loc, want = case
    mpc = _solver(**kw)
    got = mpc.solve_host(paths, offs, vmax, loc)
    assert mpc.launch_info()["launches"] >= 1
    _cases.assert_matches_golden(got, want, kw["horizon"], atol=TOL)


@pytest.mark.parametrize("track", tracks.TRACK_ORDER)
def test_kernel_matches_oracle_on_perturbed_batches(track):
    import _golden

    kw = _golden.racing_kwargs(track)
    paths, vmax = tracks.perturbed_batch(track, 256, seed=21)
    offs = np.random.default_rng(5).uniform(-0.5, 0.5, 256)
    for loc in (False, True):
        got = _solver(**kw).solve_host(paths, offs, vmax, loc)
        want = port.solve_batch(port.default_config(**kw), paths, offs, vmax, loc, nthreads=8)
        for k in ("status", "status_speed", "iters", "rho_updates"):
            assert np.array_equal(got[k], want[k]), k
        for k in ("controls", "prediction", "cum_time", "states", "v_ref", "waypoints", "cost"):
            np.testing.assert_allclose(got[k], want[k], rtol=0, atol=TOL, err_msg=k)
        # the residuals the solver reports stay under the reference tolerances (eps_abs = eps_rel = 1e-3)
        solved = got["status"] == 1
        assert solved.mean() > 0.95
        np.testing.assert_allclose(got["pri_res"], want["pri_res"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(got["dua_res"], want["dua_res"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("H", [4, 5, 20, 33, 40, 64, 80, 100, 128])
def test_horizon_sweep_matches_oracle(H):
    import _golden

    kw = _golden.racing_kwargs("spa", H)
    paths, vmax = tracks.perturbed_batch("spa", 64, horizon=H, seed=H)
    got = _solver(**kw).solve_host(paths, None, vmax, False)
    want = port.solve_batch(port.default_config(**kw), paths, None, vmax, False, nthreads=8)
    for k in ("status", "status_speed", "iters", "rho_updates"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("controls", "states", "v_ref", "cost"):
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=TOL, err_msg=k)


def test_failed_speed_profile_and_tight_iteration_cap_match_oracle():
    """max_iter below the first termination check: OSQP's 'maximum iterations reached' /
    'solved inaccurate' branch; the speed profile is then NOT written (spatial_mpc.py:119-122)."""
    paths, vmax = tracks.perturbed_batch("monza", 32, seed=2)
    for mi in (7, 25, 30):
        got = _solver(max_iter=mi).solve_host(paths, None, vmax, False)
        want = port.solve_batch(port.default_config(max_iter=mi), paths, None, vmax, False, nthreads=4)
        for k in ("status", "status_speed", "iters"):
            assert np.array_equal(got[k], want[k]), (mi, k)
        # FP-order differences are amplified when v_ref = 0 makes the model singular: compare solved ones
        ok = (want["status"] == 1) & (want["status_speed"] == 1)
        np.testing.assert_allclose(got["controls"][ok], want["controls"][ok], rtol=0, atol=TOL)
        failed = want["status_speed"] != 1
        assert np.all(got["v_ref"][failed] == 0.0)


def test_empty_and_single_instance_batches():
    mpc = _solver()
    out = mpc.solve_host(np.zeros((0, 50, 3)))
    assert out["controls"].shape == (0, 2, 49)
    paths, vmax = tracks.perturbed_batch("monza", 1, seed=9)
    a = mpc.solve_host(paths, None, vmax)
    b = port.solve_batch(port.default_config(), paths, None, vmax)
    np.testing.assert_allclose(a["controls"], b["controls"], rtol=0, atol=TOL)
    # optional inputs: offsets NULL == zeros, vmax NULL == cfg.v_max
    c = mpc.solve_host(paths)
    d = mpc.solve_host(paths, np.zeros(1), np.full(1, 84.0))
    assert np.array_equal(c["controls"], d["controls"])


def test_device_entry_point_equals_host_entry_point_and_is_deterministic():
    import torch

    paths, vmax = tracks.perturbed_batch("monza", 777, seed=4)   # ragged: not a multiple of anything
    mpc = _solver()
    host = mpc.solve_host(paths, None, vmax)
    dp = torch.from_numpy(paths).cuda()
    dv = torch.from_numpy(vmax).cuda()
    packed, views = mpc.alloc_device_outputs(777)
    for _ in range(2):
        packed.zero_()
        mpc.solve_device(dp, None, dv, False, out=views)
        torch.cuda.synchronize()
        for k, v in views.items():
            assert np.array_equal(v.cpu().numpy(), host[k]), k     # bit-exact run to run
    # a subset of output fields may be requested (NULL pointers are skipped)
    some = mpc.solve_host(paths, None, vmax, fields=["controls", "status"])
    assert set(k for k in some if not k.startswith("_")) == {"controls", "status"}
    assert np.array_equal(some["controls"], host["controls"])


def test_full_size_batch_properties():
    """BASELINE configs[1] size (4096, Monza): properties that need no oracle run --
    permutation equivariance (instances are independent), agreement with the oracle on a random
    subsample, residuals under the reference tolerances, and bound satisfaction of the controls."""
    B = 4096
    paths, vmax = tracks.perturbed_batch("monza", B, seed=1)
    mpc = _solver()
    out = mpc.solve_host(paths, None, vmax)
    perm = np.random.default_rng(0).permutation(B)
    outp = mpc.solve_host(paths[perm], None, vmax[perm])
    for k in ("controls", "status", "iters", "cost", "states"):
        assert np.array_equal(outp[k], out[k][perm]), k
    sub = np.random.default_rng(1).choice(B, 256, replace=False)
    want = port.solve_batch(port.default_config(), paths[sub], None, vmax[sub], nthreads=8)
    assert np.array_equal(out["iters"][sub], want["iters"])
    np.testing.assert_allclose(out["controls"][sub], want["controls"], rtol=0, atol=TOL)
    solved = out["status"] == 1
    assert solved.mean() > 0.99
    cfg = mpc.cfg
    eps = 1e-3
    v, delta = out["controls"][solved, 0], out["controls"][solved, 1]
    # OSQP's iterate satisfies the bounds only up to its primal tolerance eps_abs + eps_rel*|.|
    assert v.min() > cfg.input_v_min - 0.1 - (eps + eps * 85) and v.max() < cfg.input_v_max + 0.1 + (eps + eps * 85)
    assert np.abs(delta).max() < cfg.delta_max + 2e-3
    assert np.all(np.isfinite(out["prediction"][solved]))


def test_drop_in_spatial_mpc_get_control_matches_reference_attributes():
    """The reference-facing object API (build_mpc -> SpatialMPC.get_control) on the golden fixtures:
    the attributes the caller reads (controller.py:274-280) equal the reference's."""
    import _golden
    from ac_mpc_b200.control import build_mpc

    G = _golden.load()
    veh = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                         "max_steering_angle": lambda self: 0.3})()
    g = G["racing_monza"]
    for b in range(g["paths"].shape[0]):
        mpc = build_mpc(tracks.racing_config("monza", 50), veh)
        mpc.speed_profile_constraints["v_max"] = float(g["vmax"][b])      # controller.py:241-243
        mpc.get_control(g["paths"][b], bool(g["localised"][b]), float(g["offsets"][b]))
        assert mpc.infeasibility_counter == g["infeasibility_delta"][b]
        if g["status"][b] == 1:
            np.testing.assert_allclose(mpc.projected_control, g["controls"][b], rtol=0, atol=TOL)
            np.testing.assert_allclose(mpc.current_prediction, g["prediction"][b], rtol=0, atol=TOL)
            np.testing.assert_allclose(mpc.cum_time, g["cum_time"][b], rtol=0, atol=TOL)
            np.testing.assert_allclose(mpc.reference_path.velocities, g["v_ref"][b], rtol=0, atol=TOL)
            # spatial_mpc.py:208-211 by value, against the reference's own attributes
            np.testing.assert_allclose(mpc.times, g["times"][b], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(mpc.accelerations, g["accelerations"][b], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(mpc.steer_rates, g["steer_rates"][b], rtol=1e-5, atol=1e-5)


def _assert_equals_oracle(got, want, fields=("controls", "states", "v_ref", "cost", "prediction", "cum_time")):
    for k in ("status", "status_speed", "iters", "rho_updates"):
        assert np.array_equal(got[k], want[k]), k
    # a failed speed profile (v_ref = 0) makes the model singular and amplifies summation-order noise
    ok = want["status_speed"] == 1
    for k in fields:
        np.testing.assert_allclose(got[k][ok], want[k][ok], rtol=0, atol=TOL, err_msg=k)


def test_work_queue_across_launches_of_one_handle():
    """The persistent warps of the control kernel draw instances from a ticket counter that is never reset:
    launches of very different sizes on ONE handle (fewer instances than warps, more than the device holds,
    ragged) must each solve every instance exactly once."""
    mpc = _solver()
    cfg = port.default_config()
    for B, seed in ((1, 3), (3000, 4), (7, 5), (5000, 6), (2, 7), (1185, 8)):
        paths, vmax = tracks.perturbed_batch("monza", B, seed=seed)
        got = mpc.solve_host(paths, None, vmax)
        want = port.solve_batch(cfg, paths, None, vmax, nthreads=8)
        _assert_equals_oracle(got, want, fields=("controls", "cost"))


def test_work_queue_of_the_phased_kernel_across_ragged_launches():
    """Horizons 65..96 run the split layout with CTA-phased rounds (four tickets per CTA round, partial rounds included):
    launches whose sizes are not multiples of four, smaller than one CTA, larger than the device holds, on ONE handle."""
    import _golden

    kw = _golden.racing_kwargs("spa", 80)
    mpc = _solver(**kw)
    cfg = port.default_config(**kw)
    for B, seed in ((1, 3), (3001, 4), (7, 5), (2, 7), (2375, 8), (5, 9)):
        paths, vmax = tracks.perturbed_batch("spa", B, horizon=80, seed=seed)
        got = mpc.solve_host(paths, None, vmax)
        want = port.solve_batch(cfg, paths, None, vmax, nthreads=16)
        _assert_equals_oracle(got, want, fields=("controls", "cost"))


def test_host_pipeline_overlaps_batches_without_mixing_them_up():
    """HostPipeline: several different batches in flight (H2D / kernels / D2H of consecutive batches overlap), pinned and
    plain inputs; every batch must come back bit-equal to the synchronous call on the same inputs."""
    import torch

    mpc = _solver()
    B = 3000
    fields = ["controls", "prediction", "cum_time", "status", "iters", "cost"]
    pipe = mpc.pipeline(B, fields, depth=2)
    batches = []
    for seed in range(5):
        p, v = tracks.perturbed_batch("monza", B, seed=40 + seed)
        o = np.random.default_rng(seed).uniform(-0.3, 0.3, B)
        if seed & 1:
            p, v = torch.from_numpy(p).pin_memory().numpy(), torch.from_numpy(v).pin_memory().numpy()
        batches.append((p, o, v, bool(seed == 3)))
    want = [_solver().solve_host(p, o, v, loc, fields=fields) for p, o, v, loc in batches]
    tickets, got = [], []
    for i, (p, o, v, loc) in enumerate(batches):
        tickets.append(pipe.submit(p, o, v, loc))
        if i >= 1:
            got.append({k: a.copy() for k, a in pipe.wait(tickets[i - 1]).items()})
    got.append({k: a.copy() for k, a in pipe.wait(tickets[-1]).items()})
    for g, w in zip(got, want):
        for k in fields:
            assert np.array_equal(g[k], w[k]), k
    with pytest.raises(ValueError):
        pipe.wait(tickets[0])          # long out of flight


def test_baseline_config3_nordschleife_every_waypoint_sweep():
    """BASELINE.json configs[2]: one instance per metre of the Nordschleife centreline (~20.8 k instances,
    unperturbed), every instance checked against the oracle (the C port does the sweep in under a second)."""
    import _golden

    kw = _golden.racing_kwargs("nordschleife")
    cl = tracks.synthetic_centreline("nordschleife")
    idx = np.arange(0, cl.shape[0], 2)           # centreline sampled every 0.5 m -> one instance per metre
    paths = tracks.make_instances(cl, idx, 50)
    assert paths.shape[0] > 20000
    got = _solver(**kw).solve_host(paths)
    want = port.solve_batch(port.default_config(**kw), paths, nthreads=16)
    _assert_equals_oracle(got, want)
    assert (got["status"] == 1).mean() > 0.99


@pytest.mark.parametrize("H", [20, 40, 80])
def test_baseline_config4_spa_horizon_sweep_batch_16k(H):
    """BASELINE.json configs[3]: Spa racing block, horizon 20 / 40 / 80, 16384 perturbed instances each."""
    import _golden

    kw = _golden.racing_kwargs("spa", H)
    paths, vmax = tracks.perturbed_batch("spa", 16384, horizon=H, seed=3)
    got = _solver(**kw).solve_host(paths, None, vmax, fields=["controls", "states", "v_ref", "cost", "status",
                                                               "status_speed", "iters", "rho_updates"])
    want = port.solve_batch(port.default_config(**kw), paths, None, vmax, nthreads=16)
    _assert_equals_oracle(got, want, fields=("controls", "states", "v_ref", "cost"))


@pytest.mark.parametrize("track", tracks.TRACK_ORDER)
def test_baseline_config5_all_tracks_shards(track):
    """BASELINE.json configs[4] (all 7 tracks, 1 M instances over 2/4/8 GPUs), scaled to what one GPU and
    the oracle check in seconds: the per-track share of the sweep is cut into contiguous rank shards by the
    same helper the multi-GPU bench uses, each shard solved by its own launch and the gathered result
    compared with the oracle on the unsharded batch."""
    import _golden
    from ac_mpc_b200 import sharding

    kw = _golden.racing_kwargs(track)
    B = 8192 + 37
    paths, vmax = tracks.perturbed_batch(track, B, seed=5)
    mpc = _solver(**kw)
    parts = []
    for r in range(4):
        lo, hi = sharding.shard_range(B, r, 4)
        parts.append(mpc.solve_host(paths[lo:hi], None, vmax[lo:hi], fields=["controls", "status", "iters", "cost",
                                                                              "status_speed", "rho_updates"]))
    got = {k: np.concatenate([p[k] for p in parts]) for k in parts[0] if not k.startswith("_")}
    want = port.solve_batch(port.default_config(**kw), paths, None, vmax, nthreads=16)
    _assert_equals_oracle(got, want, fields=("controls", "cost"))


def test_warm_start_object_reuse_matches_reference_golden():
    """The reference test reuses ONE SpatialMPC for its 28 fixture paths (warm-started OSQP objects, carried
    rho): golden group fixture_warm, through the drop-in object API and the host entry point (keep_warm)."""
    import _golden
    from ac_mpc_b200.control import build_mpc

    G = _golden.load()
    g, paths = G["fixture_warm"], G["fixture_cold"]["paths"]
    veh = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                         "max_steering_angle": lambda self: 0.3})()
    cfg = {"horizon": 100, "step_cost": _golden.FIXTURE_CONFIG["step_cost"], "r_term": [1e-2, 10.0],
           "final_cost": [1.0, 0.0, 0.1],
           "speed_profile_constraints": {k: _golden.FIXTURE_CONFIG[k] for k in ("v_min", "v_max", "a_min", "a_max", "ay_max", "ki_min", "end_velocity")}}
    mpc = build_mpc(cfg, veh)
    for b in range(paths.shape[0]):
        mpc.get_control(paths[b])
        assert mpc.last_info["iters"] == g["iters"][b].tolist(), b
        assert mpc.last_info["rho_updates"] == g["rho_updates"][b].tolist(), b
        if g["status"][b] == 1:
            np.testing.assert_allclose(mpc.projected_control, g["controls"][b], rtol=0, atol=TOL)
            np.testing.assert_allclose(mpc.current_prediction, g["prediction"][b], rtol=0, atol=TOL)


def test_warm_start_closed_loop_replay_matches_oracle_objects():
    """BASELINE configs[2] "closed-loop replay": instances advance along the Nordschleife centreline step by step,
    warm-started from their previous solve (device entry point, one record per instance), alternating
    is_localised (two separate speed-solver objects); a sample of the records is checked against oracle
    objects driven through the same sequence, and warm_valid=0 must reproduce the cold solve."""
    import _golden
    import torch

    kw = _golden.racing_kwargs("nordschleife")
    cl = tracks.synthetic_centreline("nordschleife")
    B, steps = 2048, 4
    start = np.random.default_rng(9).integers(0, cl.shape[0], B)
    mpc = _solver(**kw)
    warm = mpc.alloc_warm(B)
    assert warm.numel() * 8 == B * mpc.warm_stride()
    sample = np.random.default_rng(1).choice(B, 24, replace=False)
    objs = {int(b): port.PortMPC(port.default_config(**kw)) for b in sample}
    packed, views = mpc.alloc_device_outputs(B, ["controls", "status", "iters", "rho_updates", "cost"])
    for t in range(steps):
        paths = tracks.make_instances(cl, (start + 30 * t) % cl.shape[0], 50)
        loc = bool(t % 2)
        mpc.solve_device(torch.from_numpy(paths).cuda(), None, None, loc, out=views, warm=warm)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in views.items()}
        for b, obj in objs.items():
            want = obj.step(paths[b], 0.0, None, loc, warm=True)
            assert got["iters"][b].tolist() == want["iters"].tolist(), (t, b)
            assert got["rho_updates"][b].tolist() == want["rho_updates"].tolist(), (t, b)
            assert got["status"][b] == want["status"]
            np.testing.assert_allclose(got["controls"][b], want["controls"], rtol=0, atol=TOL)
    # warm starts change the iteration counts of most instances after the first step
    cold = mpc.solve_host(paths, None, None, loc, fields=["controls", "iters"])
    assert (got["iters"] != cold["iters"]).any()
    mpc.solve_device(torch.from_numpy(paths).cuda(), None, None, loc, out=views, warm=warm, warm_valid=False)
    torch.cuda.synchronize()
    assert np.array_equal(views["iters"].cpu().numpy(), cold["iters"])
    assert np.array_equal(views["controls"].cpu().numpy(), cold["controls"])


def test_large_batch_131072_instances():
    """One launch far beyond what the device holds at once (work queue + longest-first order over ~110 rounds):
    every instance solved exactly once, a random sample checked against the oracle."""
    B = 131072
    base, vm = tracks.perturbed_batch("monza", 8192, seed=12)
    rep = B // base.shape[0]
    paths, vmax = np.tile(base, (rep, 1, 1)), np.tile(vm, rep)
    mpc = _solver()
    out = mpc.solve_host(paths, None, vmax, fields=["controls", "status", "iters", "cost"])
    # the batch is 16 copies of the same 8192 instances: every copy must give bit-identical results
    for k in ("controls", "status", "iters", "cost"):
        a = out[k].reshape((rep, base.shape[0]) + out[k].shape[1:])
        assert np.array_equal(a, np.broadcast_to(a[0], a.shape)), k
    sub = np.random.default_rng(2).choice(base.shape[0], 512, replace=False)
    want = port.solve_batch(port.default_config(), base[sub], None, vm[sub], nthreads=16)
    assert np.array_equal(out["iters"][sub], want["iters"])
    np.testing.assert_allclose(out["controls"][sub], want["controls"], rtol=0, atol=TOL)


SETTINGS = [dict(scaling=0), dict(scaling=3), dict(check_termination=1), dict(check_termination=10),
            dict(check_termination=7, adaptive_rho_interval=20), dict(adaptive_rho=0), dict(adaptive_rho_interval=25),
            dict(alpha=1.0), dict(rho=1.0), dict(eps_abs=1e-5, eps_rel=1e-5), dict(max_iter=60), dict(max_iter=1),
            dict(check_termination=0, max_iter=80),
            # OSQP 1.x termination semantics (duality-gap test on top of the residual tests), alone and combined
            dict(check_dualgap=1), dict(check_dualgap=1, eps_abs=1e-5, eps_rel=1e-5), dict(check_dualgap=1, scaling=0),
            dict(check_dualgap=1, adaptive_rho_interval=25, check_termination=5)]


@pytest.mark.parametrize("kw", SETTINGS, ids=[",".join(f"{k}={v}" for k, v in s.items()) for s in SETTINGS])
def test_osqp_settings_sweep_matches_oracle(kw):
    """Every OSQP setting of acmpc_config away from its default (see the emulation test of the same name)."""
    paths, vmax = tracks.perturbed_batch("monza", 96, seed=7)
    got = _solver(**kw).solve_host(paths, None, vmax, False)
    want = port.solve_batch(port.default_config(**kw), paths, None, vmax, False, nthreads=8)
    for k in ("status", "status_speed", "iters", "rho_updates"):
        assert np.array_equal(got[k], want[k]), k
    ok = (want["status"] == 1) & (want["status_speed"] == 1)
    if ok.any():
        np.testing.assert_allclose(got["controls"][ok], want["controls"][ok], rtol=0, atol=TOL)
    np.testing.assert_allclose(got["cost"], want["cost"], rtol=1e-7, atol=1e-7)


def test_track_narrower_than_the_car_is_rejected_like_osqp_does():
    paths, vmax = tracks.perturbed_batch("monza", 16, seed=7)
    narrow = paths.copy()
    narrow[:, :, 2] = 1.9
    got = _solver().solve_host(narrow, None, vmax)
    want = port.solve_batch(port.default_config(), narrow, None, vmax, nthreads=2)
    assert np.all(got["status"] == -10) and np.all(want["status"] == -10)
    assert np.array_equal(got["iters"], want["iters"]) and np.array_equal(got["status_speed"], want["status_speed"])
    np.testing.assert_allclose(got["v_ref"], want["v_ref"], rtol=0, atol=TOL)


def test_localised_batch_with_offsets_through_the_ordered_pipelined_path():
    """2560 instances: large enough for the longest-first order kernel and for the host entry point's 4-chunk
    pipeline, with is_localised=True, per-instance offsets and v_max, against the oracle."""
    import _golden

    kw = _golden.racing_kwargs("silverstone")
    B = 2560
    paths, vmax = tracks.perturbed_batch("silverstone", B, seed=31)
    offs = np.random.default_rng(8).uniform(-0.6, 0.6, B)
    got = _solver(**kw).solve_host(paths, offs, vmax, True)
    want = port.solve_batch(port.default_config(**kw), paths, offs, vmax, True, nthreads=16)
    _assert_equals_oracle(got, want)
    np.testing.assert_allclose(got["waypoints"], want["waypoints"], rtol=0, atol=TOL)
