"""-m gpu: one acmpc_handle serves the drop-in object's get_control (host entry point, persistent warm-start records),
device-resident batches, profiling and the longest-first ordering at the same time.  Regression tests for the round-1
defect where the device entry point's "grow the hand-over buffer" branch re-ran acmpc_create's initialisation
(warm-start records dropped -> silent cold start, order buffers and profiling events leaked and switched off), and
for the separation of the two entry points' work queues."""
import numpy as np
import pytest

import _golden
from ac_mpc_b200 import BatchedMPC, _capi, tracks
from oracle import port

pytestmark = pytest.mark.gpu

VEH = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                     "max_steering_angle": lambda self: 0.3})()


def _fixture_mpc():
    from ac_mpc_b200.control import build_mpc

    cfg = {"horizon": 100, "step_cost": _golden.FIXTURE_CONFIG["step_cost"], "r_term": [1e-2, 10.0],
           "final_cost": [1.0, 0.0, 0.1],
           "speed_profile_constraints": {k: _golden.FIXTURE_CONFIG[k] for k in ("v_min", "v_max", "a_min", "a_max", "ay_max",
                                                                                 "ki_min", "end_velocity")}}
    return build_mpc(cfg, VEH)


def test_device_call_without_v_ref_keeps_the_objects_warm_start_records():
    """get_control x3, then a device-entry solve WITHOUT v_ref on the same handle (first use: its hand-over buffer has to
    grow), then get_control again: every get_control must equal an oracle object driven through the same four steps
    (the reference's persistent OSQP objects, spatial_mpc.py:43-58) -- in particular the 4th is NOT a cold start."""
    import torch

    G = _golden.load()
    paths = G["fixture_cold"]["paths"]
    mpc = _fixture_mpc()
    obj = port.PortMPC(port.default_config(**_golden.FIXTURE_CONFIG))
    cold_iters = G["fixture_cold"]["iters"]
    seq = [0, 8, 15, 22]
    differs = False
    for step, b in enumerate(seq):
        if step == 3:
            h = mpc._batched()
            pb, vm = tracks.perturbed_batch("monza", 300, horizon=100, seed=3)
            _, views = h.alloc_device_outputs(300, ["controls", "status"])
            h.solve_device(torch.from_numpy(pb).cuda(), None, torch.from_numpy(vm).cuda(), False, out=views)
            torch.cuda.synchronize()
            assert int((views["status"] == 1).sum()) > 250
        mpc.get_control(paths[b])
        want = obj.step(paths[b], 0.0, None, False, warm=True)
        assert mpc.last_info["iters"] == want["iters"].tolist(), (step, b)
        assert mpc.last_info["rho_updates"] == want["rho_updates"].tolist(), (step, b)
        np.testing.assert_allclose(mpc.projected_control, want["controls"], rtol=0, atol=1e-7)
        if step == 3:
            differs = want["iters"].tolist() != cold_iters[b].tolist() or \
                np.abs(want["controls"] - G["fixture_cold"]["controls"][b]).max() > 1e-6
    assert differs, "the chosen sequence must distinguish a warm 4th step from a cold one"


def test_profiling_and_ordering_survive_a_buffer_growth():
    import torch

    mpc = BatchedMPC(_capi.default_config(), device=0)
    paths, vmax = tracks.perturbed_batch("monza", 4096, seed=1)
    dp, dv = torch.from_numpy(paths).cuda(), torch.from_numpy(vmax).cuda()
    mpc.set_profiling(True)
    # first device call of the handle, no v_ref requested: the hand-over buffer grows inside this call
    _, views = mpc.alloc_device_outputs(4096, ["controls", "status", "iters"])
    mpc.solve_device(dp, None, dv, False, out=views)
    torch.cuda.synchronize()
    assert mpc.launch_info()["launches"] == 3, "order kernel + speed kernel + control kernel (longest-first still on)"
    ms = mpc.collect_kernel_ms()
    assert ms["launches"] == 1 and ms["speed_ms"] > 0 and ms["control_ms"] > 0, "profiling must survive the growth"
    # a larger batch grows it again; profiling is still on afterwards
    big = torch.from_numpy(np.tile(paths, (2, 1, 1))).cuda()
    _, v2 = mpc.alloc_device_outputs(8192, ["controls", "status", "iters"])
    mpc.solve_device(big, None, torch.from_numpy(np.tile(vmax, 2)).cuda(), False, out=v2)
    torch.cuda.synchronize()
    assert mpc.launch_info()["launches"] == 3
    assert mpc.collect_kernel_ms()["launches"] == 1
    assert np.array_equal(v2["iters"][:4096].cpu().numpy(), views["iters"].cpu().numpy())
    assert np.array_equal(v2["controls"][4096:].cpu().numpy(), views["controls"].cpu().numpy())


def test_host_and_device_entry_points_interleave_on_one_handle():
    """Device-entry launches queued on the caller's stream while host-entry calls run on the handle's own streams:
    separate ticket counters and order buffers, so every instance of both is solved exactly once."""
    import torch

    mpc = BatchedMPC(_capi.default_config(), device=0)
    cfg = port.default_config()
    pa, va = tracks.perturbed_batch("monza", 5000, seed=11)
    pb, vb = tracks.perturbed_batch("monza", 3000, seed=12)
    want_a = port.solve_batch(cfg, pa, None, va, nthreads=16)
    want_b = port.solve_batch(cfg, pb, None, vb, nthreads=16)
    dpa, dva = torch.from_numpy(pa).cuda(), torch.from_numpy(va).cuda()
    _, views = mpc.alloc_device_outputs(5000, ["controls", "status", "iters"])
    side = torch.cuda.Stream()
    for _ in range(3):
        for v in views.values():
            v.zero_()
        torch.cuda.synchronize()
        mpc.solve_device(dpa, None, dva, False, out=views, stream=side)      # asynchronous
        got_b = mpc.solve_host(pb, None, vb, fields=["controls", "status", "iters"])   # runs while the other is in flight
        torch.cuda.synchronize()
        assert np.array_equal(views["iters"].cpu().numpy(), want_a["iters"])
        assert np.array_equal(got_b["iters"], want_b["iters"])
        np.testing.assert_allclose(views["controls"].cpu().numpy(), want_a["controls"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(got_b["controls"], want_b["controls"], rtol=0, atol=1e-7)


def test_host_outputs_are_validated_before_their_pointers_reach_the_library():
    mpc = BatchedMPC(_capi.default_config(), device=0)
    paths, vmax = tracks.perturbed_batch("monza", 8, seed=1)
    good = mpc.alloc_host_outputs(8, ["controls", "status"])
    mpc.solve_host(paths, None, vmax, out=good)
    for bad in (dict(controls=np.zeros((7, 2, 49))), dict(controls=np.zeros((8, 2, 49), np.float32)),
                dict(controls=np.zeros((8, 49, 2)).transpose(0, 2, 1)), dict(status=np.zeros(8, np.int64)),
                dict(nonsense=np.zeros(8))):
        with pytest.raises(ValueError):
            mpc.solve_host(paths, None, vmax, out=bad)


def test_track_map_keeps_extra_columns(tmp_path):
    """utils/load.py:30-35 returns track[is_not_duplicated] with every column: a map with a third column (z, width)."""
    from ac_mpc_b200.utils import load

    rng = np.random.default_rng(4)
    xy = np.cumsum(rng.uniform(0.2, 1.0, (500, 2)), axis=0)
    xy[100] = xy[99] + 5e-5
    xy[200:203] = xy[199]
    full = np.column_stack([xy, rng.uniform(8, 12, 500), np.arange(500.0)])
    d = np.hypot(*np.diff(full[:, :2], axis=0).T)
    keep = np.r_[True, d > 1e-4]
    p = str(tmp_path / "map.npy")
    load.save_track_map(p, full, full[::-1].copy(), full[:, :3].copy())
    got = load.track_map(p, BatchedMPC(_capi.default_config(), device=0))
    assert np.array_equal(got["centre"], full[keep]) and got["centre"].shape[1] == 4
    assert got["right"].shape[1] == 3 and np.array_equal(got["right"], full[keep][:, :3])
