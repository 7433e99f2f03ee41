"""Track side of the MPC step: the map format + near-duplicate removal of
utils/load.py:9-35, smooth_track_with_polyfit (perception/utils.py:107-119), _calculate_centre_track
(perception/tracks.py:247-252) and the on-device instance extraction of SURVEY.md section 8d.

Golden vectors: tests/golden/track_golden.npz, produced by the UNMODIFIED reference functions
(tests/golden/make_track_golden.py).  Bars: bit-exact for the duplicate removal (a selection of input rows);
1e-9 m for the polynomial smoothing (floating point: the reference solves the least-squares problem by numpy's SVD, the
kernel by an orthogonal-polynomial recurrence -- same polynomial, different rounding) and for the extracted paths."""
import os

import numpy as np
import pytest

from ac_mpc_b200 import _capi, tracks
from oracle import track_prep as oracle

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "track_golden.npz")
TOL = 1e-9   # metres


def _group(g):
    with np.load(_PATH) as z:
        return {k.split("/")[1]: z[k] for k in z.files if k.startswith(g + "/")}


def _ragged(d):
    off = d["offsets"]
    return [d["points"][off[b]:off[b + 1]] for b in range(off.shape[0] - 1)]


# ---- CPU: the oracle restatement against the reference's output --------------------------------------------------
def test_oracle_duplicate_removal_matches_reference():
    d = _group("map")
    for raw, key in (("centre_track", "centre"), ("outside_track", "left"), ("inside_track", "right")):
        got = oracle.remove_near_duplicate_points(d[raw])
        assert got.shape[0] < d[raw].shape[0]
        np.testing.assert_array_equal(got, d[key])


def test_oracle_track_map_reads_the_reference_format(tmp_path):
    d = _group("map")
    p = str(tmp_path / "m.npy")
    np.save(p, {k: d[k] for k in ("centre_track", "outside_track", "inside_track")}, allow_pickle=True)
    got = oracle.track_map(p)
    for key in ("centre", "left", "right"):
        np.testing.assert_array_equal(got[key], d[key])


@pytest.mark.parametrize("degree", [2, 3])
def test_oracle_polyfit_matches_reference(degree):
    d = _group(f"polyfit{degree}")
    for t, want in zip(_ragged(d), d["expected"]):
        got, _ = oracle.smooth_track_with_polyfit(t, 500, degree)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    for n in (1, 2, 50, 84):
        e = _group(f"polyfit_n{n}")
        np.testing.assert_allclose(oracle.smooth_track_with_polyfit(e["points"], n, 2)[0], e["expected"], rtol=0, atol=1e-12)


def test_oracle_centre_track_matches_reference():
    d = _group("centre")
    for l, r, want in zip(d["left"][:12], d["right"][:12], d["expected"][:12]):
        np.testing.assert_allclose(oracle.calculate_centre_track(l, r, 500), want, rtol=0, atol=1e-12)


def test_oracle_instance_matches_the_batch_generator():
    cl = tracks.synthetic_centreline("vallelunga")
    rng = np.random.default_rng(5)
    idx = np.concatenate([[0, cl.shape[0] - 1, cl.shape[0] - 150], rng.integers(0, cl.shape[0], 13)])
    lat, psi = rng.uniform(-2, 2, idx.shape[0]), rng.uniform(-0.1, 0.1, idx.shape[0])
    want = tracks.make_instances(cl, idx, 50, lat, psi)
    for b in range(idx.shape[0]):
        np.testing.assert_allclose(oracle.make_instance(cl, int(idx[b]), 50, lat[b], psi[b]), want[b], rtol=0, atol=TOL)


def test_track_entry_points_exported():
    L = _capi.load()
    assert L.acmpc_remove_near_duplicates_host(None, 1, None, 1e-4, None, None) == 1
    assert L.acmpc_smooth_tracks_polyfit_host(None, 1, None, None, 500, 2, None, None, None) == 1
    assert L.acmpc_centre_tracks_host(None, 1, 500, None, None, 500, None, None) == 1
    assert L.acmpc_extract_paths_host(None, 10, None, 1, None, None, None, 100.0, 0.5, None) == 1
    assert L.acmpc_extract_paths_device(None, 10, None, 1, None, None, None, 100.0, 0.5, None, None) == 1


# ---- GPU --------------------------------------------------------------------------------------------------------
def _solver(**kw):
    from ac_mpc_b200 import BatchedMPC

    return BatchedMPC(_capi.default_config(**kw), device=0)


@pytest.mark.gpu
def test_duplicate_removal_matches_reference_golden():
    s, d = _solver(), _group("map")
    for raw, key in (("centre_track", "centre"), ("outside_track", "left"), ("inside_track", "right")):
        np.testing.assert_array_equal(s.remove_near_duplicate_points(d[raw]), d[key])


@pytest.mark.gpu
def test_duplicate_removal_edges_and_large():
    s = _solver()
    one = np.array([[3.0, 4.0]])
    np.testing.assert_array_equal(s.remove_near_duplicate_points(one), one)
    assert s.remove_near_duplicate_points(np.zeros((0, 2))).shape == (0, 2)
    same = np.tile([[1.0, 2.0]], (1000, 1))
    np.testing.assert_array_equal(s.remove_near_duplicate_points(same), same[:1])       # everything after row 0 goes
    rng = np.random.default_rng(0)
    big = np.cumsum(rng.normal(0, 0.3, (1_000_003, 2)), axis=0)
    dup = rng.random(big.shape[0]) < 0.3
    dup[0] = False
    big[dup] = big[np.maximum.accumulate(np.where(dup, 0, np.arange(big.shape[0])))][dup] + 3e-5   # runs of copies
    d = np.diff(big, axis=0)
    keep = np.ones(big.shape[0], bool)
    keep[1:] = np.hypot(d[:, 0], d[:, 1]) > 0.0001                                      # load.py:31-34, vectorised
    got = s.remove_near_duplicate_points(big)
    np.testing.assert_array_equal(got, big[keep])
    np.testing.assert_array_equal(s.remove_near_duplicate_points(got), oracle_fixed_point(got))


def oracle_fixed_point(track):
    d = np.diff(track, axis=0)
    keep = np.ones(track.shape[0], bool)
    keep[1:] = np.hypot(d[:, 0], d[:, 1]) > 0.0001
    return track[keep]


@pytest.mark.gpu
def test_track_map_loader_matches_reference_golden(tmp_path):
    from ac_mpc_b200.utils import load

    d = _group("map")
    p = str(tmp_path / "synthetic.npy")
    load.save_track_map(p, d["centre_track"], d["outside_track"], d["inside_track"])
    got = load.track_map(p, _solver())
    for key in ("centre", "left", "right"):
        np.testing.assert_array_equal(got[key], d[key])
    assert load.find_map("synthetic", str(tmp_path)) == p and load.find_map("monza", str(tmp_path)) is None
    cl = tracks.centreline("synthetic", str(tmp_path), ds=0.5)
    seg = np.linalg.norm(np.diff(cl, axis=0), axis=1)
    assert abs(seg.mean() - 0.5) < 1e-3 and seg.max() < 0.51


@pytest.mark.gpu
@pytest.mark.parametrize("degree", [2, 3])
def test_polyfit_matches_reference_golden(degree):
    s, d = _solver(), _group(f"polyfit{degree}")
    trs = _ragged(d)
    got, status, start = s.smooth_tracks_with_polyfit(trs, 500, degree, return_info=True)
    np.testing.assert_allclose(got, d["expected"], rtol=0, atol=TOL)
    want_status = np.array([1 if t.shape[0] == 0 else 0 for t in trs])
    np.testing.assert_array_equal(status, want_status)
    want_start = np.array([oracle.smooth_track_with_polyfit(t, 500, degree)[1] for t in trs])
    np.testing.assert_array_equal(start, want_start)


@pytest.mark.gpu
def test_polyfit_single_track_signature_and_point_counts():
    from ac_mpc_b200.perception import utils as putils

    s = _solver()
    for n in (1, 2, 50, 84):
        e = _group(f"polyfit_n{n}")
        got = putils.smooth_track_with_polyfit(e["points"], n, 2, solver=s)
        assert got.shape == (n, 2)
        np.testing.assert_allclose(got, e["expected"], rtol=0, atol=TOL)
    stub = putils.smooth_track_with_polyfit(np.zeros((0, 2)), 7, 2, solver=s)           # perception/utils.py:108-111
    np.testing.assert_array_equal(stub, np.array([np.linspace(0, 0.1, 7), np.linspace(0, 2, 7)]).T)


@pytest.mark.gpu
def test_polyfit_rank_deficient_is_flagged():
    s = _solver()
    t = np.array([[1.0, 5.0], [1.2, 5.0], [3.0, 20.0], [3.1, 20.0], [2.9, 20.0]])       # two distinct abscissae
    got, status, _ = s.smooth_tracks_with_polyfit([t], 50, 2, return_info=True)
    assert status[0] == 2
    np.testing.assert_allclose(got[0], oracle.smooth_track_with_polyfit(t, 50, 1)[0], rtol=0, atol=TOL)


@pytest.mark.gpu
def test_centre_tracks_match_reference_golden():
    from ac_mpc_b200.perception import utils as putils

    s, d = _solver(), _group("centre")
    np.testing.assert_allclose(s.centre_tracks(d["left"], d["right"]), d["expected"], rtol=0, atol=TOL)
    one = putils.calculate_centre_track({"left": d["left"][3], "right": d["right"][3]}, solver=s)
    np.testing.assert_allclose(one, d["expected"][3], rtol=0, atol=TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("track,horizon", [("monza", 50), ("vallelunga", 20), ("spa", 80)])
def test_extract_paths_matches_oracle(track, horizon):
    s = _solver(horizon=horizon)
    cl = tracks.synthetic_centreline(track)
    rng = np.random.default_rng(3)
    B = 4096
    idx = rng.integers(0, cl.shape[0], B)
    idx[:3] = [0, cl.shape[0] - 1, cl.shape[0] - 100]                                   # wrap-around of the closed loop
    lat, psi = rng.uniform(-2, 2, B), rng.uniform(-0.1, 0.1, B)
    got = s.extract_paths(cl, idx, lat, psi)
    np.testing.assert_allclose(got, tracks.make_instances(cl, idx, horizon, lat, psi), rtol=0, atol=TOL)
    for b in range(8):
        np.testing.assert_allclose(got[b], oracle.make_instance(cl, int(idx[b]), horizon, lat[b], psi[b]), rtol=0, atol=TOL)
    np.testing.assert_allclose(s.extract_paths(cl, idx[:16]), tracks.make_instances(cl, idx[:16], horizon), rtol=0, atol=TOL)


@pytest.mark.gpu
def test_device_sweep_never_touches_the_host():
    """extract_paths_device -> solve_device on resident tensors == the host entry point fed the same paths, and the
    controls agree with the CPU oracle solving the host-generated instances (1e-3, the north-star bar)."""
    import torch

    from oracle import port

    s = _solver()
    cl = tracks.synthetic_centreline("monza")
    rng = np.random.default_rng(9)
    B = 512
    idx = rng.integers(0, cl.shape[0], B).astype(np.int32)
    lat, psi, vmax = rng.uniform(-2, 2, B), rng.uniform(-0.1, 0.1, B), rng.uniform(20, 84, B)
    dev = torch.device("cuda:0")
    d_paths = s.extract_paths_device(torch.from_numpy(cl).to(dev), torch.from_numpy(idx).to(dev),
                                     torch.from_numpy(lat).to(dev), torch.from_numpy(psi).to(dev))
    out = s.solve_device(d_paths, None, torch.from_numpy(vmax).to(dev), False)
    torch.cuda.synchronize()
    paths = d_paths.cpu().numpy()
    host = s.solve_host(paths, None, vmax, False)
    np.testing.assert_array_equal(out["controls"].cpu().numpy(), host["controls"])
    np.testing.assert_array_equal(out["status"].cpu().numpy(), host["status"])
    want = port.solve_batch(port.default_config(), tracks.make_instances(cl, idx, 50, lat, psi), None, vmax, False, nthreads=4)
    np.testing.assert_array_equal(host["status"], want["status"])
    assert np.abs(host["controls"] - want["controls"]).max() < 1e-3
