"""Generate tests/golden/track_golden.npz: the track side of the MPC step computed
by the UNMODIFIED reference functions -- utils/load.py (track_map / remove_near_duplicate_points, read from a map file
written here in the reference's own on-disk format), perception/utils.py:smooth_track_with_polyfit and
TrackLimitPerception._calculate_centre_track (perception/tracks.py:247-252, bound to a stand-in object: the class
itself needs the camera configuration and the simulator stack to construct).

    python tests/golden/make_track_golden.py        # needs /root/reference, so it cannot run on the GPU box
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = ["/root/reference/src", os.path.join(ROOT, "oracle", "shim"), ROOT]
# ruamel.yaml (only used by load.yaml) is absent here
ruamel = types.ModuleType("ruamel")
ruamel.yaml = types.ModuleType("ruamel.yaml")
ruamel.yaml.YAML = object
sys.modules.setdefault("ruamel", ruamel)
sys.modules.setdefault("ruamel.yaml", ruamel.yaml)

from acmpc.perception.tracks import TrackLimitPerception  # noqa: E402
from acmpc.perception.utils import smooth_track_with_polyfit  # noqa: E402
from acmpc.utils import load  # noqa: E402


def perceived_limit(rng, side, n):
    """A bird's-eye-view track limit as the segmentation hands it over: unordered-ish samples of a gently curving line
    8..140 m ahead, `side` * ~5 m off the ego axis, with pixel-quantisation noise."""
    y = np.sort(rng.uniform(rng.uniform(2, 12), rng.uniform(60, 140), n))
    c = rng.uniform(-4e-4, 4e-4), rng.uniform(-0.05, 0.05), side * rng.uniform(3.5, 6.5)
    x = c[0] * y * y + c[1] * y + c[2] + rng.normal(0, 0.15, n)
    return np.stack([x, y], axis=1)


def main():
    out = {}
    rng = np.random.default_rng(11)

    # -- map file in the reference's format, with planted near-duplicates -------------------------------------------
    from ac_mpc_b200 import tracks
    cl = tracks.synthetic_centreline("vallelunga")[::4]
    tang = np.roll(cl, -1, axis=0) - cl
    nrm = np.stack([-tang[:, 1], tang[:, 0]], axis=1) / np.linalg.norm(tang, axis=1, keepdims=True)
    lines = {"centre_track": cl, "outside_track": cl + 5.0 * nrm, "inside_track": cl - 5.0 * nrm}
    for name, line in lines.items():
        dup = np.sort(rng.choice(line.shape[0] - 1, 200, replace=False))
        jitter = rng.choice([0.0, 2e-5, 9.9e-5, 1.01e-4, 3e-4], 200)[:, None] * rng.normal(size=(200, 2))
        lines[name] = np.insert(line, dup + 1, line[dup] + jitter, axis=0)
        # runs of duplicates: every row is compared with its predecessor in the INPUT
        lines[name] = np.insert(lines[name], [10, 10, 10], lines[name][9] + 1e-6, axis=0)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "synthetic_vallelunga.npy")
        np.save(path, lines, allow_pickle=True)
        loaded = load.track_map(path)
    for name, key in (("centre_track", "centre"), ("outside_track", "left"), ("inside_track", "right")):
        out[f"map/{name}"] = lines[name]
        out[f"map/{key}"] = loaded[key]

    # -- smooth_track_with_polyfit ---------------------------------------------------------------------------------
    for degree in (2, 3):
        tr, res = [], []
        for b in range(64):
            n = int(rng.integers(6, 400))
            t = perceived_limit(rng, rng.choice([-1.0, 1.0]), n)
            if b % 16 == 5:
                t = np.zeros((0, 2))            # nothing seen: the stub line
            if b % 16 == 9:
                t[:, 0] -= t[:, 0].mean()       # the fitted line passes close to the origin
            tr.append(t)
            res.append(smooth_track_with_polyfit(t, 500, degree))
        out[f"polyfit{degree}/offsets"] = np.concatenate([[0], np.cumsum([t.shape[0] for t in tr])]).astype(np.int32)
        out[f"polyfit{degree}/points"] = np.concatenate(tr, axis=0)
        out[f"polyfit{degree}/expected"] = np.stack(res)
    t = perceived_limit(rng, 1.0, 120)
    for n_pts in (1, 2, 50, 84):
        out[f"polyfit_n{n_pts}/points"] = t
        out[f"polyfit_n{n_pts}/expected"] = smooth_track_with_polyfit(t, n_pts, 2)

    # -- _calculate_centre_track --------------------------------------------------------------------------------------
    stub = types.SimpleNamespace(_n_polyfit_points=500)
    stub._smooth_track_with_polyfit = types.MethodType(TrackLimitPerception._smooth_track_with_polyfit, stub)
    lefts, rights, centres = [], [], []
    for b in range(48):
        tracks_ = {"left": perceived_limit(rng, -1.0, int(rng.integers(20, 300))),
                   "right": perceived_limit(rng, 1.0, int(rng.integers(20, 300)))}
        tracks_["left"] = stub._smooth_track_with_polyfit(tracks_["left"])        # tracks.py:224-225
        tracks_["right"] = stub._smooth_track_with_polyfit(tracks_["right"])
        lefts.append(tracks_["left"]), rights.append(tracks_["right"])
        centres.append(TrackLimitPerception._calculate_centre_track(stub, tracks_))
    out["centre/left"], out["centre/right"], out["centre/expected"] = np.stack(lefts), np.stack(rights), np.stack(centres)

    np.savez_compressed(os.path.join(HERE, "track_golden.npz"), **out)
    print("wrote track_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    with warnings.catch_warnings():
        warnings.simplefilter("error", np.exceptions.RankWarning)
        main()
