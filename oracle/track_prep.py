"""CPU restatement of the track side of the MPC step.

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/ may import this module.  Pinned against the reference's own
functions run in the build container (tests/golden/make_track_golden.py -> tests/golden/track_golden.npz).

  remove_near_duplicate_points  /root/reference/src/acmpc/utils/load.py:30-35
  track_map                     /root/reference/src/acmpc/utils/load.py:9-27,61-65
  smooth_track_with_polyfit     /root/reference/src/acmpc/perception/utils.py:107-119
  calculate_centre_track        /root/reference/src/acmpc/perception/tracks.py:247-252
  make_instance                 SURVEY.md section 8d "instance -> get_control input" (scalar loops, small cases only)
"""
from __future__ import annotations

import math

import numpy as np


def remove_near_duplicate_points(track: np.ndarray, tol: float = 0.0001) -> np.ndarray:
    """load.py:30-35: row 0 stays, row i stays iff hypot(track[i] - track[i-1]) > tol."""
    keep = [True]
    for i in range(1, track.shape[0]):
        keep.append(math.hypot(track[i, 0] - track[i - 1, 0], track[i, 1] - track[i - 1, 1]) > tol)
    return track[np.array(keep[:track.shape[0]], dtype=bool)]


def track_map(path: str) -> dict:
    """load.py:9-27: {"outside_track","inside_track","centre_track"} pickled in a .npy -> left / right / centre."""
    d = np.load(path, allow_pickle=True).item()
    return {"left": remove_near_duplicate_points(d["outside_track"]),
            "right": remove_near_duplicate_points(d["inside_track"]),
            "centre": remove_near_duplicate_points(d["centre_track"])}


def _linspace(start: float, stop: float, num: int) -> list:
    """numpy.linspace as numpy computes it (arange * step + start, last sample = stop)."""
    if num == 1:
        return [start]
    div, delta = num - 1, stop - start
    step = delta / div
    ys = [(j / div) * delta + start if step == 0 else j * step + start for j in range(num)]
    ys[-1] = stop
    return ys


def smooth_track_with_polyfit(track: np.ndarray, num_points: int, degree: int = 3):
    """perception/utils.py:107-119.  The least-squares fit is numpy's (np.polyfit: SVD of the column-scaled
    Vandermonde matrix), the third-party arithmetic the reference itself calls; the rest is restated with scalar
    loops.  Returns (points (num_points, 2), start_index)."""
    if len(track) == 0:
        return np.array([_linspace(0, 0.1, num_points), _linspace(0, 2, num_points)]).T, 0
    ymax = max(float(v) for v in track[:, 1])
    coeffs = [float(c) for c in np.polyfit(track[:, 1], track[:, 0], degree)]

    def horner(y):
        acc = 0.0
        for c in coeffs:
            acc = acc * y + c
        return acc

    scan = _linspace(0, ymax, 500)
    best, start = math.inf, 0
    for j, y in enumerate(scan):
        x = horner(y)
        r = math.sqrt(x * x + y * y)
        if r < best:
            best, start = r, j
    ys = _linspace(scan[start], ymax, num_points)
    return np.array([[horner(y), y] for y in ys]), start


def calculate_centre_track(left: np.ndarray, right: np.ndarray, num_points: int) -> np.ndarray:
    """tracks.py:247-252: midline, 10 origin points (x of the first midline point, y = 0) in front, degree-2 fit."""
    centre = (left + right) / 2
    origin = np.zeros((10, 2))
    origin[:, 0] = centre[0][0]
    return smooth_track_with_polyfit(np.concatenate([origin, centre], axis=0), num_points, 2)[0]


def make_instance(centreline: np.ndarray, index: int, horizon: int, offset_lat: float = 0.0, offset_psi: float = 0.0,
                  lookahead: float = 100.0, ds: float = 0.5) -> np.ndarray:
    """ego pose = centre-line point `index` moved `offset_lat` along the left normal, heading =
    tangent + `offset_psi`; the next `lookahead` metres resampled to H points (linear interpolation between map
    points) in the ego frame (x right, y forward), widths linspace(10, 6, H) (controller.py:264)."""
    M = centreline.shape[0]
    i = index % M
    ox, oy = centreline[i]
    nx, ny = centreline[(i + 1) % M]
    th = math.atan2(ny - oy, nx - ox)
    gx, gy = ox - offset_lat * math.sin(th), oy + offset_lat * math.cos(th)
    the = th + offset_psi
    out = np.empty((horizon, 3))
    widths = _linspace(10.0, 6.0, horizon)
    for k, s_m in enumerate(_linspace(0.0, lookahead, horizon)):
        s = s_m / ds
        fl = math.floor(s)
        frac = s - fl
        ia = (i + fl) % M
        ib = (ia + 1) % M
        qx = centreline[ia, 0] * (1.0 - frac) + centreline[ib, 0] * frac
        qy = centreline[ia, 1] * (1.0 - frac) + centreline[ib, 1] * frac
        rx, ry = qx - gx, qy - gy
        out[k] = (rx * math.sin(the) - ry * math.cos(the), rx * math.cos(the) + ry * math.sin(the), widths[k])
    return out
