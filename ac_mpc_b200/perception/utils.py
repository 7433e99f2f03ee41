"""Drop-in for `acmpc.perception.utils.smooth_track_with_polyfit` (/root/reference/src/acmpc/perception/utils.py:107-119)
and `TrackLimitPerception._calculate_centre_track` (perception/tracks.py:247-252), batched on the GPU."""
from __future__ import annotations

import numpy as np

from .._native import default_solver


def smooth_track_with_polyfit(track, num_points: int, degree: int = 3, solver=None) -> np.ndarray:
    """One track (m,2) -> (num_points,2), the reference's signature."""
    return (solver or default_solver()).smooth_tracks_with_polyfit([track], num_points, degree)[0]


def smooth_tracks_with_polyfit(tracks, num_points: int, degree: int = 3, solver=None) -> np.ndarray:
    """A list of ragged tracks in one launch -> (B,num_points,2)."""
    return (solver or default_solver()).smooth_tracks_with_polyfit(tracks, num_points, degree)


def calculate_centre_track(tracks: dict, n_polyfit_points: int | None = None, solver=None) -> np.ndarray:
    """tracks.py:247-252 on {"left": (N,2), "right": (N,2)} -> (n_polyfit_points,2)."""
    left, right = np.asarray(tracks["left"], float), np.asarray(tracks["right"], float)
    return (solver or default_solver()).centre_tracks(left[None], right[None], n_polyfit_points)[0]
