"""Drop-in for the map part of `acmpc.utils.load` (/root/reference/src/acmpc/utils/load.py:9-65): the on-disk map
format ({"outside_track", "inside_track", "centre_track"} pickled into a .npy, or the JSON variant with "Outside" /
"Inside" / "Centre") and the near-duplicate removal, which runs on the GPU (acmpc_remove_near_duplicates_host).
`save_track_map` writes the same format, so synthetic centre lines and downloaded maps are interchangeable."""
from __future__ import annotations

import glob
import json as _json
import os
from typing import Dict, Optional

import numpy as np

from .._native import default_solver


def remove_near_duplicate_points(track: np.ndarray, solver=None) -> np.ndarray:
    """load.py:30-35."""
    return (solver or default_solver()).remove_near_duplicate_points(track, 0.0001)


def npy(filepath: str) -> Dict:
    """load.py:61-62."""
    return np.load(filepath, allow_pickle=True).item()


def json(filepath: str) -> Dict:
    """load.py:54-57."""
    with open(filepath) as file:
        return _json.load(file)


def _load_json_track(path: str) -> Dict:
    """load.py:38-45."""
    data = json(path)
    return {"centre_track": np.array(data["Centre"]), "outside_track": np.array(data["Outside"]),
            "inside_track": np.array(data["Inside"])}


EXTENSION_TO_METHOD = {"npy": npy, "json": _load_json_track}


def track_map(path: str, solver=None) -> Dict:
    """load.py:9-27: {"left", "right", "centre"} with near-duplicate points removed."""
    track_dict = EXTENSION_TO_METHOD[path.split(".")[-1]](path)
    tracks = {"left": track_dict["outside_track"], "right": track_dict["inside_track"],
              "centre": track_dict["centre_track"]}
    return {k: remove_near_duplicate_points(np.asarray(v, dtype=np.float64), solver) for k, v in tracks.items()}


def save_track_map(path: str, centre: np.ndarray, outside: np.ndarray, inside: np.ndarray) -> None:
    """Write a map the reference's `load.track_map` (and this one) reads."""
    np.save(path, {"outside_track": np.asarray(outside), "inside_track": np.asarray(inside),
                   "centre_track": np.asarray(centre)}, allow_pickle=True)


def find_map(track: str, map_dir: str) -> Optional[str]:
    """data/maps/<track>*.npy as scripts/download_assets.sh:42-48 lays the assets out; None if absent."""
    hits = sorted(glob.glob(os.path.join(map_dir, f"{track}*.npy")) + glob.glob(os.path.join(map_dir, track, "*.npy")))
    return hits[0] if hits else None
