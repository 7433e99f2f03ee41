"""Synthetic centrelines and MPC problem instances.

The reference's map assets (data/maps/<track>*.npy, read by /root/reference/src/acmpc/utils/load.py:9-35)
are not in its tree, so the benchmark runs on closed, C2-smooth synthetic loops of the named lengths.
An instance is what ControlProcess._reference_path hands to get_control
(/root/reference/src/acmpc/control/controller.py:257-267): H points (x, y, width) in the ego frame
(x right, y forward, ego heading == pi/2, spatial_mpc.py:186-187), widths linspace(10, 6, H).
"""
from __future__ import annotations

import numpy as np

# nominal lap lengths in metres (general knowledge; not in the reference tree)
TRACK_LENGTHS = {
    "monza": 5793.0,
    "spa": 7004.0,
    "silverstone": 5891.0,
    "vallelunga": 4085.0,
    "yas_marina": 5554.0,
    "bathurst": 6213.0,
    "nordschleife": 20832.0,
}
TRACK_ORDER = list(TRACK_LENGTHS)

# racing.control blocks of /root/reference/configs/<track>.yaml:67-81 (all: horizon 50, v_max 84,
# a_max 1.0, r_term [1e-2, 10], final_cost [1, 0, 0.1])
RACING_CONTROL = {
    "monza": dict(v_min=8.0, a_min=-1.3, ay_max=5.5, ki_min=0.005, end_velocity=14.0, step_cost=[4e-3, 5e-2, 0.0]),
    "spa": dict(v_min=5.0, a_min=-1.0, ay_max=4.0, ki_min=0.003, end_velocity=20.0, step_cost=[1e-3, 0.0, 0.0]),
    "silverstone": dict(v_min=8.0, a_min=-1.0, ay_max=5.0, ki_min=0.003, end_velocity=20.0, step_cost=[2e-3, 5e-2, 0.0]),
    "vallelunga": dict(v_min=8.0, a_min=-1.0, ay_max=3.5, ki_min=0.003, end_velocity=None, step_cost=[8e-3, 5e-3, 0.0]),
    "yas_marina": dict(v_min=8.0, a_min=-1.0, ay_max=3.0, ki_min=0.0, end_velocity=14.0, step_cost=[2e-3, 2e-2, 0.0]),
    "nordschleife": dict(v_min=12.0, a_min=-1.0, ay_max=3.0, ki_min=0.0, end_velocity=14.0, step_cost=[2e-4, 0.0, 0.0]),
    "bathurst": dict(v_min=8.0, a_min=-1.0, ay_max=3.0, ki_min=0.0, end_velocity=14.0, step_cost=[1e-3, 2e-2, 0.0]),
}

# map_speed_profile_constraints blocks of /root/reference/configs/<track>.yaml (controller.py:89-91)
MAP_PROFILE = {
    "monza": dict(ay_max=7.0, a_min=-0.15), "spa": dict(ay_max=6.5, a_min=-0.15),
    "silverstone": dict(ay_max=8.0, a_min=-0.1), "vallelunga": dict(ay_max=5.0, a_min=-0.15),
    "yas_marina": dict(ay_max=2.0, a_min=-0.15), "bathurst": dict(ay_max=2.0, a_min=-0.15),
    "nordschleife": dict(ay_max=2.0, a_min=-0.15),
}
ROAD_WIDTH_MAP = 9.5   # agent.py:288


def map_track(centreline: np.ndarray) -> np.ndarray:
    """(M,3) input of Controller.compute_track_speed_profile, as agent.py:287-296 builds it."""
    return np.column_stack([centreline[:, 0], centreline[:, 1], np.full(len(centreline), ROAD_WIDTH_MAP)])


def racing_config(track: str = "monza", horizon: int = 50) -> dict:
    """The dict `build_mpc` receives (controller.py:19-29) for a track's racing.control block."""
    r = RACING_CONTROL[track]
    return {
        "horizon": horizon,
        "speed_profile_constraints": {
            "v_min": r["v_min"], "v_max": 84.0, "a_min": r["a_min"], "a_max": 1.0,
            "ay_max": r["ay_max"], "ki_min": r["ki_min"], "end_velocity": r["end_velocity"],
        },
        "step_cost": list(r["step_cost"]),
        "r_term": [1e-2, 10.0],
        "final_cost": [1.0, 0.0, 0.1],
    }


def _resample_closed(xy: np.ndarray, ds: float) -> np.ndarray:
    seg = np.linalg.norm(np.diff(np.vstack([xy, xy[:1]]), axis=0), axis=1)
    s = np.concatenate([[0.0], np.cumsum(seg)])
    total = s[-1]
    m = int(np.floor(total / ds))
    t = np.arange(m) * ds
    ext = np.vstack([xy, xy[:1]])
    return np.stack([np.interp(t, s, ext[:, 0]), np.interp(t, s, ext[:, 1])], axis=1)


def synthetic_centreline(track: str, ds: float = 0.5, min_radius: float = 15.0) -> np.ndarray:
    """Closed band-limited Fourier loop, seed = track index, rescaled to the named length and
    resampled every `ds` metres (map density, mapping/map_maker.py:203).  Returns (M, 2) float64."""
    idx = TRACK_ORDER.index(track)
    length = TRACK_LENGTHS[track]
    rng = np.random.default_rng(1000 + idx)
    n_harm = max(9, int(round(length / 350.0)))
    amp = rng.uniform(0.05, 0.30, n_harm) / np.arange(2, n_harm + 2) ** 1.0
    phase = rng.uniform(0, 2 * np.pi, n_harm)
    theta = np.linspace(0.0, 2 * np.pi, 200001)[:-1]
    for _ in range(80):
        r = 1.0 + sum(a * np.cos((k + 2) * theta + p) for k, (a, p) in enumerate(zip(amp, phase)))
        xy = np.stack([r * np.cos(theta), r * np.sin(theta)], axis=1)
        seg = np.linalg.norm(np.diff(np.vstack([xy, xy[:1]]), axis=0), axis=1)
        xy = xy * (length / seg.sum())
        d1 = np.gradient(xy, axis=0)
        d2 = np.gradient(d1, axis=0)
        curv = np.abs(d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]) / (np.linalg.norm(d1, axis=1) ** 3 + 1e-300)
        if 1.0 / curv.max() >= min_radius:
            break
        amp *= 0.93
    return _resample_closed(xy, ds)


def centreline(track: str, map_dir: str | None = None, ds: float = 0.5) -> np.ndarray:
    """The centre line the benchmark and the sweeps run on: the downloaded map (data/maps/<track>*.npy in the
    reference's on-disk format, utils/load.py:9-35; resampled to `ds` metres) when `map_dir` holds one, else the
    synthetic loop of the named length."""
    if map_dir:
        from .utils import load

        path = load.find_map(track, map_dir)
        if path:
            return _resample_closed(load.track_map(path)["centre"][:, :2], ds)
    return synthetic_centreline(track, ds)


def make_instances(centreline: np.ndarray, indices, horizon: int, offset_lat=None, offset_psi=None,
                   lookahead: float = 100.0, ds: float = 0.5) -> np.ndarray:
    """(B, H, 3) get_control inputs: ego pose = centreline point i displaced `offset_lat` along the
    left normal with heading tangent + `offset_psi`; the next `lookahead` metres of centreline are
    resampled to H points equally spaced in arc length and expressed in the ego frame."""
    indices = np.asarray(indices, dtype=np.int64)
    B, H, M = indices.shape[0], horizon, centreline.shape[0]
    offset_lat = np.zeros(B) if offset_lat is None else np.asarray(offset_lat, float)
    offset_psi = np.zeros(B) if offset_psi is None else np.asarray(offset_psi, float)
    s = np.linspace(0.0, lookahead, H) / ds                       # fractional sample offsets
    i0 = np.floor(s).astype(np.int64)
    frac = s - i0
    ia = (indices[:, None] + i0[None, :]) % M
    ib = (ia + 1) % M
    pts = centreline[ia] * (1.0 - frac)[None, :, None] + centreline[ib] * frac[None, :, None]
    tang = centreline[(indices + 1) % M] - centreline[indices]
    th = np.arctan2(tang[:, 1], tang[:, 0])
    left = np.stack([-np.sin(th), np.cos(th)], axis=1)
    origin = centreline[indices] + offset_lat[:, None] * left
    the = th + offset_psi
    fwd = np.stack([np.cos(the), np.sin(the)], axis=1)
    right = np.stack([np.sin(the), -np.cos(the)], axis=1)
    rel = pts - origin[:, None, :]
    out = np.empty((B, H, 3))
    out[:, :, 0] = np.einsum("bhd,bd->bh", rel, right)
    out[:, :, 1] = np.einsum("bhd,bd->bh", rel, fwd)
    out[:, :, 2] = np.linspace(10.0, 6.0, H)[None, :]
    return out


def perturbed_batch(track: str, B: int, horizon: int = 50, seed: int = 1, centreline=None):
    """BASELINE.json configs[1]-style batch: waypoint index uniform over the lap,
    offset_lat ~ U(-2, 2) m, offset_psi ~ U(-0.1, 0.1) rad, v_max ~ U(20, 84) m/s."""
    cl = synthetic_centreline(track) if centreline is None else centreline
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, cl.shape[0], B)
    lat = rng.uniform(-2.0, 2.0, B)
    psi = rng.uniform(-0.1, 0.1, B)
    vmax = rng.uniform(20.0, 84.0, B)
    return make_instances(cl, idx, horizon, lat, psi), vmax
