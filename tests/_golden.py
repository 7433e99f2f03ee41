"""Loader for tests/golden/mpc_golden.npz (made by tests/golden/make_golden.py)."""
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_golden.npz")
_CACHE = {}

FIXTURE_CONFIG = dict(horizon=100, v_min=12.0, v_max=84.0, a_min=-1.0, a_max=1.0, ay_max=5.5, ki_min=0.005,
                      end_velocity=14.0, has_end_velocity=1, step_cost=[2e-3, 5e-2, 0.0], r_term=[1e-2, 10.0],
                      final_cost=[1.0, 0.0, 0.1], input_v_min=12.0, input_v_max=84.0)


def load():
    if not _CACHE:
        with np.load(_PATH) as z:
            for k in z.files:
                if k.startswith("meta/"):
                    continue
                g, name = k.split("/")
                _CACHE.setdefault(g, {})[name] = z[k]
    return _CACHE


def racing_kwargs(track, horizon=50):
    """Config-field kwargs (acmpc_config / oracle.port.Config) for a track's racing.control block."""
    from ac_mpc_b200 import tracks

    r = tracks.RACING_CONTROL[track]
    return dict(horizon=horizon, v_min=r["v_min"], v_max=84.0, a_min=r["a_min"], a_max=1.0, ay_max=r["ay_max"],
                ki_min=r["ki_min"], end_velocity=0.0 if r["end_velocity"] is None else r["end_velocity"],
                has_end_velocity=0 if r["end_velocity"] is None else 1, step_cost=r["step_cost"],
                r_term=[1e-2, 10.0], final_cost=[1.0, 0.0, 0.1], input_v_min=r["v_min"], input_v_max=84.0)


def groups():
    """(group name, config kwargs) for every cold-start golden group."""
    from ac_mpc_b200 import tracks

    out = [("fixture_cold", FIXTURE_CONFIG)]
    out += [(f"racing_{t}", racing_kwargs(t)) for t in tracks.TRACK_ORDER]
    out += [(f"spa_h{H}", racing_kwargs("spa", H)) for H in (20, 40, 80)]
    return out
