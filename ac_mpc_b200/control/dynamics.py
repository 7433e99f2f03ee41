"""Drop-in mirror of the reference's `SpatialBicycleModel`
(/root/reference/src/acmpc/control/dynamics.py:9-103): same constructor, attributes and the three methods
`t2s`, `s2t`, `linearise`.

Inside a `get_control` step the three transforms are fused into the control kernel
(ac_mpc_b200/csrc/mpc_warp.cuh: ControlQP::setup / control_instance).  Called stand-alone, as the reference's
object API allows, they run as small batched kernels behind the C ABI (acmpc_t2s_host / acmpc_s2t_host /
acmpc_linearise_host, ac_mpc_b200/csrc/model.cuh); the `*_batch` variants take B instances per launch.
There is no host-side arithmetic here: without a B200 the calls raise."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .._native import default_solver


def _rows(reference_path) -> np.ndarray:
    """(7,n) array of a ReferencePath (this package's or the reference's, paths.py:4-72)."""
    rows = reference_path.as_array() if hasattr(reference_path, "as_array") else reference_path._reference_path
    return np.ascontiguousarray(rows, dtype=np.float64)


class SpatialBicycleModel:
    def __init__(self, vehicle_data, velocity_limits: Dict, device: int = 0):
        self.length = float(vehicle_data.vehicle_data.wheelbase)
        self.width = float(vehicle_data.vehicle_data.width)
        self.delta_max = float(vehicle_data.max_steering_angle())
        self.margin = self.width / 2
        self.min_velocity = velocity_limits["min"]
        self.max_velocity = velocity_limits["max"]
        kappa_max = np.tan(self.delta_max) / self.length
        self.min_u = np.array([self.min_velocity, -kappa_max])
        self.max_u = np.array([self.max_velocity, kappa_max])
        self._eps = 1e-12
        self._device = int(device)

    # handy for callers that built the model themselves (build_mpc passes a SteeringGeometry-like)
    @property
    def vehicle_data(self):
        from types import SimpleNamespace

        return SimpleNamespace(wheelbase=self.length, width=self.width)

    def max_steering_angle(self) -> float:
        return self.delta_max

    def _solver(self):
        return default_solver(self._device)

    # -- dynamics.py:23-40 -------------------------------------------------------------------------------------------
    def t2s(self, reference_waypoint: np.ndarray, reference_state: np.ndarray) -> np.ndarray:
        """(x, y, psi) of the reference waypoint and of the vehicle -> spatial state (e_y, e_psi, t = 0)."""
        w = np.asarray(reference_waypoint, dtype=np.float64).reshape(1, 3)
        x = np.asarray(reference_state, dtype=np.float64).reshape(1, 3)
        return self._solver().t2s(w, x)[0]

    def t2s_batch(self, reference_waypoints, reference_states) -> np.ndarray:
        return self._solver().t2s(reference_waypoints, reference_states)

    # -- dynamics.py:42-63 -------------------------------------------------------------------------------------------
    def s2t(self, reference_waypoints, reference_states: np.ndarray) -> np.ndarray:
        """ReferencePath + spatial states (n,3) -> np.array([xs, ys, psis]) (3,n)."""
        x = np.ascontiguousarray(reference_states, dtype=np.float64)
        return self._solver().s2t(_rows(reference_waypoints)[None], x[None, :, :3])[0]

    # -- dynamics.py:65-103 ------------------------------------------------------------------------------------------
    def linearise(self, reference_path) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """f (n,3), A (n,3,3), B (n,3,2) around the path's distances, kappas and velocities."""
        f, A, B = self._solver().linearise(_rows(reference_path)[None])
        return f[0], A[0], B[0]
