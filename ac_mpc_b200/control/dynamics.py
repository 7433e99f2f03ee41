"""Vehicle-model parameter object with the reference's attribute surface
(/root/reference/src/acmpc/control/dynamics.py:9-21).

In the reference this class also linearises the model on the host (dynamics.py:65-103); here the
linearisation, the t2s initial-state transform and the s2t rollout all happen inside the CUDA kernel
(ac_mpc_b200/csrc/mpc_warp.cuh: ControlQP::setup / control_instance), so the object only carries the
constants that become fields of `acmpc_config`."""
from __future__ import annotations

from typing import Dict

import numpy as np


class SpatialBicycleModel:
    def __init__(self, vehicle_data, velocity_limits: Dict):
        self.length = float(vehicle_data.vehicle_data.wheelbase)
        self.width = float(vehicle_data.vehicle_data.width)
        self.delta_max = float(vehicle_data.max_steering_angle())
        self.margin = self.width / 2
        self.min_velocity = velocity_limits["min"]
        self.max_velocity = velocity_limits["max"]
        kappa_max = np.tan(self.delta_max) / self.length
        self.min_u = np.array([self.min_velocity, -kappa_max])
        self.max_u = np.array([self.max_velocity, kappa_max])

    # handy for callers that built the model themselves (build_mpc passes a SteeringGeometry-like)
    @property
    def vehicle_data(self):
        from types import SimpleNamespace

        return SimpleNamespace(wheelbase=self.length, width=self.width)

    def max_steering_angle(self) -> float:
        return self.delta_max
