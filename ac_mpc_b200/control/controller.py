"""`build_mpc` with the reference's signature (/root/reference/src/acmpc/control/controller.py:19-29)."""
from __future__ import annotations

from typing import Dict

from .dynamics import SpatialBicycleModel
from .spatial_mpc import SpatialMPC


def build_mpc(control_config: Dict, vehicle_data, device: int = 0, **osqp_overrides) -> SpatialMPC:
    velocity_limits = {
        "max": control_config["speed_profile_constraints"]["v_max"],
        "min": control_config["speed_profile_constraints"]["v_min"],
    }
    model = SpatialBicycleModel(vehicle_data, velocity_limits, device=device)
    return SpatialMPC(control_config, model, device=device, **osqp_overrides)
