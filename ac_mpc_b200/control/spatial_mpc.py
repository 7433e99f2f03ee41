"""Drop-in mirror of the reference's `SpatialMPC` (/root/reference/src/acmpc/control/spatial_mpc.py:20-217)
on top of the CUDA library.  Same constructor arguments, same `get_control` signature, and results
delivered as the same mutated attributes the caller reads (controller.py:274-280):
`projected_control (2,n)`, `current_prediction (n,2)`, `cum_time (n,)`, `times`, `accelerations`,
`steer_rates`, `reference_path`, `speed_profile`, `infeasibility_counter`.

`get_control` is the B = 1 case of `get_control_batch`; both run the fused sm_100a kernel.  There is
no host-side solver: without a B200 the call raises.
"""
from __future__ import annotations

import copy
from typing import Dict, Optional

import numpy as np

from .. import _capi
from ..solver import BatchedMPC, config_from_reference
from .paths import ReferencePath

try:  # the reference logs through loguru; fall back to logging if it is absent
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("ac_mpc_b200")

MAX_SOLVER_ITERATIONS_MAP = 40000
MAX_SOLVER_ITERATIONS = 4000


class SpatialMPC:
    def __init__(self, config: Dict, model, device: int = 0, **osqp_overrides):
        self.MPC_horizon = config["horizon"]
        self.cum_time = np.zeros((1))
        self.model = model
        self.nx = 3
        self.nu = 2
        # live dict, mutated by the caller before every step (controller.py:241-243)
        self.speed_profile_constraints = config["speed_profile_constraints"]
        self.ay_max = self.speed_profile_constraints["ay_max"]
        self.delta_max = model.delta_max
        self.current_prediction = None
        self.infeasibility_counter = 0
        self.projected_control = np.zeros((self.nu, self.MPC_horizon))
        self._config = copy.deepcopy(config)
        self._config["max_iterations"] = MAX_SOLVER_ITERATIONS
        self._device = device
        self._osqp_overrides = dict(osqp_overrides)
        self._solver: Optional[BatchedMPC] = None      # created lazily (fork safety, controller.py:293)
        self._solver_key = None
        self._out1 = self._out1_solver = None
        self.last_info: Dict = {}

    # ------------------------------------------------------------------------------------------
    def _batched(self) -> BatchedMPC:
        """(Re)build the native handle when a constraint other than v_max was edited in place."""
        spc = self.speed_profile_constraints
        key = tuple((k, spc.get(k)) for k in ("v_min", "a_min", "a_max", "ay_max", "ki_min", "end_velocity"))
        if self._solver is None or key != self._solver_key:
            cfg_dict = dict(self._config)
            cfg_dict["speed_profile_constraints"] = dict(spc)
            cfg = config_from_reference(cfg_dict, self.model, **self._osqp_overrides)
            # input bounds keep the model's build-time limits (dynamics.py:15-20)
            cfg.input_v_min, cfg.input_v_max = float(self.model.min_velocity), float(self.model.max_velocity)
            if self._solver is not None:
                self._solver.close()
            self._solver = BatchedMPC(cfg, self._device)
            self._solver_key = key
        return self._solver

    def get_control_batch(self, reference_paths, offsets=None, v_max=None, is_localised: bool = False,
                          fields=None, out=None, keep_warm: bool = False) -> Dict[str, np.ndarray]:
        """B independent MPC steps: reference_paths (B,H,3), offsets (B,), v_max (B,) -> dict of arrays
        (fields of `acmpc_outputs`).  Cold start per instance unless keep_warm (then slot b of consecutive
        calls is one persistent solver object, like the reference's).  `out`: optional preallocated (e.g.
        pinned) host arrays from `BatchedMPC.alloc_host_outputs`."""
        solver = self._batched()
        paths = np.ascontiguousarray(reference_paths, dtype=np.float64)
        B = paths.shape[0]
        if v_max is None:
            v_max = np.full(B, float(self.speed_profile_constraints["v_max"]))
        return solver.solve_host(paths, offsets, v_max, is_localised, out=out, fields=fields, keep_warm=keep_warm)

    def get_control(self, reference_path: np.ndarray, is_localised: bool = False, offset: float = 0.0):
        """One MPC step; signature and side effects of spatial_mpc.py:170-217."""
        path = np.ascontiguousarray(reference_path, dtype=np.float64)[None]
        # the reference keeps its OSQP objects between calls (warm start, carried rho): so does the handle
        # one set of output arrays per object, reused by every call (the attributes below are copies)
        solver = self._batched()
        if self._out1 is None or self._out1_solver is not solver:
            self._out1, self._out1_solver = solver.alloc_host_outputs(1), solver
            self._off1, self._vmax1 = np.zeros(1), np.zeros(1)
        self._off1[0] = float(offset)
        self._vmax1[0] = float(self.speed_profile_constraints["v_max"])
        out = solver.solve_host(path, self._off1, self._vmax1, is_localised, out=self._out1, keep_warm=True)
        n = self.MPC_horizon - 1
        status, status_speed = int(out["status"][0]), int(out["status_speed"][0])
        self.last_info = dict(status=_capi.STATUS_STRINGS.get(status, str(status)),
                              status_speed=_capi.STATUS_STRINGS.get(status_speed, str(status_speed)),
                              iters=out["iters"][0].tolist(), rho_updates=out["rho_updates"][0].tolist(),
                              cost=float(out["cost"][0]), pri_res=float(out["pri_res"][0]),
                              dua_res=float(out["dua_res"][0]))
        waypoints = ReferencePath(n, out["waypoints"][0].copy())
        if status_speed == 1:
            self.speed_profile = waypoints.velocities.copy()
        else:
            failed = np.hstack([waypoints.xs, waypoints.ys])
            logger.warning("Infeasible problem! reference path:\n" + f"{failed}")
        if status == 1:
            self.projected_control = out["controls"][0].copy()
            self.current_prediction = out["prediction"][0].copy()
            self.reference_path = waypoints
            self.cum_time = out["cum_time"][0].copy()
            # times, accelerations (sic: diff of the e_y column, spatial_mpc.py:210), steer_rates: the kernel's
            # `derived` rows
            self.times, self.accelerations, self.steer_rates = (out["derived"][0][k].copy() for k in range(3))
            self.infeasibility_counter = 0
        else:
            logger.warning(f"Infeasible problem! Failed {self.infeasibility_counter} time(s).")
            self.infeasibility_counter += 1

    # -- the whole-track profile (spatial_mpc.py:60-87, :125-154) ------
    def construct_waypoints(self, waypoint_coordinates) -> ReferencePath:
        """(M,3) array of (x, y, width) -> ReferencePath of M-1 waypoints (spatial_mpc.py:125-154), on the GPU."""
        rows = self._batched().construct_waypoints(np.asarray(waypoint_coordinates, dtype=np.float64))
        return ReferencePath(rows.shape[1], rows)

    def compute_map_speed_profile(self, reference_path: ReferencePath, ay_max: float, a_min: float) -> ReferencePath:
        """One cooperative launch solves the whole-track QP (max_iter = 40000).  Velocities are assigned only when
        OSQP's status is "solved" (spatial_mpc.py:115-123), otherwise a warning is logged."""
        rows = reference_path.as_array()
        x, info = self._batched().map_speed_profile(rows, float(self.speed_profile_constraints["v_max"]), ay_max,
                                                    a_min, MAX_SOLVER_ITERATIONS_MAP)
        self.last_map_info = info
        if info["status"] == 1:
            self.speed_profile = x
        else:
            failed = np.hstack([reference_path.xs, reference_path.ys])
            logger.warning("Infeasible problem! reference path:\n" + f"{failed}")
        return reference_path

    def compute_speed_profile(self, reference_path: ReferencePath, is_localised: bool = False,
                              end_vel=None) -> ReferencePath:
        """spatial_mpc.py:89-123: the speed-profile QP alone (the speed kernel without the control kernel) on a
        ReferencePath.  Like the reference it runs on the object's persistent solver (the localised one when
        `is_localised`), which get_control shares: warm start and carried rho continue across both entry points.
        Velocities are assigned only when OSQP's status is "solved"."""
        rows = reference_path.as_array()
        if rows.shape != (7, self.MPC_horizon - 1):
            raise ValueError(f"reference path must have {self.MPC_horizon - 1} waypoints (the solver's horizon)")
        buf = np.ascontiguousarray(rows, dtype=np.float64)[None].copy()
        res = self._batched().speed_profile_host(buf, np.array([float(self.speed_profile_constraints["v_max"])]),
                                                 is_localised, end_vel, keep_warm=True)
        status = int(res["status"][0])
        self.last_speed_info = dict(status=_capi.STATUS_STRINGS.get(status, str(status)), iters=int(res["iters"][0]),
                                    rho_updates=int(res["rho_updates"][0]))
        if status == 1:
            reference_path.velocities = res["x"][0]
            self.speed_profile = res["x"][0].copy()
        else:
            failed = np.hstack([reference_path.xs, reference_path.ys])
            logger.warning("Infeasible problem! reference path:\n" + f"{failed}")
        return reference_path

    def update_prediction(self, spatial_state_prediction: np.ndarray, reference_path: ReferencePath) -> np.ndarray:
        """spatial_mpc.py:156-168: predicted spatial states (n,3) -> predicted (x, y) locations (n,2) (s2t rollout)."""
        x = np.ascontiguousarray(spatial_state_prediction, dtype=np.float64)
        rows = np.ascontiguousarray(reference_path.as_array(), dtype=np.float64)
        return self._batched().s2t(rows[None], x[None, :, :3], prediction=True)[0]
