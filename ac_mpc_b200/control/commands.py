"""Mirrors of the reference's command lookups (/root/reference/src/acmpc/control/commands.py) over the CUDA library:
`TemporalCommandSelector` (:8-38, what `ControlProcess.desired_state` uses, controller.py:100-112) and
`TemporalCommandInterpolator` (:41-99).  Same constructor (`controller` = any object with `control_cumtime` (n,) and
`control_inputs` (n,2)), same call signature, same quirks -- an elapsed time before the first command makes the
selector return the LAST command (index -1), and everything is computed in the dtype of the arrays (float32 for the
shared-memory views).  The arithmetic runs in `acmpc_select_commands_*_host`; there is no host implementation."""
from __future__ import annotations

import numpy as np

from ..solver import BatchedMPC


class _Lookup:
    _interpolate = False

    def __init__(self, controller, solver: BatchedMPC | None = None):
        self._controller = controller
        self._solver = solver

    def _native(self) -> BatchedMPC:
        if self._solver is None:
            mpc = getattr(self._controller, "model_predictive_controller", None)
            self._solver = mpc._batched() if mpc is not None else None
        if self._solver is None:
            raise RuntimeError("command lookup needs a BatchedMPC (pass solver=..., or a controller that owns a SpatialMPC)")
        return self._solver

    @property
    def _cum_time(self) -> np.ndarray:
        return self._controller.control_cumtime

    def _command_rows(self) -> np.ndarray:
        raise NotImplementedError

    def __call__(self, elapsed_time: float) -> np.ndarray:
        return self.get_command(elapsed_time)

    def get_command(self, elapsed_time: float) -> np.ndarray:
        ct = np.asarray(self._cum_time)
        cm = np.asarray(self._command_rows(), dtype=ct.dtype)
        out = self._native().select_commands(ct[None], cm[None], np.array([float(elapsed_time)]),
                                             interpolate=self._interpolate)
        return out[0]


class TemporalCommandSelector(_Lookup):
    """commands.py:8-38: the command whose time stamp is the latest one not after `elapsed_time`."""

    def _command_rows(self) -> np.ndarray:
        return self._controller.control_inputs          # (n,2), commands.py:16-18


class TemporalCommandInterpolator(_Lookup):
    """commands.py:41-99: linear interpolation between the two commands around `elapsed_time`.  The reference
    reads `control_inputs.T` as its (n,2) command table (:52-53; its test passes projected_control (2,n))."""

    _interpolate = True

    def _command_rows(self) -> np.ndarray:
        return np.asarray(self._controller.control_inputs).T
