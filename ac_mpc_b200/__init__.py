"""ac_mpc_b200 -- B200-native batched MPC step with ac-mpc's controller entry points.

Only the hot path of the reference is here: SpatialMPC.get_control as hand-written sm_100a CUDA
behind a C ABI (include/acmpc_b200.h).  See DESIGN.md / INTEGRATION.md.
"""
from . import tracks  # noqa: F401
from ._capi import STATUS_STRINGS, Config, default_config  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from .solver import BatchedMPC, config_from_reference, fp64_peak_tflops  # noqa: F401

__version__ = "0.1.0"
