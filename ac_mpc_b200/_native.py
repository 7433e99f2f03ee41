"""Process-wide default handle for the entry points that need no MPC configuration (map loading, track smoothing)."""
from __future__ import annotations

_default = None


def default_solver(device: int = 0):
    """One lazily created BatchedMPC on `device` (raises if the CUDA library or the GPU is missing: no CPU fallback)."""
    global _default
    if _default is None or _default.device != device:
        from . import _capi
        from .solver import BatchedMPC

        _default = BatchedMPC(_capi.default_config(), device=device)
    return _default
