"""TEST-ONLY: ctypes loader of the one-lane CPU emulation of the kernel body (see emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import port

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libacmpc_emul.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "emul.cpp"), os.path.join(_ROOT, "ac_mpc_b200", "csrc", "mpc_warp.cuh"), os.path.join(_ROOT, "ac_mpc_b200", "csrc", "simt.cuh"),
                os.path.join(_ROOT, "include", "acmpc_b200.h")]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-o", _SO, srcs[0]],
                           check=True, capture_output=True)
        _lib = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        _lib.acmpc_emul_solve_batch.argtypes = [C.POINTER(port.Config), C.c_int, dp, dp, dp, C.c_int,
                                                C.POINTER(port.Outputs), dp, C.c_int]
    return _lib


def warm_buffer(cfg, B):
    """Zeroed warm-start records for B instances (the emulation's counterpart of acmpc_warm_stride)."""
    return np.zeros((B, lib().acmpc_emul_warm_doubles(int(cfg.horizon))), dtype=np.float64)


def solve_batch(cfg, paths, offsets=None, vmax=None, is_localised=False, warm=None, use_warm=True):
    paths = np.ascontiguousarray(paths, dtype=np.float64)
    B, H, _ = paths.shape
    offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    vmax = None if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
    arrs, o = port.alloc_outputs(B, H)
    rc = lib().acmpc_emul_solve_batch(C.byref(cfg), B, port._dptr(paths), port._dptr(offsets), port._dptr(vmax),
                                      int(bool(is_localised)), C.byref(o), port._dptr(warm), int(bool(use_warm)))
    assert rc == 0
    return arrs


def speed_profile(cfg, waypoints, vmax=None, is_localised=False, end_vel=None, warm=None, use_warm=True):
    """speed_instance's stand-alone mode (SpatialMPC.compute_speed_profile) on (B,7,n) rows, in place."""
    import copy

    w = waypoints
    assert w.dtype == np.float64 and w.flags.c_contiguous and w.ndim == 3
    B, n = w.shape[0], w.shape[2]
    cfg = copy.copy(cfg)
    cfg.has_end_velocity, cfg.end_velocity = int(end_vel is not None), float(0.0 if end_vel is None else end_vel)
    vmax = None if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
    x = np.zeros((B, n))
    st, it, ru = (np.zeros(B, np.intc) for _ in range(3))
    ip = C.POINTER(C.c_int)
    L = lib()
    L.acmpc_emul_speed_profile.argtypes = [C.POINTER(port.Config), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.c_int, C.POINTER(C.c_double), ip, ip, ip, C.POINTER(C.c_double), C.c_int]
    L.acmpc_emul_speed_profile(C.byref(cfg), B, port._dptr(w), port._dptr(vmax), int(bool(is_localised)), port._dptr(x),
                               st.ctypes.data_as(ip), it.ctypes.data_as(ip), ru.ctypes.data_as(ip), port._dptr(warm),
                               int(bool(use_warm)))
    return dict(x=x, status=st, iters=it, rho_updates=ru)
