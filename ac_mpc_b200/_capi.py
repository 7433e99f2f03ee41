"""ctypes binding of the C ABI in include/acmpc_b200.h (ac_mpc_b200/csrc/libacmpc_b200.so).

The library is CUDA-only.  If it cannot be loaded, or no sm_100 device is present when a solver is
created, this module raises: there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# ACMPC_B200_LIB: load another build of the same sources (instrumented experiments under tools/)
LIB_PATH = os.environ.get("ACMPC_B200_LIB") or os.path.join(CSRC, "libacmpc_b200.so")
_SOURCES = [os.path.join(CSRC, "acmpc_b200.cu"), os.path.join(CSRC, "mpc_warp.cuh"), os.path.join(CSRC, "simt.cuh"),
            os.path.join(CSRC, "map_profile.cuh"), os.path.join(CSRC, "publish.cuh"), os.path.join(CSRC, "track_prep.cuh"),
            os.path.join(CSRC, "model.cuh"),
            os.path.join(os.path.dirname(_HERE), "include", "acmpc_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# symbols declared in include/acmpc_b200.h
EXPORTED = [
    "acmpc_abi_version", "acmpc_default_config", "acmpc_create", "acmpc_destroy", "acmpc_last_error",
    "acmpc_warm_stride", "acmpc_solve_batch_device", "acmpc_solve_batch_host", "acmpc_last_launch_info",
    "acmpc_fp64_peak_tflops", "acmpc_set_profiling", "acmpc_collect_kernel_ms",
    "acmpc_construct_waypoints_host", "acmpc_map_speed_profile_host", "acmpc_track_speed_profile_host",
    "acmpc_reference_speeds_host",
    "acmpc_reference_paths_host", "acmpc_publish_host", "acmpc_select_commands_f32_host",
    "acmpc_select_commands_f64_host",
    "acmpc_remove_near_duplicates_host", "acmpc_smooth_tracks_polyfit_host", "acmpc_centre_tracks_host",
    "acmpc_extract_paths_device", "acmpc_extract_paths_host",
    "acmpc_speed_profile_batch_device", "acmpc_speed_profile_batch_host", "acmpc_t2s_host", "acmpc_s2t_host",
    "acmpc_linearise_host", "acmpc_remove_near_duplicates_cols_host",
    "acmpc_attach_completion", "acmpc_stream_wait_value32",
]

RC_NAMES = {0: "ACMPC_OK", 1: "ACMPC_ERR_INVALID", 2: "ACMPC_ERR_CUDA", 3: "ACMPC_ERR_NO_DEVICE"}
STATUS_STRINGS = {
    1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate", 4: "dual infeasible inaccurate",
    -2: "maximum iterations reached", -3: "primal infeasible", -4: "dual infeasible",
    -7: "problem non convex", -10: "unsolved",
}


class Config(C.Structure):
    """`acmpc_config`."""

    _fields_ = [
        ("horizon", C.c_int32), ("max_iter", C.c_int32),
        ("v_min", C.c_double), ("v_max", C.c_double), ("a_min", C.c_double), ("a_max", C.c_double),
        ("ay_max", C.c_double), ("ki_min", C.c_double), ("end_velocity", C.c_double),
        ("has_end_velocity", C.c_int32), ("reserved0", C.c_int32),
        ("step_cost", C.c_double * 3), ("r_term", C.c_double * 2), ("final_cost", C.c_double * 3),
        ("wheelbase", C.c_double), ("width", C.c_double), ("delta_max", C.c_double),
        ("input_v_min", C.c_double), ("input_v_max", C.c_double),
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("scaling", C.c_int32), ("check_termination", C.c_int32),
        ("adaptive_rho", C.c_int32), ("adaptive_rho_interval", C.c_int32),
        ("check_dualgap", C.c_int32), ("reserved1", C.c_int32),
    ]


class MapInfo(C.Structure):
    """`acmpc_map_info`."""

    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("rho_updates", C.c_int32), ("ctas", C.c_int32),
                ("pri_res", C.c_double), ("dua_res", C.c_double), ("obj_val", C.c_double), ("rho", C.c_double),
                ("kernel_ms", C.c_double)]


OUTPUT_FIELDS = ["controls", "prediction", "cum_time", "states", "v_ref", "cost", "pri_res", "dua_res",
                 "status", "status_speed", "iters", "rho_updates", "waypoints", "derived"]


class Outputs(C.Structure):
    """`acmpc_outputs`."""

    _fields_ = [(name, C.c_void_p) for name in OUTPUT_FIELDS]


def output_spec(H: int):
    """name -> (per-instance shape, numpy dtype string) in ABI order."""
    n = H - 1
    return {
        "controls": ((2, n), "float64"), "prediction": ((n, 2), "float64"), "cum_time": ((n,), "float64"),
        "states": ((H, 3), "float64"), "v_ref": ((n,), "float64"), "cost": ((), "float64"),
        "pri_res": ((), "float64"), "dua_res": ((), "float64"), "status": ((), "int32"),
        "status_speed": ((), "int32"), "iters": ((2,), "int32"), "rho_updates": ((2,), "int32"),
        "waypoints": ((7, n), "float64"), "derived": ((3, n - 1), "float64"),
    }


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree with nvcc (cross-compiles without a GPU)."""
    stale = force or not os.path.exists(LIB_PATH)
    if not stale:
        t = os.path.getmtime(LIB_PATH)
        stale = any(os.path.getmtime(s) > t for s in _SOURCES)
    if stale:
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, _SOURCES[0]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"ac_mpc_b200: CUDA extension {LIB_PATH} is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); there is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTED:
        if not hasattr(L, name):
            raise RuntimeError(f"ac_mpc_b200: {LIB_PATH} does not export {name}")
    dp, vp = C.POINTER(C.c_double), C.c_void_p
    L.acmpc_abi_version.restype = C.c_int32
    L.acmpc_default_config.argtypes = [C.POINTER(Config)]
    L.acmpc_default_config.restype = None
    L.acmpc_create.argtypes = [C.POINTER(Config), C.c_int32, C.POINTER(vp)]
    L.acmpc_create.restype = C.c_int32
    L.acmpc_destroy.argtypes = [vp]
    L.acmpc_destroy.restype = C.c_int32
    L.acmpc_last_error.argtypes = [vp]
    L.acmpc_last_error.restype = C.c_char_p
    L.acmpc_warm_stride.argtypes = [vp]
    L.acmpc_warm_stride.restype = C.c_int64
    L.acmpc_solve_batch_device.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_int32, vp, C.c_int32,
                                           C.POINTER(Outputs), vp]
    L.acmpc_solve_batch_device.restype = C.c_int32
    L.acmpc_solve_batch_host.argtypes = [vp, C.c_int32, dp, dp, dp, C.c_int32, C.c_int32, C.POINTER(Outputs)]
    L.acmpc_solve_batch_host.restype = C.c_int32
    L.acmpc_last_launch_info.argtypes = [vp] + [C.POINTER(C.c_int32)] * 4
    L.acmpc_last_launch_info.restype = C.c_int32
    L.acmpc_fp64_peak_tflops.argtypes = [C.c_int32, dp]
    L.acmpc_fp64_peak_tflops.restype = C.c_int32
    L.acmpc_set_profiling.argtypes = [vp, C.c_int32]
    L.acmpc_set_profiling.restype = C.c_int32
    L.acmpc_collect_kernel_ms.argtypes = [vp, dp, dp, C.POINTER(C.c_int32)]
    L.acmpc_collect_kernel_ms.restype = C.c_int32
    L.acmpc_construct_waypoints_host.argtypes = [vp, C.c_int32, dp, dp]
    L.acmpc_construct_waypoints_host.restype = C.c_int32
    L.acmpc_map_speed_profile_host.argtypes = [vp, C.c_int32, dp, C.c_double, C.c_double, C.c_double, C.c_int32, dp,
                                               C.POINTER(MapInfo)]
    L.acmpc_map_speed_profile_host.restype = C.c_int32
    L.acmpc_track_speed_profile_host.argtypes = [vp, C.c_int32, dp, C.c_double, C.c_double, C.c_double, C.c_int32,
                                                 dp, dp, C.POINTER(MapInfo)]
    L.acmpc_track_speed_profile_host.restype = C.c_int32
    L.acmpc_reference_speeds_host.argtypes = [vp, C.c_int32, dp, C.c_int32, C.c_int32, dp, dp]
    L.acmpc_reference_speeds_host.restype = C.c_int32
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.acmpc_reference_paths_host.argtypes = [vp, C.c_int32, C.c_int32, fp, dp]
    L.acmpc_reference_paths_host.restype = C.c_int32
    L.acmpc_publish_host.argtypes = [vp, C.c_int32, dp, dp, dp, fp, fp, fp]
    L.acmpc_publish_host.restype = C.c_int32
    L.acmpc_select_commands_f32_host.argtypes = [vp, C.c_int32, C.c_int32, fp, fp, dp, C.c_int32, fp, ip]
    L.acmpc_select_commands_f32_host.restype = C.c_int32
    L.acmpc_select_commands_f64_host.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, C.c_int32, dp, ip]
    L.acmpc_select_commands_f64_host.restype = C.c_int32
    L.acmpc_remove_near_duplicates_host.argtypes = [vp, C.c_int32, dp, C.c_double, dp, ip]
    L.acmpc_remove_near_duplicates_host.restype = C.c_int32
    L.acmpc_smooth_tracks_polyfit_host.argtypes = [vp, C.c_int32, ip, dp, C.c_int32, C.c_int32, dp, ip, ip]
    L.acmpc_smooth_tracks_polyfit_host.restype = C.c_int32
    L.acmpc_centre_tracks_host.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, C.c_int32, dp, ip]
    L.acmpc_centre_tracks_host.restype = C.c_int32
    L.acmpc_extract_paths_device.argtypes = [vp, C.c_int32, vp, C.c_int32, vp, vp, vp, C.c_double, C.c_double, vp, vp]
    L.acmpc_extract_paths_device.restype = C.c_int32
    L.acmpc_extract_paths_host.argtypes = [vp, C.c_int32, dp, C.c_int32, ip, dp, dp, C.c_double, C.c_double, dp]
    L.acmpc_extract_paths_host.restype = C.c_int32
    L.acmpc_speed_profile_batch_device.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, C.c_int32, C.c_double, vp, C.c_int32,
                                                   vp, vp, vp, vp, vp]
    L.acmpc_speed_profile_batch_device.restype = C.c_int32
    L.acmpc_speed_profile_batch_host.argtypes = [vp, C.c_int32, dp, dp, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                                 dp, ip, ip, ip]
    L.acmpc_speed_profile_batch_host.restype = C.c_int32
    L.acmpc_t2s_host.argtypes = [vp, C.c_int32, dp, dp, dp]
    L.acmpc_t2s_host.restype = C.c_int32
    L.acmpc_s2t_host.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, dp]
    L.acmpc_s2t_host.restype = C.c_int32
    L.acmpc_linearise_host.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, dp]
    L.acmpc_linearise_host.restype = C.c_int32
    L.acmpc_remove_near_duplicates_cols_host.argtypes = [vp, C.c_int32, C.c_int32, dp, C.c_double, dp, ip]
    L.acmpc_remove_near_duplicates_cols_host.restype = C.c_int32
    L.acmpc_attach_completion.argtypes = [vp, vp, C.c_uint32, vp, C.c_int32, C.c_uint32, vp, C.c_uint32]
    L.acmpc_attach_completion.restype = C.c_int32
    L.acmpc_stream_wait_value32.argtypes = [vp, vp, C.c_uint32, vp]
    L.acmpc_stream_wait_value32.restype = C.c_int32
    _lib = L
    return L


def default_config(**overrides) -> Config:
    cfg = Config()
    load().acmpc_default_config(C.byref(cfg))
    apply_overrides(cfg, overrides)
    return cfg


def apply_overrides(cfg: Config, overrides: dict) -> Config:
    for k, v in overrides.items():
        if not any(k == f[0] for f in Config._fields_):
            raise TypeError(f"unknown acmpc_config field {k!r}")
        if k in ("step_cost", "r_term", "final_cost"):
            arr = getattr(cfg, k)
            if len(v) != len(arr):
                raise ValueError(f"{k} needs {len(arr)} entries")
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(cfg, k, v)
    return cfg
