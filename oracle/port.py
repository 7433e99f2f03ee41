"""ctypes bindings of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  PARITY UNPINNED (see osqp_port.h).

Two layers are exposed:
  * `PortOSQP`   -- the OSQP restatement for a general sparse QP (osqp_port.c); the shim package
                    oracle/shim/osqp wraps it in the `osqp.OSQP` object API the reference calls
                    (/root/reference/src/acmpc/control/solvers/control.py:88-106).
  * `PortMPC`    -- one SpatialMPC.get_control (acmpc_port.c), single instance (optionally warm)
                    or a multi-threaded cold-start batch.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

STATUS_STRINGS = {
    1: "solved",
    2: "solved inaccurate",
    3: "primal infeasible inaccurate",
    4: "dual infeasible inaccurate",
    -2: "maximum iterations reached",
    -3: "primal infeasible",
    -4: "dual infeasible",
    -7: "problem non convex",
    -10: "unsolved",
}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only) and return the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("osqp_port.c", "acmpc_port.c", "osqp_port.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "acmpc_b200.h"))
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


class Settings(C.Structure):
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("scaling", C.c_int), ("max_iter", C.c_int), ("check_termination", C.c_int),
        ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
        ("warm_start", C.c_int), ("scaled_termination", C.c_int), ("check_dualgap", C.c_int),
    ]


class Info(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("iter", C.c_int), ("rho_updates", C.c_int),
        ("obj_val", C.c_double), ("pri_res", C.c_double), ("dua_res", C.c_double),
        ("rho_estimate", C.c_double), ("duality_gap", C.c_double),
    ]


class Config(C.Structure):
    """Mirror of `acmpc_config` (include/acmpc_b200.h)."""

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


class Outputs(C.Structure):
    """Mirror of `acmpc_outputs`."""

    _fields_ = [
        ("controls", C.c_void_p), ("prediction", C.c_void_p), ("cum_time", C.c_void_p),
        ("states", C.c_void_p), ("v_ref", C.c_void_p), ("cost", C.c_void_p),
        ("pri_res", C.c_void_p), ("dua_res", C.c_void_p), ("status", C.c_void_p),
        ("status_speed", C.c_void_p), ("iters", C.c_void_p), ("rho_updates", C.c_void_p),
        ("waypoints", C.c_void_p), ("derived", C.c_void_p),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.opq_default_settings.argtypes = [C.POINTER(Settings)]
        L.opq_setup.restype = C.c_void_p
        L.opq_setup.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, ip, ip, dp, dp, dp, C.POINTER(Settings)]
        L.opq_free.argtypes = [C.c_void_p]
        L.opq_update.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.opq_update.restype = C.c_int
        L.opq_cold_start.argtypes = [C.c_void_p]
        L.opq_warm_start.argtypes = [C.c_void_p, dp, dp]
        L.opq_solve.argtypes = [C.c_void_p, dp, dp, C.POINTER(Info)]
        L.opq_solve.restype = C.c_int
        L.opq_get_vec.restype = dp
        L.opq_get_vec.argtypes = [C.c_void_p, C.c_char_p, ip]
        L.opq_get_scalar.restype = C.c_double
        L.opq_get_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.acmpc_port_create.restype = C.c_void_p
        L.acmpc_port_create.argtypes = [C.POINTER(Config)]
        L.acmpc_port_destroy.argtypes = [C.c_void_p]
        L.acmpc_port_step.argtypes = [C.c_void_p, dp, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(Outputs)]
        L.acmpc_port_step.restype = C.c_int
        L.acmpc_port_solve_batch.argtypes = [C.POINTER(Config), C.c_int, dp, dp, dp, C.c_int, C.c_int, C.POINTER(Outputs)]
        L.acmpc_port_solve_batch.restype = C.c_int
        L.acmpc_port_default_config.argtypes = [C.POINTER(Config)]
        L.acmpc_port_speed_profile.argtypes = [C.c_void_p, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, dp,
                                               ip, ip]
        L.acmpc_port_speed_profile.restype = C.c_int
        L.acmpc_port_waypoints.restype = dp
        L.acmpc_port_waypoints.argtypes = [C.c_void_p]
        L.acmpc_port_get_qp.argtypes = [C.c_void_p, C.c_int, ip, ip, C.POINTER(ip), C.POINTER(ip),
                                        C.POINTER(dp), C.POINTER(dp), C.POINTER(dp), C.POINTER(dp), C.POINTER(dp)]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def default_settings(**kw) -> Settings:
    s = Settings()
    lib().opq_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise TypeError(f"unknown OSQP setting {k!r}")
        setattr(s, k, v)
    return s


class PortOSQP:
    """General sparse QP through the C restatement.  P: any scipy sparse (upper triangle is used)."""

    def __init__(self, P, q, A, l, u, **settings):
        from scipy import sparse

        L = lib()
        P = sparse.triu(sparse.csc_matrix(P), format="csc")
        A = sparse.csc_matrix(A)
        P.sort_indices()
        A.sort_indices()
        self.n, self.m = A.shape[1], A.shape[0]
        self._Pp = np.ascontiguousarray(P.indptr, dtype=np.intc)
        self._Pi = np.ascontiguousarray(P.indices, dtype=np.intc)
        self._Px = np.ascontiguousarray(P.data, dtype=np.float64)
        self._Ap = np.ascontiguousarray(A.indptr, dtype=np.intc)
        self._Ai = np.ascontiguousarray(A.indices, dtype=np.intc)
        self._Ax = np.ascontiguousarray(A.data, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        l = np.ascontiguousarray(l, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        self.settings = default_settings(**settings)
        self._w = L.opq_setup(self.n, self.m, _iptr(self._Pp), _iptr(self._Pi), _dptr(self._Px), _dptr(q),
                              _iptr(self._Ap), _iptr(self._Ai), _dptr(self._Ax), _dptr(l), _dptr(u),
                              C.byref(self.settings))
        if not self._w:
            raise ValueError("opq_setup failed (singular KKT?)")

    def __del__(self):
        if getattr(self, "_w", None):
            lib().opq_free(self._w)
            self._w = None

    def update(self, q=None, l=None, u=None, Ax=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (q, l, u, Ax)]
        if arrs[3] is not None and arrs[3].shape[0] != self._Ax.shape[0]:
            raise ValueError("Ax has the wrong number of stored entries")
        rc = lib().opq_update(self._w, *[_dptr(a) for a in arrs])
        if rc != 0:
            raise ValueError("lower bound must be lower than or equal to upper bound / refactor failed")

    def cold_start(self):
        lib().opq_cold_start(self._w)

    def warm_start(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        lib().opq_warm_start(self._w, _dptr(x), _dptr(y))

    def solve(self):
        x = np.empty(self.n)
        y = np.empty(self.m)
        info = Info()
        lib().opq_solve(self._w, _dptr(x), _dptr(y), C.byref(info))
        return x, y, info

    def vec(self, name: str) -> np.ndarray:
        ln = C.c_int(0)
        p = lib().opq_get_vec(self._w, name.encode(), C.byref(ln))
        return np.ctypeslib.as_array(p, shape=(ln.value,)).copy()

    def scalar(self, name: str) -> float:
        return lib().opq_get_scalar(self._w, name.encode())


def default_config(**kw) -> Config:
    c = Config()
    lib().acmpc_port_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError(f"unknown config field {k!r}")
        if k in ("step_cost", "r_term", "final_cost"):
            arr = getattr(c, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(c, k, v)
    return c


OUTPUT_SPEC = {
    # name: (per-instance shape as a function of (H, n), dtype)
    "controls": (lambda H, n: (2, n), np.float64),
    "prediction": (lambda H, n: (n, 2), np.float64),
    "cum_time": (lambda H, n: (n,), np.float64),
    "states": (lambda H, n: (H, 3), np.float64),
    "v_ref": (lambda H, n: (n,), np.float64),
    "cost": (lambda H, n: (), np.float64),
    "pri_res": (lambda H, n: (), np.float64),
    "dua_res": (lambda H, n: (), np.float64),
    "status": (lambda H, n: (), np.int32),
    "status_speed": (lambda H, n: (), np.int32),
    "iters": (lambda H, n: (2,), np.int32),
    "rho_updates": (lambda H, n: (2,), np.int32),
    "waypoints": (lambda H, n: (7, n), np.float64),
    "derived": (lambda H, n: (3, n - 1), np.float64),
}


def alloc_outputs(B: int, H: int):
    n = H - 1
    arrs = {k: np.zeros((B,) + shp(H, n), dtype=dt) for k, (shp, dt) in OUTPUT_SPEC.items()}
    o = Outputs()
    for k, a in arrs.items():
        setattr(o, k, a.ctypes.data)
    return arrs, o


class PortMPC:
    """One stateful SpatialMPC restated in C (acmpc_port.c)."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.H = cfg.horizon
        self._p = lib().acmpc_port_create(C.byref(cfg))
        if not self._p:
            raise ValueError("acmpc_port_create failed")

    def __del__(self):
        if getattr(self, "_p", None):
            lib().acmpc_port_destroy(self._p)
            self._p = None

    def step(self, path, offset=0.0, v_max=None, is_localised=False, warm=False):
        path = np.ascontiguousarray(path, dtype=np.float64)
        assert path.shape == (self.H, 3)
        arrs, o = alloc_outputs(1, self.H)
        lib().acmpc_port_step(self._p, _dptr(path), float(offset),
                              float(self.cfg.v_max if v_max is None else v_max),
                              int(bool(is_localised)), int(bool(warm)), C.byref(o))
        return {k: v[0] for k, v in arrs.items()}

    def speed_profile(self, waypoints, v_max=None, is_localised=False, end_vel=None, warm=True):
        """compute_speed_profile (spatial_mpc.py:89-123) on this object's persistent speed solvers: `waypoints`
        (7,n) is updated IN PLACE (velocities row only when "solved").  Returns dict(status, x, iters, rho_updates)."""
        w = waypoints
        assert w.dtype == np.float64 and w.flags.c_contiguous and w.shape == (7, self.H - 1)
        x = np.zeros(self.H - 1)
        it, ru = C.c_int(), C.c_int()
        st = lib().acmpc_port_speed_profile(self._p, _dptr(w), float(self.cfg.v_max if v_max is None else v_max),
                                            int(bool(is_localised)), int(end_vel is not None),
                                            float(0.0 if end_vel is None else end_vel), int(bool(warm)), _dptr(x),
                                            C.byref(it), C.byref(ru))
        return dict(status=int(st), x=x, iters=it.value, rho_updates=ru.value)

    def waypoints(self) -> np.ndarray:
        n = self.H - 1
        p = lib().acmpc_port_waypoints(self._p)
        return np.ctypeslib.as_array(p, shape=(7, n)).copy()

    def qp(self, which: str):
        """QP data of the last step: which in {"speed", "control"} -> dict(A (csc), Pdiag, q, l, u)."""
        from scipy import sparse

        n, m = C.c_int(), C.c_int()
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        Ap, Ai, Ax, Pd, q, l, u = ip(), ip(), dp(), dp(), dp(), dp(), dp()
        lib().acmpc_port_get_qp(self._p, 0 if which == "speed" else 1, C.byref(n), C.byref(m),
                                C.byref(Ap), C.byref(Ai), C.byref(Ax), C.byref(Pd), C.byref(q),
                                C.byref(l), C.byref(u))
        n, m = n.value, m.value
        indptr = np.ctypeslib.as_array(Ap, shape=(n + 1,)).copy()
        nnz = int(indptr[-1])
        A = sparse.csc_matrix((np.ctypeslib.as_array(Ax, shape=(nnz,)).copy(),
                               np.ctypeslib.as_array(Ai, shape=(nnz,)).copy(), indptr), shape=(m, n))
        g = lambda p_, k: np.ctypeslib.as_array(p_, shape=(k,)).copy()
        return dict(A=A, Pdiag=g(Pd, n), q=g(q, n), l=g(l, m), u=g(u, m))


def solve_batch(cfg: Config, paths, offsets=None, vmax=None, is_localised=False, nthreads=1):
    """Cold-start batch on `nthreads` host threads; returns dict of numpy arrays."""
    paths = np.ascontiguousarray(paths, dtype=np.float64)
    B, H, _ = paths.shape
    assert H == cfg.horizon
    offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
    vmax = None if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
    arrs, o = alloc_outputs(B, H)
    rc = lib().acmpc_port_solve_batch(C.byref(cfg), B, _dptr(paths), _dptr(offsets), _dptr(vmax),
                                      int(bool(is_localised)), int(nthreads), C.byref(o))
    if rc != 0:
        raise RuntimeError("acmpc_port_solve_batch failed")
    return arrs


# ---- whole-track speed profile -------------------------------------------------
def construct_waypoints(coords) -> np.ndarray:
    """numpy restatement of SpatialMPC.construct_waypoints (/root/reference/src/acmpc/control/
    spatial_mpc.py:125-154) over an (M,3) array -> (7, M-1) ReferencePath rows (velocities = 0)."""
    c = np.asarray(coords, dtype=np.float64)
    cur, nxt = c[:-1, :2], c[1:, :2]
    prev = np.vstack([c[-1:, :2], c[:-2, :2]])          # :135-137: "previous" of point 0 is the last point
    a, b = nxt - cur, cur - prev
    psi = np.arctan2(a[:, 1], a[:, 0])
    d = np.sqrt(a[:, 0] ** 2 + a[:, 1] ** 2)
    behind = np.arctan2(b[:, 1], b[:, 0])
    dang = np.mod(psi - behind + np.pi, 2 * np.pi) - np.pi
    kap = dang / (d + 1e-12) + 1e-12
    kap[0] = kap[1]
    out = np.zeros((7, len(d)))
    out[0], out[1], out[2], out[3], out[4], out[5] = cur[:, 0], cur[:, 1], psi, kap, d, c[1:, 2]
    return out


def map_speed_profile(waypoints, constraints: dict, ay_max: float, a_min: float, max_iter: int = 40000,
                      **settings):
    """compute_map_speed_profile (spatial_mpc.py:60-87) = SpeedProfileSolver.solve (solvers/speed_profile.py:
    15-86, end_velocity None) on a (7,n) ReferencePath array, through the C OSQP restatement.
    Returns (x, info) -- `x` is dec.x; the caller assigns velocities only if info.status == 1."""
    from scipy import sparse

    w = np.asarray(waypoints, dtype=np.float64)
    kap, d = w[3], w[4]
    n = w.shape[1]
    v_max = np.full(n, float(constraints["v_max"]))
    v_dyn = np.sqrt(ay_max / (np.abs(kap) + 1e-12))
    v_dyn[np.abs(kap) < constraints["ki_min"]] = constraints["v_max"]
    vb = np.maximum(constraints["v_min"], np.minimum(v_dyn, v_max)) + 2.0
    D1 = sparse.diags([-1 / (2 * d[:-1]), 1 / (2 * d[:-1])], offsets=[0, 1], shape=[n - 1, n])
    A = sparse.vstack([D1, sparse.eye(n)], format="csc")
    lo = np.hstack([np.full(n - 1, float(a_min)), np.full(n, float(constraints["v_min"]))])
    hi = np.hstack([np.full(n - 1, float(constraints["a_max"])), vb])
    solver = PortOSQP(sparse.eye(n, format="csc"), -1 * vb, A, lo, hi, max_iter=max_iter, **settings)
    x, _, info = solver.solve()
    return x, info


def reference_speeds(velocities, behind: int = 25, ahead: int = 75):
    """agent.py:300 (savgol_filter(v, 21, 3)) and agent.py:137-143 for every map index."""
    from scipy.signal import savgol_filter

    sm = savgol_filter(np.asarray(velocities, dtype=np.float64), 21, 3)
    n = len(sm)
    idx = (np.arange(n)[:, None] + np.arange(-behind, ahead)[None, :])
    return sm, np.mean(sm.take(idx, mode="wrap"), axis=1)


# ---- SpatialBicycleModel (dynamics.py:23-103) and update_prediction (spatial_mpc.py:156-168) ------------------
def t2s(reference_waypoint, reference_state):
    """dynamics.py:23-40: (x, y, psi) of the waypoint and of the vehicle -> (e_y, e_psi, t = 0)."""
    rx, ry, rpsi = reference_waypoint
    x, y, psi = reference_state
    e_y = np.cos(rpsi) * (y - ry) - np.sin(rpsi) * (x - rx)
    e_psi = np.mod(psi - rpsi + np.pi, 2 * np.pi) - np.pi
    return np.array([e_y, e_psi, 0.0])


def s2t(waypoints, states):
    """dynamics.py:42-63: ReferencePath rows (7,n) and spatial states (n,3) -> (3,n) rows X, Y, Psi."""
    w, s = np.asarray(waypoints), np.asarray(states)
    return np.array([w[0] - s[:, 0] * np.sin(w[2]), w[1] + s[:, 0] * np.cos(w[2]), w[2] + s[:, 1]])


def update_prediction(waypoints, states):
    """spatial_mpc.py:156-168: s2t(...)[:-1].T = (n,2) rows (X, Y)."""
    return s2t(waypoints, states)[:-1].T


def linearise(waypoints):
    """dynamics.py:65-103: (7,n) rows -> f (n,3), A (n,3,3), B (n,3,2)."""
    w = np.asarray(waypoints)
    ka, d, v = w[3], w[4], w[6]
    n = w.shape[1]
    eps = 1e-12
    A, B, f = np.zeros((n, 3, 3)), np.zeros((n, 3, 2)), np.zeros((n, 3))
    A[:, 0, 0] = 1.0
    A[:, 0, 1] = d
    A[:, 1, 0] = -(ka ** 2) * d
    A[:, 1, 1] = 1.0
    A[:, 2, 0] = -ka / (v * d + eps)
    A[:, 2, 2] = 1.0
    B[:, 1, 1] = d
    B[:, 2, 0] = -1 / (v ** 2 * d + eps)
    f[:, 2] = 1 / (v * d + eps)
    return f, A, B


# ---- caller side of the step ----------------------------------------------------
def select_command(cum_time, commands, elapsed_time: float):
    """TemporalCommandSelector.get_command (/root/reference/src/acmpc/control/commands.py:22-35);
    cum_time (n,), commands (n,2); arithmetic in the dtype of cum_time, as numpy does with a Python float."""
    distances = cum_time - elapsed_time
    index = int(np.argmin(abs(distances)))
    if distances[index] > 0:
        index -= 1
    n = len(commands)
    index = index if index < n else n - 1
    return commands[index]            # index -1 = the last command, as in the reference


def interpolate_command(cum_time, commands, elapsed_time: float):
    """TemporalCommandInterpolator.get_command (commands.py:59-99); commands (n,2)."""
    distances = cum_time - elapsed_time
    a = int(np.argmin(abs(distances)))
    if a == 0 or a == len(commands) - 1:
        b = a
    elif distances[a] < 0:
        b = a + 1
    else:
        b = a - 1
    if a == b:
        return commands[a]
    x_a, y_a, x_b, y_b = cum_time[a], commands[a], cum_time[b], commands[b]
    return y_a * ((x_b - elapsed_time) / (x_b - x_a)) + y_b * ((elapsed_time - x_a) / (x_b - x_a))


def reference_path(centreline, horizon: int):
    """ControlProcess._reference_path (control/controller.py:257-267)."""
    ds = int(len(centreline) / horizon)
    return np.stack([centreline[0::ds, 0], centreline[0::ds, 1], np.linspace(10.0, 6.0, horizon)]).T
