"""Batched MPC step on one B200: host-side wrapper of the C ABI (include/acmpc_b200.h).

`BatchedMPC` is the batched counterpart of the reference's `SpatialMPC.get_control`
(/root/reference/src/acmpc/control/spatial_mpc.py:170-217): B independent instances per launch, one
warp per instance.  PyTorch is used only to own device buffers and streams; numpy for host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _capi


def config_from_reference(control_config: dict, vehicle_data=None, **osqp_overrides) -> _capi.Config:
    """Translate the dict `build_mpc` receives (controller.py:19-29) + the vehicle object
    (`.vehicle_data.wheelbase`, `.vehicle_data.width`, `.max_steering_angle()`, dynamics.py:11-13)
    into an `acmpc_config`."""
    spc = control_config["speed_profile_constraints"]
    cfg = _capi.default_config()
    cfg.horizon = int(control_config["horizon"])
    cfg.max_iter = int(control_config.get("max_iterations", 4000))   # MAX_SOLVER_ITERATIONS, spatial_mpc.py:17
    cfg.v_min, cfg.v_max = float(spc["v_min"]), float(spc["v_max"])
    cfg.a_min, cfg.a_max = float(spc["a_min"]), float(spc["a_max"])
    cfg.ay_max, cfg.ki_min = float(spc["ay_max"]), float(spc["ki_min"])
    end_v = spc.get("end_velocity")
    cfg.has_end_velocity = 0 if end_v is None else 1
    cfg.end_velocity = 0.0 if end_v is None else float(end_v)
    _capi.apply_overrides(cfg, dict(step_cost=control_config["step_cost"], r_term=control_config["r_term"],
                                    final_cost=control_config["final_cost"]))
    if vehicle_data is not None:
        cfg.wheelbase = float(vehicle_data.vehicle_data.wheelbase)
        cfg.width = float(vehicle_data.vehicle_data.width)
        cfg.delta_max = float(vehicle_data.max_steering_angle())
    # build-time limits of SpatialBicycleModel (controller.py:20-24): frozen here, like the reference
    cfg.input_v_min, cfg.input_v_max = float(spc["v_min"]), float(spc["v_max"])
    _capi.apply_overrides(cfg, osqp_overrides)
    return cfg


class BatchedMPC:
    """Owns one `acmpc_handle`.  The handle (and the CUDA context) is created lazily on first use so
    an object constructed before a fork (controller.py:293-297) initialises CUDA in the child."""

    def __init__(self, cfg: _capi.Config, device: int = 0):
        self.cfg = cfg
        self.device = int(device)
        self.H = int(cfg.horizon)
        self.n = self.H - 1
        self._h = None
        self._lib = _capi.load()
        self._checked_out = (None, 0, None)

    # -- lifetime -------------------------------------------------------------------------------
    def _handle(self):
        if self._h is None:
            h = C.c_void_p()
            rc = self._lib.acmpc_create(C.byref(self.cfg), self.device, C.byref(h))
            if rc != 0:
                raise RuntimeError(
                    f"acmpc_create failed: {_capi.RC_NAMES.get(rc, rc)} "
                    "(the MPC step is CUDA-only; a B200 (sm_100) device is required, there is no CPU fallback)")
            self._h = h
        return self._h

    def close(self):
        if self._h is not None:
            self._lib.acmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.acmpc_last_error(self._h).decode() if self._h else ""
            raise RuntimeError(f"acmpc call failed: {_capi.RC_NAMES.get(rc, rc)} {msg}")

    def launch_info(self) -> Dict[str, int]:
        vals = [C.c_int32() for _ in range(4)]
        self._check(self._lib.acmpc_last_launch_info(self._handle(), *[C.byref(v) for v in vals]))
        return dict(zip(("launches", "smem_bytes", "threads_per_cta", "instances_per_cta"), (v.value for v in vals)))

    def set_profiling(self, on: bool):
        """Record CUDA events around the two kernels of every following launch (acmpc_set_profiling)."""
        self._check(self._lib.acmpc_set_profiling(self._handle(), int(bool(on))))

    def collect_kernel_ms(self) -> Dict[str, float]:
        """Summed device times of the speed-profile and control kernels since the last collect."""
        s, c, k = C.c_double(), C.c_double(), C.c_int32()
        self._check(self._lib.acmpc_collect_kernel_ms(self._handle(), C.byref(s), C.byref(c), C.byref(k)))
        return {"speed_ms": s.value, "control_ms": c.value, "launches": k.value}

    # -- host buffers ---------------------------------------------------------------------------
    def alloc_host_outputs(self, B: int, fields=None, pinned: bool = False):
        """numpy arrays (optionally views of pinned torch tensors) for every requested field."""
        spec = _capi.output_spec(self.H)
        fields = list(spec) if fields is None else list(fields)
        arrs = {}
        keep = []
        for name in fields:
            shp, dt = spec[name]
            if pinned:
                import torch

                t = torch.empty((B,) + shp, dtype=getattr(torch, dt)).pin_memory()
                keep.append(t)
                arrs[name] = t.numpy()
            else:
                arrs[name] = np.empty((B,) + shp, dtype=dt)
        arrs["_keepalive"] = keep
        return arrs

    def _check_host_outputs(self, out, B: int):
        """A caller-supplied `out` dict hands raw pointers to the library: every array must be exactly what the
        ABI writes (shape (B,) + field shape, dtype, C-contiguous, writeable)."""
        spec = _capi.output_spec(self.H)
        for name, a in out.items():
            if name.startswith("_"):
                continue
            if name not in spec:
                raise ValueError(f"unknown output field {name!r}")
            shp, dt = spec[name]
            if not isinstance(a, np.ndarray) or a.shape != (B,) + shp or a.dtype != np.dtype(dt) \
                    or not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError(f"out[{name!r}] must be a writeable C-contiguous {dt} array of shape {(B,) + shp}")

    def solve_host(self, paths, offsets=None, vmax=None, is_localised: bool = False, out=None, fields=None,
                   keep_warm: bool = False):
        """HOST buffers in, HOST buffers out (acmpc_solve_batch_host): H2D + kernels + D2H + sync.
        keep_warm: consecutive calls with the same B behave like B persistent reference objects (OSQP warm
        start + carried rho, spatial_mpc.py:43-58); False = cold start per call."""
        paths = np.ascontiguousarray(paths, dtype=np.float64)
        if paths.ndim != 3 or paths.shape[1:] != (self.H, 3):
            raise ValueError(f"paths must be (B, {self.H}, 3), got {paths.shape}")
        B = paths.shape[0]
        offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64)
        vmax = None if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
        for a, nm in ((offsets, "offsets"), (vmax, "vmax")):
            if a is not None and a.shape != (B,):
                raise ValueError(f"{nm} must have shape ({B},)")
        if out is None:
            out = self.alloc_host_outputs(B, fields)
        elif out is not self._checked_out[0] or B != self._checked_out[1]:
            self._check_host_outputs(out, B)       # once per buffer set: a control loop reuses the same arrays
            o = _capi.Outputs()
            for name in _capi.OUTPUT_FIELDS:
                if name in out:
                    setattr(o, name, out[name].ctypes.data)
            self._checked_out = (out, B, o)
        if out is self._checked_out[0]:
            o = self._checked_out[2]
        else:
            o = _capi.Outputs()
            for name in _capi.OUTPUT_FIELDS:
                if name in out:
                    setattr(o, name, out[name].ctypes.data)
        dp = C.POINTER(C.c_double)
        ptr = lambda a: a.ctypes.data_as(dp) if a is not None else None
        self._check(self._lib.acmpc_solve_batch_host(self._handle(), B, ptr(paths), ptr(offsets), ptr(vmax),
                                                     int(bool(is_localised)), int(bool(keep_warm)), C.byref(o)))
        return out

    # -- stand-alone pieces of the step (cut A) --------------------------------
    def speed_profile_host(self, waypoints, vmax=None, is_localised: bool = False, end_vel=None,
                           keep_warm: bool = False):
        """SpatialMPC.compute_speed_profile (spatial_mpc.py:89-123) for B ReferencePaths: `waypoints` (B,7,n) float64,
        updated IN PLACE (velocities row of the instances whose QP is "solved").  Returns dict(x (B,n) = dec.x,
        status, iters, rho_updates (B,))."""
        w = waypoints
        if not (isinstance(w, np.ndarray) and w.dtype == np.float64 and w.flags.c_contiguous and w.ndim == 3
                and w.shape[1:] == (7, self.n)):
            raise ValueError(f"waypoints must be a C-contiguous float64 (B, 7, {self.n}) array")
        B = w.shape[0]
        vmax = None if vmax is None else np.ascontiguousarray(vmax, dtype=np.float64)
        if vmax is not None and vmax.shape != (B,):
            raise ValueError(f"vmax must have shape ({B},)")
        x = np.zeros((B, self.n))
        st, it, ru = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._check(self._lib.acmpc_speed_profile_batch_host(
            self._handle(), B, w.ctypes.data_as(dp), None if vmax is None else vmax.ctypes.data_as(dp),
            int(bool(is_localised)), int(end_vel is not None), float(0.0 if end_vel is None else end_vel),
            int(bool(keep_warm)), x.ctypes.data_as(dp), st.ctypes.data_as(ip), it.ctypes.data_as(ip), ru.ctypes.data_as(ip)))
        return dict(x=x, status=st, iters=it, rho_updates=ru)

    def speed_profile_device(self, waypoints, vmax=None, is_localised: bool = False, end_vel=None, warm=None,
                             warm_valid: bool = True, stream=None):
        """Same with DEVICE tensors, asynchronous on `stream`: waypoints (B,7,n) float64 CUDA tensor, in place.
        Returns dict of CUDA tensors x (B,n), status (B,), iters (B,2), rho_updates (B,2) (column 0 is written)."""
        import torch

        w = waypoints
        if not (w.is_cuda and w.dtype == torch.float64 and w.is_contiguous() and w.dim() == 3
                and tuple(w.shape[1:]) == (7, self.n)):
            raise ValueError(f"waypoints must be a contiguous float64 CUDA tensor of shape (B, 7, {self.n})")
        B = w.shape[0]
        if vmax is not None and not (vmax.is_cuda and vmax.dtype == torch.float64 and vmax.is_contiguous() and vmax.numel() == B):
            raise ValueError("vmax must be a contiguous float64 CUDA tensor of B elements")
        x = torch.empty((B, self.n), dtype=torch.float64, device=w.device)
        st = torch.empty(B, dtype=torch.int32, device=w.device)
        it = torch.zeros((B, 2), dtype=torch.int32, device=w.device)
        ru = torch.zeros((B, 2), dtype=torch.int32, device=w.device)
        s = torch.cuda.current_stream(w.device) if stream is None else stream
        self._check(self._lib.acmpc_speed_profile_batch_device(
            self._handle(), B, w.data_ptr(), None if vmax is None else vmax.data_ptr(), int(bool(is_localised)),
            int(end_vel is not None), float(0.0 if end_vel is None else end_vel),
            None if warm is None else warm.data_ptr(), int(bool(warm_valid and warm is not None)), x.data_ptr(),
            st.data_ptr(), it.data_ptr(), ru.data_ptr(), C.c_void_p(s.cuda_stream)))
        return dict(x=x, status=st, iters=it, rho_updates=ru)

    def t2s(self, waypoints, states) -> np.ndarray:
        """SpatialBicycleModel.t2s (dynamics.py:23-40) for B (waypoint, state) pairs: (B,3), (B,3) -> (B,3)."""
        w = np.ascontiguousarray(waypoints, dtype=np.float64)
        x = np.ascontiguousarray(states, dtype=np.float64)
        if w.ndim != 2 or w.shape[1] != 3 or x.shape != w.shape:
            raise ValueError("waypoints and states must both be (B, 3)")
        out = np.empty_like(w)
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_t2s_host(self._handle(), w.shape[0], w.ctypes.data_as(dp), x.ctypes.data_as(dp),
                                             out.ctypes.data_as(dp)))
        return out

    def s2t(self, waypoints, states, prediction: bool = False) -> np.ndarray:
        """SpatialBicycleModel.s2t (dynamics.py:42-63): waypoints (B,7,n), states (B,n,3) -> (B,3,n) rows X, Y, Psi;
        prediction=True returns SpatialMPC.update_prediction's (B,n,2) instead (spatial_mpc.py:156-168)."""
        w = np.ascontiguousarray(waypoints, dtype=np.float64)
        x = np.ascontiguousarray(states, dtype=np.float64)
        if w.ndim != 3 or w.shape[1] != 7 or x.shape != (w.shape[0], w.shape[2], 3):
            raise ValueError("waypoints (B,7,n) and states (B,n,3) expected")
        B, _, n = w.shape
        out = np.empty((B, n, 2)) if prediction else np.empty((B, 3, n))
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_s2t_host(self._handle(), B, n, w.ctypes.data_as(dp), x.ctypes.data_as(dp),
                                             None if prediction else out.ctypes.data_as(dp),
                                             out.ctypes.data_as(dp) if prediction else None))
        return out

    def linearise(self, waypoints):
        """SpatialBicycleModel.linearise (dynamics.py:65-103): waypoints (B,7,n) -> f (B,n,3), A (B,n,3,3), B (B,n,3,2)."""
        w = np.ascontiguousarray(waypoints, dtype=np.float64)
        if w.ndim != 3 or w.shape[1] != 7:
            raise ValueError("waypoints (B,7,n) expected")
        B, _, n = w.shape
        f, A, Bm = np.empty((B, n, 3)), np.empty((B, n, 3, 3)), np.empty((B, n, 3, 2))
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_linearise_host(self._handle(), B, n, w.ctypes.data_as(dp), f.ctypes.data_as(dp),
                                                   A.ctypes.data_as(dp), Bm.ctypes.data_as(dp)))
        return f, A, Bm

    # -- whole-track speed profile -----------------------------------
    @staticmethod
    def _map_info(info: "_capi.MapInfo") -> Dict:
        return dict(status=int(info.status), status_str=_capi.STATUS_STRINGS.get(int(info.status), str(info.status)),
                    iters=int(info.iters), rho_updates=int(info.rho_updates), ctas=int(info.ctas),
                    pri_res=float(info.pri_res), dua_res=float(info.dua_res), obj_val=float(info.obj_val),
                    rho=float(info.rho), kernel_ms=float(info.kernel_ms))

    def construct_waypoints(self, track) -> np.ndarray:
        """spatial_mpc.py:125-154 over an (M,3) array -> (7, M-1) rows (acmpc_construct_waypoints_host)."""
        track = np.ascontiguousarray(track, dtype=np.float64)
        if track.ndim != 2 or track.shape[1] != 3 or track.shape[0] < 3:
            raise ValueError(f"track must be (M, 3) with M >= 3, got {track.shape}")
        way = np.zeros((7, track.shape[0] - 1))
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_construct_waypoints_host(self._handle(), track.shape[0], track.ctypes.data_as(dp),
                                                             way.ctypes.data_as(dp)))
        return way

    def map_speed_profile(self, waypoints: np.ndarray, v_max: float, ay_max: float, a_min: float,
                          max_iter: int = 0):
        """spatial_mpc.py:60-87 on a (7,n) ReferencePath array, IN PLACE (velocities row written when "solved").
        Returns (dec_x, info dict)."""
        if not (isinstance(waypoints, np.ndarray) and waypoints.dtype == np.float64 and waypoints.ndim == 2
                and waypoints.shape[0] == 7 and waypoints.flags.c_contiguous):
            raise ValueError("waypoints must be a C-contiguous float64 (7, n) array")
        n = waypoints.shape[1]
        x = np.zeros(n)
        info = _capi.MapInfo()
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_map_speed_profile_host(self._handle(), n, waypoints.ctypes.data_as(dp), float(v_max),
                                                           float(ay_max), float(a_min), int(max_iter),
                                                           x.ctypes.data_as(dp), C.byref(info)))
        return x, self._map_info(info)

    def track_speed_profile(self, track, v_max: float, ay_max: float, a_min: float, max_iter: int = 0):
        """controller.py:49-57 in one launch: (M,3) track -> ((7, M-1) rows incl. velocities, dec_x, info)."""
        track = np.ascontiguousarray(track, dtype=np.float64)
        if track.ndim != 2 or track.shape[1] != 3 or track.shape[0] < 3:
            raise ValueError(f"track must be (M, 3) with M >= 3, got {track.shape}")
        n = track.shape[0] - 1
        way, x, info = np.zeros((7, n)), np.zeros(n), _capi.MapInfo()
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_track_speed_profile_host(self._handle(), track.shape[0], track.ctypes.data_as(dp),
                                                             float(v_max), float(ay_max), float(a_min), int(max_iter),
                                                             way.ctypes.data_as(dp), x.ctypes.data_as(dp),
                                                             C.byref(info)))
        return way, x, self._map_info(info)

    def reference_speeds(self, velocities, behind: int = 25, ahead: int = 75):
        """agent.py:300 + agent.py:137-143: (savgol_filter(v, 21, 3), window mean for every map index)."""
        v = np.ascontiguousarray(velocities, dtype=np.float64)
        if v.ndim != 1 or v.shape[0] < 21:
            raise ValueError("velocities must be a vector of at least 21 samples")
        sm, wm = np.zeros_like(v), np.zeros_like(v)
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_reference_speeds_host(self._handle(), v.shape[0], v.ctypes.data_as(dp), int(behind),
                                                          int(ahead), sm.ctypes.data_as(dp), wm.ctypes.data_as(dp)))
        return sm, wm

    # -- caller side of the step -------------------------------------
    def reference_paths(self, centrelines) -> np.ndarray:
        """ControlProcess._reference_path (controller.py:257-267) for B perceived centre lines:
        (B,P,2) float32 -> (B,H,3) float64.  Raises where the reference's np.stack raises (P // H stride not
        yielding H rows)."""
        c = np.ascontiguousarray(centrelines, dtype=np.float32)
        if c.ndim != 3 or c.shape[2] != 2:
            raise ValueError(f"centrelines must be (B, P, 2), got {c.shape}")
        out = np.empty((c.shape[0], self.H, 3))
        self._check(self._lib.acmpc_reference_paths_host(self._handle(), c.shape[0], c.shape[1],
                                                         c.ctypes.data_as(C.POINTER(C.c_float)),
                                                         out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def publish(self, controls, cum_time, prediction):
        """ControlProcess._update_shared_memory (controller.py:274-280): the float32 arrays the agent process reads,
        (control_inputs (B,n,2) = projected_control.T, control_cumtime (B,n), predicted_locations (B,n,2))."""
        controls = np.ascontiguousarray(controls, dtype=np.float64)
        cum_time = np.ascontiguousarray(cum_time, dtype=np.float64)
        prediction = np.ascontiguousarray(prediction, dtype=np.float64)
        B, n = cum_time.shape
        if n != self.n or controls.shape != (B, 2, n) or prediction.shape != (B, n, 2):
            raise ValueError("controls (B,2,n), cum_time (B,n), prediction (B,n,2) expected")
        ci, ct, pl = (np.empty((B, n, 2), np.float32), np.empty((B, n), np.float32), np.empty((B, n, 2), np.float32))
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        self._check(self._lib.acmpc_publish_host(self._handle(), B, controls.ctypes.data_as(dp), cum_time.ctypes.data_as(dp),
                                                 prediction.ctypes.data_as(dp), ci.ctypes.data_as(fp),
                                                 ct.ctypes.data_as(fp), pl.ctypes.data_as(fp)))
        return ci, ct, pl

    def select_commands(self, cum_time, commands, elapsed, interpolate: bool = False, return_indices: bool = False):
        """commands.py: TemporalCommandSelector (default) / TemporalCommandInterpolator lookup for B instances.
        cum_time (B,n), commands (B,n,2), elapsed (B,).  float32 inputs are processed in float32 (the shared-memory
        views), anything else in float64."""
        cum_time = np.asarray(cum_time)
        f32 = cum_time.dtype == np.float32
        dt, ct_t = (np.float32, C.c_float) if f32 else (np.float64, C.c_double)
        cum_time = np.ascontiguousarray(cum_time, dtype=dt)
        commands = np.ascontiguousarray(commands, dtype=dt)
        elapsed = np.ascontiguousarray(elapsed, dtype=np.float64)
        B, n = cum_time.shape
        if commands.shape != (B, n, 2) or elapsed.shape != (B,):
            raise ValueError("cum_time (B,n), commands (B,n,2), elapsed (B,) expected")
        out, idx = np.empty((B, 2), dt), np.empty((B, 2), np.int32)
        fn = self._lib.acmpc_select_commands_f32_host if f32 else self._lib.acmpc_select_commands_f64_host
        tp = C.POINTER(ct_t)
        self._check(fn(self._handle(), B, n, cum_time.ctypes.data_as(tp), commands.ctypes.data_as(tp),
                       elapsed.ctypes.data_as(C.POINTER(C.c_double)), int(bool(interpolate)), out.ctypes.data_as(tp),
                       idx.ctypes.data_as(C.POINTER(C.c_int32))))
        return (out, idx) if return_indices else out

    # -- track side of the step ---------------------------------
    def remove_near_duplicate_points(self, track, tol: float = 0.0001) -> np.ndarray:
        """utils/load.py:30-35 on the device: (M,cols>=2) -> the rows whose (x, y) = columns 0, 1 lie farther than `tol`
        from their predecessor's; further columns travel with their row (numpy's track[is_not_duplicated])."""
        t = np.ascontiguousarray(track, dtype=np.float64)
        if t.ndim != 2 or t.shape[1] < 2:
            raise ValueError(f"track must be (M, >=2), got {t.shape}")
        out, kept = np.empty_like(t), C.c_int32(0)
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_remove_near_duplicates_cols_host(self._handle(), t.shape[0], t.shape[1],
                                                                     t.ctypes.data_as(dp), float(tol),
                                                                     out.ctypes.data_as(dp), C.byref(kept)))
        return out[:kept.value].copy()

    def smooth_tracks_with_polyfit(self, tracks, num_points: int, degree: int = 3, return_info: bool = False):
        """perception/utils.py:107-119 for a list of (m_b, 2) tracks (ragged; empty tracks allowed) ->
        (B, num_points, 2).  return_info adds (status[B], start_index[B])."""
        tracks = [np.asarray(t, dtype=np.float64).reshape(-1, 2) for t in tracks]
        B = len(tracks)
        off = np.zeros(B + 1, np.int32)
        off[1:] = np.cumsum([t.shape[0] for t in tracks])
        pts = np.ascontiguousarray(np.concatenate(tracks, axis=0)) if off[-1] else np.zeros((0, 2))
        out = np.empty((B, int(num_points), 2))
        st, si = np.empty(B, np.int32), np.empty(B, np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        self._check(self._lib.acmpc_smooth_tracks_polyfit_host(
            self._handle(), B, off.ctypes.data_as(ip), pts.ctypes.data_as(dp) if off[-1] else None, int(num_points),
            int(degree), out.ctypes.data_as(dp), st.ctypes.data_as(ip), si.ctypes.data_as(ip)))
        return (out, st, si) if return_info else out

    def centre_tracks(self, left, right, num_points: Optional[int] = None) -> np.ndarray:
        """TrackLimitPerception._calculate_centre_track (perception/tracks.py:247-252) for B frames:
        left / right (B,N,2) -> (B,num_points,2) (num_points defaults to N = n_polyfit_points)."""
        left = np.ascontiguousarray(left, dtype=np.float64)
        right = np.ascontiguousarray(right, dtype=np.float64)
        if left.ndim != 3 or left.shape != right.shape or left.shape[2] != 2:
            raise ValueError("left and right must both be (B, N, 2)")
        B, N = left.shape[:2]
        num_points = N if num_points is None else int(num_points)
        out = np.empty((B, num_points, 2))
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_centre_tracks_host(self._handle(), B, N, left.ctypes.data_as(dp), right.ctypes.data_as(dp),
                                                       num_points, out.ctypes.data_as(dp), None))
        return out

    def extract_paths(self, centreline, indices, offset_lat=None, offset_psi=None, lookahead: float = 100.0,
                      ds: float = 0.5) -> np.ndarray:
        """SURVEY.md section 8d instances from a map centre line (M,2), host buffers: -> (B,H,3)."""
        cl = np.ascontiguousarray(centreline, dtype=np.float64)
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        B = idx.shape[0]
        lat = None if offset_lat is None else np.ascontiguousarray(offset_lat, dtype=np.float64)
        psi = None if offset_psi is None else np.ascontiguousarray(offset_psi, dtype=np.float64)
        out = np.empty((B, self.H, 3))
        dp = C.POINTER(C.c_double)
        self._check(self._lib.acmpc_extract_paths_host(
            self._handle(), cl.shape[0], cl.ctypes.data_as(dp), B, idx.ctypes.data_as(C.POINTER(C.c_int32)),
            None if lat is None else lat.ctypes.data_as(dp), None if psi is None else psi.ctypes.data_as(dp),
            float(lookahead), float(ds), out.ctypes.data_as(dp)))
        return out

    def extract_paths_device(self, centreline, indices, offset_lat=None, offset_psi=None, lookahead: float = 100.0,
                             ds: float = 0.5, out=None, stream=None):
        """Same with DEVICE tensors (centreline (M,2) f64, indices (B,) int32, offsets (B,) f64), asynchronous on
        `stream`; the (B,H,3) result feeds solve_device without touching the host."""
        import torch

        if not (centreline.is_cuda and centreline.dtype == torch.float64 and centreline.is_contiguous() and
                centreline.dim() == 2 and centreline.shape[1] == 2):
            raise ValueError("centreline must be a contiguous (M, 2) float64 CUDA tensor")
        if not (indices.is_cuda and indices.dtype == torch.int32 and indices.is_contiguous() and indices.dim() == 1):
            raise ValueError("indices must be a contiguous int32 CUDA tensor")
        B = indices.shape[0]
        for a in (offset_lat, offset_psi):
            if a is not None and not (a.is_cuda and a.dtype == torch.float64 and a.is_contiguous() and a.numel() == B):
                raise ValueError("offset_lat / offset_psi must be contiguous float64 CUDA tensors of B elements")
        if out is None:
            out = torch.empty((B, self.H, 3), dtype=torch.float64, device=centreline.device)
        s = torch.cuda.current_stream(centreline.device) if stream is None else stream
        self._check(self._lib.acmpc_extract_paths_device(
            self._handle(), centreline.shape[0], centreline.data_ptr(), B, indices.data_ptr(),
            None if offset_lat is None else offset_lat.data_ptr(), None if offset_psi is None else offset_psi.data_ptr(),
            float(lookahead), float(ds), out.data_ptr(), C.c_void_p(s.cuda_stream)))
        return out

    # -- device buffers -------------------------------------------------------------------------
    def alloc_device_outputs(self, B: int, fields=None):
        """One packed uint8 CUDA tensor with a 256-byte aligned slab per field (so a multi-GPU run
        gathers ONE buffer), plus typed views.  Returns (packed, views)."""
        import torch

        spec = _capi.output_spec(self.H)
        fields = list(spec) if fields is None else list(fields)
        offs, total = {}, 0
        for name in fields:
            shp, dt = spec[name]
            nbytes = B * int(np.prod(shp, dtype=np.int64)) * np.dtype(dt).itemsize
            offs[name] = (total, nbytes)
            total = (total + nbytes + 255) // 256 * 256
        packed = torch.empty(max(total, 256), dtype=torch.uint8, device=f"cuda:{self.device}")
        views = {}
        for name in fields:
            shp, dt = spec[name]
            o, nb = offs[name]
            views[name] = packed[o:o + nb].view(getattr(torch, dt)).view((B,) + shp)
        return packed, views

    @staticmethod
    def unpack(packed, B: int, H: int, fields=None):
        """Typed views of a packed output buffer (same layout as alloc_device_outputs)."""
        import torch

        spec = _capi.output_spec(H)
        fields = list(spec) if fields is None else list(fields)
        views, total = {}, 0
        for name in fields:
            shp, dt = spec[name]
            nbytes = B * int(np.prod(shp, dtype=np.int64)) * np.dtype(dt).itemsize
            views[name] = packed[total:total + nbytes].view(getattr(torch, dt)).view((B,) + shp)
            total = (total + nbytes + 255) // 256 * 256
        return views

    def warm_stride(self) -> int:
        """Bytes of one instance's warm-start record."""
        return int(self._lib.acmpc_warm_stride(self._handle()))

    def alloc_warm(self, B: int):
        """Zero-filled device buffer of B warm-start records ("no solver object yet")."""
        import torch

        return torch.zeros(B * self.warm_stride() // 8, dtype=torch.float64, device=f"cuda:{self.device}")

    def attach_completion(self, flag_ptr: int = 0, flag_value: int = 0, credit_table_ptr: int = 0, credit_n: int = 0,
                          credit_value: int = 0, credit_wait_ptr: int = 0, credit_need: int = 0):
        """One-shot for the next solve_device (acmpc_attach_completion): raw device addresses, see include/acmpc_b200.h."""
        self._check(self._lib.acmpc_attach_completion(self._handle(), C.c_void_p(flag_ptr or None), int(flag_value),
                                                      C.c_void_p(credit_table_ptr or None), int(credit_n), int(credit_value),
                                                      C.c_void_p(credit_wait_ptr or None), int(credit_need)))

    def stream_wait_value32(self, addr: int, value: int, stream=None):
        """`stream` (torch stream, default: current) waits until the LOCAL device word at `addr` is >= value."""
        import torch

        s = torch.cuda.current_stream(self.device) if stream is None else stream
        self._check(self._lib.acmpc_stream_wait_value32(self._handle(), C.c_void_p(addr), int(value), C.c_void_p(s.cuda_stream)))

    def pipeline(self, B: int, fields=None, depth: int = 2):
        """A HostPipeline (ac_mpc_b200/pipeline.py) over this solver: submit()/wait() for streams of B-instance batches,
        the H2D / kernels / D2H of consecutive batches overlapped."""
        from .pipeline import HostPipeline

        return HostPipeline(self, B, fields, depth)

    def solve_device(self, paths, offsets=None, vmax=None, is_localised: bool = False, out=None, stream=None,
                     warm=None, warm_valid: bool = True):
        """DEVICE tensors in/out (acmpc_solve_batch_device); asynchronous on `stream` (torch stream or
        None = torch's current stream).  `out` = dict of CUDA tensors from alloc_device_outputs.
        `warm`: tensor from alloc_warm (read when warm_valid, always rewritten) or None = cold start."""
        import torch

        if not (paths.is_cuda and paths.dtype == torch.float64 and paths.is_contiguous()):
            raise ValueError("paths must be a contiguous float64 CUDA tensor")
        if paths.dim() != 3 or tuple(paths.shape[1:]) != (self.H, 3):
            raise ValueError(f"paths must be (B, {self.H}, 3)")
        B = paths.shape[0]
        for a in (offsets, vmax):
            if a is not None and not (a.is_cuda and a.dtype == torch.float64 and a.is_contiguous() and a.numel() == B):
                raise ValueError("offsets / vmax must be contiguous float64 CUDA tensors of B elements")
        if out is None:
            _, out = self.alloc_device_outputs(B)
        o = _capi.Outputs()
        for name in _capi.OUTPUT_FIELDS:
            if name in out:
                setattr(o, name, out[name].data_ptr())
        s = torch.cuda.current_stream(paths.device) if stream is None else stream
        self._check(self._lib.acmpc_solve_batch_device(
            self._handle(), B, paths.data_ptr(), None if offsets is None else offsets.data_ptr(),
            None if vmax is None else vmax.data_ptr(), int(bool(is_localised)),
            None if warm is None else warm.data_ptr(), int(bool(warm_valid and warm is not None)), C.byref(o),
            C.c_void_p(s.cuda_stream)))
        return out


def fp64_peak_tflops(device: int = 0) -> float:
    """Measured FP64 FMA throughput of the device (roofline denominator of this path)."""
    v = C.c_double()
    rc = _capi.load().acmpc_fp64_peak_tflops(int(device), C.byref(v))
    if rc != 0:
        raise RuntimeError(f"acmpc_fp64_peak_tflops failed: {_capi.RC_NAMES.get(rc, rc)}")
    return v.value
