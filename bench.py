#!/usr/bin/env python
"""bench.py -- batched MPC solves/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                     # the CPU arm (real osqp + reference Python when both are
                                                             # importable, else the oracle's C port), all host threads
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1, one rank per GPU

A "step" is one pass of the hot path (SpatialMPC.get_control: waypoints + speed-profile QP + linearise/assemble +
control QP + unpack/rollout/cost) over one batch of synthetic instances: BASELINE.json configs[1] = Monza racing block
(H = 50), 4096 perturbed initial states per GPU.  At N > 1 every rank solves its own 4096-instance shard (weak scaling,
no data-path collective) through the PRODUCT call `ac_mpc_b200.sharded.ShardedMPC`, whose single exchange lands every
rank's results on rank 0: by default the kernels store straight into rank 0's slab over NVLink (peer-mapped memory) and
the per-step collective is a one-element completion signal; --transport nccl = one NCCL gather per step.  The exchange
of step i overlaps the kernels of step i + 1; every exchange lies inside a timed region (the last one has its own).

One JSON line on stdout (rank 0).  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the same
metric through the public host API (N = 1: SpatialMPC.get_control_batch; N > 1: ShardedMPC.solve) with pinned HOST
buffers, H2D / D2H (and the exchange) inside the timed region.  Further legs (outside the timed steps, rank 0, N = 1
unless stated): BASELINE configs[2] (Nordschleife every-waypoint sweep, cold + warm replay), configs[3] (Spa H = 20 / 40 /
80 x 16384), H = 100, configs[4] (all 7 tracks, 1 M instances split over the N ranks, solve-only and solve + exchange).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched MPC solves/sec"
UNIT = "solves/s"
BENCH_FIELDS = ["controls", "prediction", "cum_time", "v_ref", "cost", "pri_res", "dua_res", "status",
                "status_speed", "iters", "rho_updates"]
VEH = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                     "max_steering_angle": lambda self: 0.30})()


# ------------------------------------------------------------------------------------------------
def workload(track: str, B: int, H: int, rank: int):
    """Synthetic batch: waypoint index uniform over the synthetic centreline of the
    named length, lateral U(-2,2) m, heading U(-0.1,0.1) rad, v_max U(20,84) m/s; seed = 1 + rank."""
    from ac_mpc_b200 import tracks

    cl = tracks.synthetic_centreline(track)
    paths, vmax = tracks.perturbed_batch(track, B, horizon=H, seed=1 + rank, centreline=cl)
    return paths, vmax


def config_kwargs(track: str, H: int):
    from ac_mpc_b200 import tracks

    r = tracks.RACING_CONTROL[track]
    return dict(horizon=H, v_min=r["v_min"], v_max=84.0, a_min=r["a_min"], a_max=1.0, ay_max=r["ay_max"],
                ki_min=r["ki_min"], end_velocity=0.0 if r["end_velocity"] is None else r["end_velocity"],
                has_end_velocity=0 if r["end_velocity"] is None else 1, step_cost=r["step_cost"],
                r_term=[1e-2, 10.0], final_cost=[1.0, 0.0, 0.1], input_v_min=r["v_min"], input_v_max=84.0)


def bench_config(args, world: int) -> dict:
    """The `config` object of the JSON line: identical for both arms (--impl b200 / reference) of one invocation."""
    return {"workload": f"{args.track} racing block H={args.horizon}, {args.batch} perturbed initial states per GPU "
                        "(BASELINE configs[1]), cold start per instance",
            "track": args.track, "horizon": args.horizon, "batch_per_gpu": args.batch, "global_batch": world * args.batch,
            "n_gpus": world, "osqp": "eps_abs=eps_rel=1e-3, check 25, adaptive rho interval 50, max_iter 4000"}


def flops_per_batch(H: int, iters: np.ndarray, rho_updates: np.ndarray) -> tuple[float, float]:
    """Algorithmic FP64 flop model of SURVEY.md 8(d) / DESIGN.md, summed over the kernel-reported
    per-instance ADMM iteration counts (column 0 speed QP, column 1 control QP).  Returns the flops of the
    (speed-profile kernel, control kernel): waypoints + speed QP / assembly + Ruiz + control QP + rollout."""
    Ks, Kc = iters[:, 0].astype(np.float64), iters[:, 1].astype(np.float64)
    Fs, Fc = 1.0 + rho_updates[:, 0], 1.0 + rho_updates[:, 1]
    control = Kc * (319 * H - 70) + (Kc / 25.0) * (96 * H + 12 * 13 * H) + Fc * 300 * H + 40 * (16 * H - 10) + 100 * H
    speed = Ks * 45 * H + Fs * 10 * H + 50 * H
    return float(speed.sum()), float(control.sum())


def bytes_per_solve(H: int, fields) -> tuple[int, int]:
    from ac_mpc_b200 import _capi

    spec = _capi.output_spec(H)
    out = sum(int(np.prod(spec[f][0], dtype=np.int64)) * np.dtype(spec[f][1]).itemsize for f in fields)
    return 8 * (3 * H + 1), out     # paths + v_max in, requested fields out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the run (B200_PROFILING.md recipe).  Started before the warm-up so the
    20 ms poll has produced samples by the time the (12 ms long) timed region runs; `mark()` brackets the timed region
    and the summary reports the samples inside the marks when there are any, else all samples under load."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.marks = index, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for t, ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                rows.append((t, float(parts[0]), float(parts[1]), float(parts[2]),
                             [nm for nm, val in zip(names, parts[3:7]) if val.lower().startswith("active")]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if len(self.marks) >= 2 and self.marks[0] - 0.02 <= r[0] <= self.marks[-1] + 0.02]
        loaded = [r for r in rows if r[3] > 0.5 * max(x[3] for x in rows)]
        use, scope = (inside, "timed region") if inside else (loaded, "whole run, samples under load")
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "power_w_max": float(max(r[3] for r in rows)), "samples": len(use), "samples_total": len(rows),
                "scope": scope, "reasons": sorted({nm for r in use for nm in r[4]})}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def recorded_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, taken from the
    committed ncu capture (profiles/traffic.json) -- null until such a capture exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f)
        except Exception:
            return None
    return None


def p50(fn, n):
    ts = []
    for i in range(n + 20):
        t0 = time.perf_counter()
        fn(i)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts[20:]) * 1e3)


# ------------------------------------------------------------------------------------------------
# CPU arm
def cpu_port_rate(cfg_kw, paths, vmax, threads: int, min_seconds: float, max_reps: int):
    """Oracle port (oracle/acmpc_port.c, cold start per instance) on `threads` host threads."""
    from oracle import port

    cfg = port.default_config(**cfg_kw)
    port.solve_batch(cfg, paths[:64], None, vmax[:64], False, nthreads=threads)      # warm the code path
    done, t0 = 0, time.perf_counter()
    for _ in range(max_reps):
        port.solve_batch(cfg, paths, None, vmax, False, nthreads=threads)
        done += paths.shape[0]
        if time.perf_counter() - t0 >= min_seconds:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


_REF = {}


def _ref_init(track, H, src, force_port):
    """Pool worker: import the UNMODIFIED reference (`src`) on the selected osqp module."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shim"), src]
    from oracle import osqp_select

    mod, label = osqp_select.select(prefer_real=not force_port)
    osqp_select.install(mod)
    from ace.steering import SteeringGeometry
    from acmpc.control.controller import build_mpc

    from ac_mpc_b200 import tracks

    _REF.update(build=build_mpc, veh=SteeringGeometry(), cfg=tracks.racing_config(track, H), label=label)


def _ref_chunk(job):
    """Cold start per instance, like the GPU arm: a fresh SpatialMPC (fresh OSQP objects) for every instance."""
    paths, vmax = job
    out = np.zeros((paths.shape[0], 2, paths.shape[1] - 1))
    for b in range(paths.shape[0]):
        cfg = dict(_REF["cfg"], speed_profile_constraints=dict(_REF["cfg"]["speed_profile_constraints"], v_max=float(vmax[b])))
        mpc = _REF["build"](cfg, _REF["veh"])
        mpc.get_control(paths[b], False, 0.0)
        out[b] = mpc.projected_control
    return out


def reference_python_rate(track, H, paths, vmax, procs, src, force_port, steps=1, warmup=0):
    """The reference's own Python (SpatialMPC.get_control) over a multiprocessing pool; returns (rate, label, seconds)."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_ref_init, initargs=(track, H, src, force_port)) as pool:
        chunks = [(paths[i:i + 16], vmax[i:i + 16]) for i in range(0, paths.shape[0], 16)]
        for _ in range(warmup):
            pool.map(_ref_chunk, chunks[: 2 * procs])
        t0 = time.perf_counter()
        for _ in range(steps):
            res = pool.map(_ref_chunk, chunks)
        dt = time.perf_counter() - t0
        label = pool.apply(_ref_label)
    return paths.shape[0] * steps / dt, label, dt, np.concatenate(res)


def _ref_label():
    return _REF["label"]


def reference_arm_mode(args):
    """("python", src) when the unmodified reference can run here on a REAL osqp wheel (or --ref-python forces it onto the
    stand-in), else ("port", None): /root/reference is absent on the GPU box and this image has no osqp wheel."""
    from oracle import osqp_select

    src = args.reference_src
    have_src = os.path.isdir(os.path.join(src, "acmpc", "control"))
    if have_src and (osqp_select.info().real or args.ref_python):
        return "python", src
    return "port", None


def run_reference(args, rank: int):
    """--impl reference: the CPU implementation of the path on the box's host cores, full batch per step.  Tries the real
    thing first (unmodified reference Python on an importable `osqp` wheel, multiprocessing pool); in this image and on the
    GPU box neither the wheel nor /root/reference exist, so the arm times the oracle's C port of the same path on all host
    threads -- faster than the reference's Python (no ~6 ms of scipy.sparse glue per call): a conservative baseline."""
    if rank != 0:
        return
    from oracle import osqp_select, port

    threads = len(os.sched_getaffinity(0))
    kw = config_kwargs(args.track, args.horizon)
    paths, vmax = workload(args.track, args.batch, args.horizon, 0)
    mode, src = reference_arm_mode(args)
    if mode == "python":
        rate, label, dt, _ = reference_python_rate(args.track, args.horizon, paths, vmax, threads, src,
                                                   force_port=not osqp_select.info().real, steps=args.steps,
                                                   warmup=min(args.warmup, 1))
        value, kind, ms = rate, "reference", dt / args.steps * 1e3
        sample = (f"all {args.batch} instances per step, {args.steps} steps, unmodified reference Python "
                  f"(SpatialMPC.get_control, fresh object per instance) over a {threads}-process pool on {label}")
        solver = label
    else:
        cfg = port.default_config(**kw)
        for _ in range(args.warmup):
            port.solve_batch(cfg, paths, None, vmax, False, nthreads=threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            port.solve_batch(cfg, paths, None, vmax, False, nthreads=threads)
        dt = time.perf_counter() - t0
        value, kind, ms = args.batch * args.steps / dt, "port", dt / args.steps * 1e3
        sample = (f"all {args.batch} instances per step, {args.steps} steps, oracle/acmpc_port.c on {threads} pthreads "
                  f"(osqp wheel importable: {osqp_select.info().real}; reference sources present: "
                  f"{os.path.isdir(args.reference_src)})")
        solver = osqp_select.PORT_LABEL
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "solver": solver, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# extra legs (BASELINE configs[2], [3], [4], H = 100); all outside the timed steps
def _device_rate(mpc, d_paths, d_vmax, views, reps, warm=None, warm_valid=True, loc=False):
    """Device-timed solves/s of solve_device on resident inputs (CUDA events on the launching stream, L2 flushed by the
    caller's inputs being larger than L2 or by the flush tensor between reps)."""
    import torch

    for _ in range(2):
        mpc.solve_device(d_paths, None, d_vmax, loc, out=views, warm=warm, warm_valid=warm_valid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for _ in range(reps):
        _FLUSH["t"].fill_(1)
        e0.record()
        mpc.solve_device(d_paths, None, d_vmax, loc, out=views, warm=warm, warm_valid=warm_valid)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return d_paths.shape[0] * reps / (ms * 1e-3), ms / reps


_FLUSH = {}


def _spot_check(kw, paths, vmax, got, sample=64, seed=0, loc=False):
    """Parity of a random sample against the oracle port (not timed): identical iteration counts, max |d controls|."""
    from oracle import port

    B = paths.shape[0]
    sub = np.random.default_rng(seed).choice(B, min(sample, B), replace=False)
    want = port.solve_batch(port.default_config(**kw), paths[sub], None, None if vmax is None else vmax[sub], loc,
                            nthreads=min(16, len(os.sched_getaffinity(0))))
    ok = want["status"] == 1
    return {"sample": int(sub.shape[0]), "iters_equal": bool(np.array_equal(got["iters"][sub], want["iters"])),
            "status_equal": bool(np.array_equal(got["status"][sub], want["status"])),
            "max_abs_dcontrols": float(np.abs(got["controls"][sub][ok] - want["controls"][ok]).max()) if ok.any() else None}


def leg_horizon(track, H, B, fp64_peak, local_rank, reps=5):
    """One (track, horizon, batch) configuration: device-timed rate, roofline fraction of the whole step, parity sample."""
    import torch

    from ac_mpc_b200 import BatchedMPC, _capi

    kw = config_kwargs(track, H)
    paths, vmax = workload(track, B, H, 7)
    mpc = BatchedMPC(_capi.default_config(**kw), device=local_rank)
    dev = torch.device("cuda", local_rank)
    dp, dv = torch.from_numpy(paths).to(dev), torch.from_numpy(vmax).to(dev)
    _, views = mpc.alloc_device_outputs(B, ["controls", "status", "iters", "rho_updates", "cost"])
    rate, ms = _device_rate(mpc, dp, dv, views, reps)
    got = {k: v.cpu().numpy() for k, v in views.items()}
    fs, fc = flops_per_batch(H, got["iters"], got["rho_updates"])
    info = mpc.launch_info()
    mpc.close()
    return {"track": track, "horizon": H, "batch": B, "value": rate, "unit": UNIT, "ms_per_batch": ms,
            "roofline_frac_whole_step": (fs + fc) / (ms * 1e-3) / 1e12 / fp64_peak,
            "iters_mean": [float(got["iters"][:, 0].mean()), float(got["iters"][:, 1].mean())],
            "solved_frac": float((got["status"] == 1).mean()), "smem_bytes_per_cta": info["smem_bytes"],
            "parity": _spot_check(kw, paths, vmax, got)}


def leg_nordschleife(fp64_peak, local_rank):
    """BASELINE configs[2]: one instance per metre of the Nordschleife centre line (~20.8 k, unperturbed), paths built on
    the device from the resident centre line; cold, then a closed-loop warm replay (every instance advances 2 m per step,
    warm-started from its own previous solve; the s2t prediction and the cost come back with the controls)."""
    import torch

    from ac_mpc_b200 import BatchedMPC, _capi, tracks
    from oracle import port

    kw = config_kwargs("nordschleife", 50)
    cl = tracks.synthetic_centreline("nordschleife")
    idx = np.arange(0, cl.shape[0], 2, dtype=np.int32)
    B = idx.shape[0]
    dev = torch.device("cuda", local_rank)
    mpc = BatchedMPC(_capi.default_config(**kw), device=local_rank)
    d_cl, d_idx = torch.from_numpy(cl).to(dev), torch.from_numpy(idx).to(dev)
    fields = ["controls", "prediction", "cost", "status", "iters", "rho_updates"]
    _, views = mpc.alloc_device_outputs(B, fields)
    d_paths = mpc.extract_paths_device(d_cl, d_idx)
    cold_rate, cold_ms = _device_rate(mpc, d_paths, None, views, 5)
    cold = {k: v.cpu().numpy() for k, v in views.items()}
    paths0 = d_paths.cpu().numpy()
    # closed-loop replay: T steps of 2 m, warm records per instance, paths re-extracted on the device every step
    T = 6
    warm = mpc.alloc_warm(B)
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
    M = cl.shape[0]
    step_idx = [torch.from_numpy(((idx.astype(np.int64) + 4 * t) % M).astype(np.int32)).to(dev) for t in range(T)]
    sample = np.random.default_rng(3).choice(B, 16, replace=False)
    objs = {int(b): port.PortMPC(port.default_config(**kw)) for b in sample}
    iters_equal, dmax, its, it_max, n_solved = True, 0.0, [], [], []
    torch.cuda.synchronize()
    for t in range(T):
        _FLUSH["t"].fill_(1)
        e[t][0].record()
        mpc.extract_paths_device(d_cl, step_idx[t], out=d_paths)
        mpc.solve_device(d_paths, None, None, False, out=views, warm=warm)
        e[t][1].record()
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in views.items()}
        its.append(float(got["iters"][:, 1].mean()))
        it_max.append([int(got["iters"][:, 0].max()), int(got["iters"][:, 1].max())])
        n_solved.append(int((got["status"] == 1).sum()))
        p_host = d_paths.cpu().numpy()
        for b, obj in objs.items():
            want = obj.step(p_host[b], 0.0, None, False, warm=True)
            iters_equal = iters_equal and got["iters"][b].tolist() == want["iters"].tolist()
            if want["status"] == 1:
                dmax = max(dmax, float(np.abs(got["controls"][b] - want["controls"]).max()))
    warm_ms = [e[t][0].elapsed_time(e[t][1]) for t in range(1, T)]
    mpc.close()
    return {"instances": int(B), "cold": {"value": cold_rate, "unit": UNIT, "ms_per_sweep": cold_ms,
                                           "solved_frac": float((cold["status"] == 1).mean()),
                                           "parity": _spot_check(kw, paths0, None, cold)},
            "warm_replay": {"steps": T, "value": B / (float(np.mean(warm_ms)) * 1e-3), "unit": UNIT,
                            "ms_per_step": float(np.mean(warm_ms)), "includes": "device path extraction + both kernels",
                            "control_iters_mean_per_step": its, "iters_max_per_step_speed_control": it_max,
                            "solved_per_step": n_solved,
                            "note": "a step lasts as long as its slowest instance: one warp runs one instance, so a single "
                                    "instance that needs hundreds of ADMM iterations from its warm start sets the step time",
                            "parity_vs_oracle_objects": {
                                "sample": int(sample.shape[0]), "iters_equal": bool(iters_equal), "max_abs_dcontrols": dmax}},
            "outputs": "controls, s2t prediction, cost, status per instance"}


def leg_all_tracks(args, rank, local_rank, world, fp64_peak):
    """BASELINE configs[4]: all 7 racing blocks, 1 M instances split evenly over the tracks and over the N ranks (strong
    scaling: total work fixed), paths built on the device; per rank one ShardedMPC per track (its own config), timed
    solve-only (kernels of all 7 tracks) and solve + exchange to rank 0; 64 instances per track checked against the port."""
    import torch
    import torch.distributed as dist

    from ac_mpc_b200 import _capi, sharding, tracks
    from ac_mpc_b200.sharded import ShardedMPC

    dev = torch.device("cuda", local_rank)
    total = args.all_tracks_instances
    per_track = total // len(tracks.TRACK_ORDER)
    fields = ["controls", "status", "iters", "rho_updates", "cost"]
    work = []
    for ti, tr in enumerate(tracks.TRACK_ORDER):
        cl = tracks.synthetic_centreline(tr)
        rng = np.random.default_rng(500 + ti)
        idx = rng.integers(0, cl.shape[0], per_track).astype(np.int32)
        lat, psi, vm = rng.uniform(-2, 2, per_track), rng.uniform(-0.1, 0.1, per_track), rng.uniform(20, 84, per_track)
        lo, hi = sharding.shard_range(per_track, rank, world)
        sh = ShardedMPC(_capi.default_config(**config_kwargs(tr, 50)), fields=fields, device=local_rank,
                        transport=args.transport)
        d = [torch.from_numpy(a[lo:hi].copy()).to(dev) for a in (idx, lat, psi, vm)]
        d_paths = sh.solver.extract_paths_device(torch.from_numpy(cl).to(dev), d[0], d[1], d[2])
        work.append((tr, sh, d_paths, d[3], (cl, idx, lat, psi, vm)))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(exchange: bool, reps: int):
        ms = []
        for _ in range(reps + 1):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tickets = []
            for tr, sh, d_paths, d_vm, _ in work:
                if exchange:
                    tickets.append(sh.submit_device(d_paths, None, d_vm, False, B_total=per_track))
                else:
                    _, v = sh.solver.alloc_device_outputs(d_paths.shape[0], fields) if "v" not in sh.__dict__ else (None, sh.v)
                    sh.v = v
                    sh.solver.solve_device(d_paths, None, d_vm, False, out=v)
            if exchange:
                for (tr, sh, *_), t in zip(work, tickets):
                    sh.wait(t)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t.item()))
        return float(np.min(ms[1:])), tickets

    solve_ms, _ = timed(False, 2)
    both_ms, tickets = timed(True, 2)
    line = None
    if rank == 0:
        parity, solved, n = [], 0.0, per_track * len(work)
        for (tr, sh, _, _, (cl, idx, lat, psi, vm)), t in zip(work, tickets):
            views = sh.wait(t)
            got = {k: torch.cat(v, dim=0).cpu().numpy() for k, v in views.items()}
            sub = np.random.default_rng(1).choice(per_track, 64, replace=False)
            p = tracks.make_instances(cl, idx[sub], 50, lat[sub], psi[sub])
            chk = _spot_check(config_kwargs(tr, 50), p, vm[sub], {k: g[sub] for k, g in got.items()}, sample=64)
            parity.append(dict(chk, track=tr))
            solved += float((got["status"] == 1).sum())
        line = {"instances": n, "tracks": len(work), "n_gpus": world, "scaling": "strong",
                "instances_per_gpu": n // world, "transport": work[0][1].transport,
                "solve_only": {"value": n / (solve_ms * 1e-3), "unit": UNIT, "ms": solve_ms},
                "solve_plus_exchange": {"value": n / (both_ms * 1e-3), "unit": UNIT, "ms": both_ms,
                                        "bytes_to_rank0": int(sum(sh._cap for _, sh, *_ in work)) * (world - 1)},
                "solved_frac": solved / n,
                "parity_all_ok": all(c["iters_equal"] and c["status_equal"] and (c["max_abs_dcontrols"] or 0) < 1e-6 for c in parity),
                "parity": parity, "timed": "max over ranks of CUDA events around all 7 tracks, best of 2 after a warm-up pass"}
    barrier()
    return line


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from ac_mpc_b200 import BatchedMPC, _capi, fp64_peak_tflops, sharding, tracks
    from ac_mpc_b200.control import build_mpc
    from ac_mpc_b200.sharded import NBUF, ShardedMPC

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, H, K, W = args.batch, args.horizon, args.steps, max(args.warmup, 3)
    kw = config_kwargs(args.track, H)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # before the warm-up: the timed region is only ~12 ms long
    paths, vmax = workload(args.track, B, H, rank)
    sh = ShardedMPC(_capi.default_config(**kw), fields=BENCH_FIELDS, device=local_rank, transport=args.transport)
    mpc = sh.solver
    # Timed steps run back to back with NO L2 flush in between; instead step i reads copy i % NROT of the inputs, and the
    # NROT copies together (plus the rotating output buffers) are larger than the 126 MB L2
    bytes_in = B * (3 * H + 1) * 8
    NROT = max(3, -(-(160 << 20) // bytes_in))
    d_paths_rot = torch.from_numpy(paths).to(dev).unsqueeze(0).repeat(NROT, 1, 1, 1).contiguous()
    d_vmax_rot = torch.from_numpy(vmax).to(dev).unsqueeze(0).repeat(NROT, 1).contiguous()
    d_paths, d_vmax = d_paths_rot[0], d_vmax_rot[0]
    _FLUSH["t"] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # the legs flush explicitly (> 126 MB L2)
    main = torch.cuda.current_stream()

    def timed_pass(n_steps, first):
        """n_steps pipelined steps back to back + the completion of the LAST exchange, between ONE pair of events on the
        launching stream: every kernel and every exchange of the pass lies inside [start, end].  The exchange of step i
        overlaps the kernels of step i + 1 (ShardedMPC); per-step events give the step-time distribution."""
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        start.record()
        t = None
        for i in range(n_steps):
            k = (first + i) % NROT
            t = sh.submit_device(d_paths_rot[k], None, d_vmax_rot[k], False, B_total=B * world)
            marks[i].record()
        views = sh.wait(t)                 # the launching stream waits for the exchange of the last step
        end.record()
        return start, marks, end, views

    # warm-up = the EXACT timed path (same streams, all NBUF output buffers, pipelined), at least NBUF + 1 steps
    n_warm = max(W, NBUF + 1)
    timed_pass(n_warm, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    mpc.set_profiling(True)        # events around each of the two kernels, on the launching stream

    def aligned_start():
        """Ranks leave an NCCL barrier up to a few hundred microseconds apart (and the first broadcast of a process can take
        tens of milliseconds), while rank 0's timed region ends only when the LAST rank's results have landed: with 12 ms
        regions that skew is 1-3 % of the figure, or far more.  All ranks of one node share CLOCK_MONOTONIC, so rank 0
        publishes a deadline well ahead and everybody starts on it.  Returns how late THIS rank was (ns, 0 = on time)."""
        if world == 1:
            return 0
        dist.barrier()                      # everybody is HERE before the deadline is drawn
        go = torch.tensor([time.monotonic_ns() + 100_000_000], dtype=torch.int64, device=dev)
        dist.broadcast(go, src=0)
        go = int(go.item())
        torch.cuda.synchronize()
        late = max(time.monotonic_ns() - go, 0)
        while time.monotonic_ns() < go:
            pass
        return late

    aligned_start()                         # also the warm-up of the broadcast itself
    start_late_us, attempts = 0.0, 0
    for attempts in range(1, 4):            # a pass whose start was not aligned is measured again (it costs 12 ms)
        mpc.collect_kernel_ms()
        late = aligned_start()
        sampler.mark()
        start, marks, end, views = timed_pass(K, n_warm)
        torch.cuda.synchronize()
        sampler.mark()
        lt = torch.tensor([float(late)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        start_late_us = float(lt.item()) / 1e3
        if start_late_us < 100.0:
            break
    per_kernel = mpc.collect_kernel_ms()
    mpc.set_profiling(False)
    if world > 1:
        dist.barrier()
    total_ms = start.elapsed_time(end)
    edges = [start] + marks
    step_ms = [edges[i].elapsed_time(edges[i + 1]) for i in range(K)]
    drain_ms = marks[-1].elapsed_time(end)
    kernel_ms = float(np.median(step_ms))
    t = torch.tensor([total_ms, max(step_ms), float(np.median(step_ms)), -min(step_ms), drain_ms], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, step_max, step_med, step_min, drain_ms = float(t[0]), float(t[1]), float(t[2]), -float(t[3]), float(t[4])
    # per-rank kernel times (the library's own events): do the ranks whose stores cross NVLink run slower?
    nl_ = max(per_kernel["launches"], 1)
    kr = torch.tensor([per_kernel["speed_ms"] / nl_, per_kernel["control_ms"] / nl_], dtype=torch.float64, device=dev)
    kr_all = [torch.zeros_like(kr) for _ in range(world)]
    if world > 1:
        dist.all_gather(kr_all, kr)
    else:
        kr_all = [kr]
    kernels_per_rank = [[round(float(x), 5) for x in k.tolist()] for k in kr_all]

    # ---- verification of what the exchange delivered (not timed): per-rank checksums + an oracle sample of another rank
    local = {k: v.cpu().numpy() for k, v in sh.last_local_views().items()}     # what THIS rank's kernels wrote last
    csum = torch.tensor([float(local["controls"].view(np.int64).sum() % (1 << 52)), float(local["iters"].sum()),
                         float(local["status"].sum())], dtype=torch.float64, device=dev)
    sums = [torch.zeros_like(csum) for _ in range(world)]
    if world > 1:
        dist.all_gather(sums, csum)
    else:
        sums = [csum]
    gather_verified = None
    if rank == 0:
        ok = True
        for r in range(world):
            c = views["controls"][r].cpu().numpy()
            got = [float(c.view(np.int64).sum() % (1 << 52)), float(views["iters"][r].sum().item()),
                   float(views["status"][r].sum().item())]
            ok = ok and got == [float(x) for x in sums[r].tolist()]
        r_chk = world - 1                              # a shard rank 0 did not compute
        p_chk, v_chk = workload(args.track, B, H, r_chk)
        chk = _spot_check(kw, p_chk, v_chk, {k: views[k][r_chk].cpu().numpy() for k in ("controls", "iters", "status")},
                          sample=48, seed=2)
        gather_verified = bool(ok and chk["iters_equal"] and chk["status_equal"] and (chk["max_abs_dcontrols"] or 0) < 1e-6)
        gather_check = {"checksums_equal_all_ranks": bool(ok), "oracle_sample_of_rank": r_chk, **chk}
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public host API with pinned host buffers -----------------------------------------
    bin_, bout = bytes_per_solve(H, BENCH_FIELDS)
    api = build_mpc(tracks.racing_config(args.track, H), VEH, device=local_rank)
    if world == 1:
        h_paths = torch.from_numpy(paths).pin_memory()
        h_vmax = torch.from_numpy(vmax).pin_memory()
        h_out = api._batched().alloc_host_outputs(B, BENCH_FIELDS, pinned=True)
        np_paths, np_vmax = h_paths.numpy(), h_vmax.numpy()
        e2e_call = lambda: api.get_control_batch(np_paths, None, np_vmax, False, out=h_out)
        e2e_api = "SpatialMPC.get_control_batch -> acmpc_solve_batch_host"
    else:
        e2e_res = {}
        hp_, hv_ = torch.from_numpy(paths).pin_memory(), torch.from_numpy(vmax).pin_memory()
        np_paths, np_vmax = hp_.numpy(), hv_.numpy()

        def e2e_call():
            e2e_res["out"] = sh.solve(np_paths, None, np_vmax, False, local=True, B_total=B * world)
        e2e_api = ("ShardedMPC.solve(local shard, host arrays), synchronous: H2D per rank -> kernels -> D2H of the shard over "
                   "the rank's own PCIe link into the shared host segment rank 0 reads")
    for _ in range(W):
        e2e_call()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_call()
    e2e_s = time.perf_counter() - t0
    e2e_sync = None
    if world > 1:
        # streams of batches at N > 1: every rank runs its own copy-in / compute / copy-out pipeline and delivers its shard
        # over its OWN PCIe link into the host segment rank 0 reads (ShardedMPC.submit_host / wait_host)
        t_ = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        e2e_sync = {"value": world * B * K / float(t_.item()), "unit": UNIT, "api": e2e_api}
        from ac_mpc_b200.sharded import ShmUnavailable

        try:
            for rep in range(2):
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                prev = None
                for _ in range(K):
                    tk = sh.submit_host(np_paths, None, np_vmax, False, B_total=B * world)
                    if prev is not None:
                        sh.wait_host(prev)
                    prev = tk
                e2e_res["out"] = sh.wait_host(prev)
                torch.cuda.synchronize()
                e2e_s = time.perf_counter() - t0
            pipelined = True
        except ShmUnavailable:                     # raised by every rank together: keep the synchronous figure
            pipelined = False
        if pipelined:
            e2e_api = ("ShardedMPC.submit_host / wait_host: per rank H2D from pinned memory -> kernels -> D2H of its shard over its "
                       "own PCIe link into one shared host segment that rank 0 reads; consecutive steps overlapped (depth 2)")
    if world == 1:
        # The synchronous call above returns when its results are in host memory: PCIe time sits next to kernel time.  A
        # stream of batches (sweep, replay) goes through the public pipeline API instead -- every batch still pays its own
        # H2D from pinned memory and its own D2H, overlapped with the neighbouring batches' kernels.
        e2e_sync = {"value": B * K / e2e_s, "unit": UNIT, "api": e2e_api}
        pipe = api._batched().pipeline(B, BENCH_FIELDS, depth=2)
        for rep in range(2):                       # the first pass is the warm-up of the exact timed path
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            prev = None
            for _ in range(K):
                tk = pipe.submit(np_paths, None, np_vmax, False)
                if prev is not None:
                    h_pipe = pipe.wait(prev)
                prev = tk
            h_pipe = pipe.wait(prev)
            e2e_s = time.perf_counter() - t0
        pipe_same = all(np.array_equal(h_pipe[k], h_out[k]) for k in ("controls", "status", "iters"))
        e2e_api = ("BatchedMPC.pipeline(B).submit / wait (ac_mpc_b200/pipeline.py): H2D from pinned memory, both kernels and "
                   "the D2H of every batch inside the timed region, consecutive batches overlapped (depth 2)")
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    if world == 1:
        same = bool(pipe_same and all(np.array_equal(views[k][0].cpu().numpy(), h_out[k]) for k in ("controls", "status", "iters")))
        iters, rhou, status = h_out["iters"], h_out["rho_updates"], h_out["status"]
    else:
        out = e2e_res.get("out")
        same = None if rank != 0 else all(np.array_equal(torch.cat(views[k], 0).cpu().numpy(), out[k])
                                          for k in ("controls", "status", "iters"))
        iters, rhou, status = local["iters"], local["rho_updates"], local["status"]

    # configs[4] runs on every rank (strong scaling over the N GPUs); the other legs on rank 0 at N = 1
    fp64_peak = fp64_peak_tflops(local_rank)
    all_tracks = None
    if not args.no_legs and args.all_tracks_instances > 0:
        all_tracks = leg_all_tracks(args, rank, local_rank, world, fp64_peak)
    if rank != 0:
        return
    flops_speed, flops_control = flops_per_batch(H, iters, rhou)
    flops = flops_speed + flops_control
    peaks, how = measured_peaks()
    k_s = kernel_ms * 1e-3
    nl = max(per_kernel["launches"], 1)
    control_ms, speed_ms = per_kernel["control_ms"] / nl, per_kernel["speed_ms"] / nl
    achieved_tf = flops_control / (control_ms * 1e-3) / 1e12
    hbm_gbs = B * (bin_ + bout) / k_s / 1e9
    traffic = recorded_traffic()
    launch_info = mpc.launch_info()

    legs = {}
    lat = {}
    if world == 1:
        # batch-1 latency (BASELINE configs[0]: Monza, waypoint 0, no perturbation), host buffers in and out.
        # cold = a fresh solve per call; warm = the reference's real call pattern, get_control on ONE object.
        from oracle import port

        cl = tracks.synthetic_centreline(args.track)
        p1 = tracks.make_instances(cl, [0], H)[0]
        seq = tracks.make_instances(cl, (np.arange(max(args.latency_reps, 40) + 20) * 4) % cl.shape[0], H)   # 2 m per step
        o1 = api._batched().alloc_host_outputs(1, None, pinned=True)
        lat["latency_b1_p50_ms"] = p50(lambda i: api.get_control_batch(p1[None], None, None, False, out=o1), args.latency_reps)
        lat["latency_b1_warm_p50_ms"] = p50(lambda i: api.get_control(seq[i]), args.latency_reps)
        pc = port.PortMPC(port.default_config(**kw))
        lat["cpu_latency_b1_p50_ms"] = p50(lambda i: pc.step(p1, 0.0, None, False, warm=False), 40)
        pw = port.PortMPC(port.default_config(**kw))
        lat["cpu_latency_b1_warm_p50_ms"] = p50(lambda i: pw.step(seq[i], 0.0, None, False, warm=True), 40)
        lat["latency_note"] = ("cold: one fresh solve per call; warm: consecutive get_control calls on one object, the car "
                               "advancing 2 m per call (OSQP warm start + carried rho, as the reference runs); cpu = the "
                               "oracle's C port without the reference's Python glue (README: 7-8 ms per call measured)")
    if world == 1 and not args.no_legs:
        legs["configs2_nordschleife_sweep"] = leg_nordschleife(fp64_peak, local_rank)
        legs["configs3_spa_horizon_sweep"] = [leg_horizon("spa", h, 16384, fp64_peak, local_rank) for h in (20, 40, 80)]
        legs["horizon_100_monza"] = leg_horizon("monza", 100, 4096, fp64_peak, local_rank)
    legs["configs4_all_tracks_1m"] = all_tracks

    cpu = None
    if world == 1:
        cores = len(os.sched_getaffinity(0))
        cpu_rate, cpu_done, cpu_dt = cpu_port_rate(kw, paths, vmax, cores, args.cpu_seconds, 16)
        from oracle import osqp_select

        cpu = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port", "solver": osqp_select.PORT_LABEL,
               "osqp_wheel_importable": osqp_select.info().real,
               "sample": f"{cpu_done} solves of the same batch in {cpu_dt:.1f} s, cold start, "
                         f"oracle/acmpc_port.c on {cores} pthreads"}

    # start-up path and device-resident sweep (scope rows 8f-1, 8f-4), N = 1 only
    map_line = sweep_line = None
    if world == 1 and not args.no_map_profile:
        from oracle import port

        cl = tracks.synthetic_centreline(args.track)
        trk = tracks.map_track(cl)
        mp_c = tracks.MAP_PROFILE[args.track]
        spc = tracks.racing_config(args.track)["speed_profile_constraints"]
        best = None
        for _ in range(3):
            _, xg, mi = api._batched().track_speed_profile(trk, spc["v_max"], mp_c["ay_max"], mp_c["a_min"])
            if best is None or mi["kernel_ms"] < best["kernel_ms"]:
                best = mi
        t0 = time.perf_counter()
        xo, io = port.map_speed_profile(port.construct_waypoints(trk), spc, mp_c["ay_max"], mp_c["a_min"])
        cpu_s = time.perf_counter() - t0
        map_line = {"track": args.track, "waypoints": int(xg.shape[0]), "ctas": best["ctas"], "status": best["status_str"],
                    "admm_iterations": best["iters"], "kernel_ms": best["kernel_ms"],
                    "us_per_iteration": 1e3 * best["kernel_ms"] / max(best["iters"], 1),
                    "cpu_port_s": cpu_s, "cpu_iterations": int(io.iter), "max_abs_diff_m_s": float(np.abs(xo - xg).max()),
                    "api": "acmpc_track_speed_profile_host (construct_waypoints + compute_map_speed_profile, "
                           "spatial_mpc.py:60-87,125-154)"}
    if world == 1 and not args.no_sweep:
        cl_s = tracks.synthetic_centreline(args.track)
        rng = np.random.default_rng(1 + rank)
        s_idx = rng.integers(0, cl_s.shape[0], B).astype(np.int32)
        s_lat, s_psi, s_vmax = rng.uniform(-2.0, 2.0, B), rng.uniform(-0.1, 0.1, B), rng.uniform(20.0, 84.0, B)
        h_in = [torch.from_numpy(a).pin_memory() for a in (s_idx, s_lat, s_psi, s_vmax)]
        d_in = [torch.empty_like(a, device=dev) for a in h_in]
        d_cl = torch.from_numpy(cl_s).to(dev)
        d_sp = torch.empty((B, H, 3), dtype=torch.float64, device=dev)
        _, sv = mpc.alloc_device_outputs(B, ["controls", "status"])
        h_ctl = torch.empty((B, 2, H - 1), dtype=torch.float64).pin_memory()
        h_st = torch.empty(B, dtype=torch.int32).pin_memory()

        def sweep_step():
            for d_a, h_a in zip(d_in, h_in):
                d_a.copy_(h_a, non_blocking=True)
            mpc.extract_paths_device(d_cl, d_in[0], d_in[1], d_in[2], out=d_sp)
            mpc.solve_device(d_sp, None, d_in[3], False, out=sv)
            h_ctl.copy_(sv["controls"], non_blocking=True)
            h_st.copy_(sv["status"], non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            sweep_step()
        t0 = time.perf_counter()
        for _ in range(K):
            sweep_step()
        sweep_s = time.perf_counter() - t0
        sweep_line = {"value": B * K / sweep_s, "unit": UNIT, "h2d_bytes_per_step": B * 28,
                      "d2h_bytes_per_step": B * (2 * (H - 1) * 8 + 4),
                      "max_abs_path_diff_m": float(np.abs(d_sp.cpu().numpy() - paths).max()),
                      "controls_equal_timed_steps": bool(np.abs(h_ctl.numpy() - h_out["controls"]).max() < 1e-3),
                      "api": "acmpc_extract_paths_device -> acmpc_solve_batch_device, centre line resident in HBM"}

    line = {
        "metric": METRIC, "value": world * B * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(W, NBUF + 1), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world),
        "run": {"l2": f"no flush: step i reads input copy i % {NROT} ({NROT} x {bytes_in / 1e6:.1f} MB of inputs + {NBUF} rotating "
                      "output buffers > 126 MB L2), steps run back to back inside ONE pair of CUDA events",
                "parallelism": (f"instances sharded over {world} GPUs (ShardedMPC), one exchange per step to rank 0 via "
                                f"'{sh.transport}' transport, overlapped with the next step's kernels; the last one timed on its own"
                                if world > 1 else "single GPU"),
                "transport": sh.transport, "transport_note": sh.transport_note,
                "warmup": "the exact timed path (same streams and buffers, pipelined)",
                "step_ms": {"min": step_min, "median": step_med, "max": step_max, "last_exchange": drain_ms,
                            "over": "max over ranks; a step = the interval between consecutive step-end events"},
                "kernel_ms_per_rank_speed_control": kernels_per_rank,
                "aligned_start": {"max_rank_lateness_us": start_late_us, "attempts": attempts}},
        "e2e": {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": world * B * bin_,
                "d2h_bytes_per_step": world * B * bout, "api": e2e_api, "equals_device_path": same,
                "synchronous_call": e2e_sync},
        "gpu_launches": K * launch_info["launches"],
        "gather_verified": gather_verified, "gather_check": gather_check,
        "kernel_ms": kernel_ms,
        "kernels": {"acmpc_speed_kernel": {"ms": speed_ms, "flops_per_launch": flops_speed,
                                           "tflops": flops_speed / (speed_ms * 1e-3) / 1e12},
                    "acmpc_control_kernel": {"ms": control_ms, "flops_per_launch": flops_control,
                                             "tflops": flops_control / (control_ms * 1e-3) / 1e12},
                    "timed": f"CUDA events on the launching stream around each kernel, mean of {nl} launches"},
        "roofline": {"kernel": "acmpc_control_kernel", "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak, "traffic": traffic,
                     "whole_step": {"achieved": flops / k_s / 1e12, "frac": flops / k_s / 1e12 / fp64_peak},
                     "peak_source": "DFMA micro-benchmark in this run (acmpc_fp64_peak_tflops); "
                                    "MEASURED_PEAKS.json has no FP64 figure",
                     "flops_per_solve": flops / B,
                     "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm_gbs / peaks["hbm_gbs"], "of": how, "bytes_per_solve": bin_ + bout}},
        "cpu_baseline": cpu,
        **lat,
        "iters_mean": [float(iters[:, 0].mean()), float(iters[:, 1].mean())],
        "solved_frac": float((status == 1).mean()),
        "launch": launch_info, "clocks": clocks, "legs": legs, "map_speed_profile": map_line, "track_sweep": sweep_line,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU per step")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: how the results reach rank 0 (peer-mapped stores over NVLink / one NCCL gather per step)")
    ap.add_argument("--track", default="monza")
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--no-sweep", action="store_true", help="skip the device-resident track-sweep leg")
    ap.add_argument("--no-map-profile", action="store_true", help="skip the whole-track speed-profile leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the BASELINE configs[2..4] / H = 100 legs")
    ap.add_argument("--all-tracks-instances", type=int, default=1 << 20, help="size of the configs[4] leg (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=4.0, help="wall-clock budget of the cpu_baseline leg")
    ap.add_argument("--latency-reps", type=int, default=200)
    ap.add_argument("--reference-src", default="/root/reference/src", help="--impl reference: ac-mpc sources, if present")
    ap.add_argument("--ref-python", action="store_true",
                    help="--impl reference: run the reference's Python even without a real osqp wheel (on the stand-in)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=__import__("torch").device("cuda", local_rank))
    elif args.gpus > 1:
        sys.exit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    try:
        run_b200(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
