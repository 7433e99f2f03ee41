#!/usr/bin/env python
"""bench.py -- batched MPC solves/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                     # the CPU arm (oracle port, all host threads)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1, one rank per GPU

A "step" is one pass of the hot path (SpatialMPC.get_control: waypoints + speed-profile QP +
linearise/assemble + control QP + unpack/rollout/cost) over one batch of synthetic instances:
BASELINE.json configs[1] = Monza racing block (H = 50), 4096 perturbed initial states per GPU.
At N > 1 every rank solves its own 4096-instance shard (weak scaling, no data-path collective) and every
step has the single NCCL collective that brings the packed outputs to rank 0 (--collective all_gather: to every rank).
By default (--gather pipelined) the collective of step i runs on a side stream while step i + 1 computes, from the
other of two output buffers; it starts after the start event of the step it overlaps and ends before that step's end
event, and the collective of the last step is a timed region of its own -- all K collectives are inside the timed time.
--gather in_step keeps each collective inside its own step.

One JSON line on stdout (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same metric through the reference-facing API (SpatialMPC.get_control_batch -> C ABI host
entry point) with pinned HOST buffers, H2D/D2H inside the timed region.
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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1])), pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


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


# ------------------------------------------------------------------------------------------------
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


def run_reference(args, rank: int):
    """--impl reference: the CPU implementation of the path on the box's host cores.  The reference is
    pure Python over the absent `osqp` wheel and /root/reference does not exist on the GPU box, so this
    arm times the oracle's C port (faster than the reference's Python glue: a conservative baseline)."""
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    kw = config_kwargs(args.track, args.horizon)
    sample = min(args.batch, args.ref_sample)
    paths, vmax = workload(args.track, sample, args.horizon, 0)
    from oracle import port

    cfg = port.default_config(**kw)
    for _ in range(args.warmup):
        port.solve_batch(cfg, paths, None, vmax, False, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.solve_batch(cfg, paths, None, vmax, False, nthreads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.track} racing block H={args.horizon}, {args.batch} perturbed initial states "
                               "(BASELINE configs[1])", "track": args.track, "horizon": args.horizon,
                   "batch_per_gpu": args.batch, "cold_start": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} of the {args.batch} instances per step, {args.steps} steps, "
                                   f"oracle/acmpc_port.c on {threads} pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from ac_mpc_b200 import BatchedMPC, _capi, fp64_peak_tflops, tracks
    from ac_mpc_b200.control import build_mpc

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, H, K, W = args.batch, args.horizon, args.steps, args.warmup
    kw = config_kwargs(args.track, H)
    paths, vmax = workload(args.track, B, H, rank)
    mpc = BatchedMPC(_capi.default_config(**kw), device=local_rank)
    d_paths = torch.from_numpy(paths).to(dev)
    d_vmax = torch.from_numpy(vmax).to(dev)
    packed, views = mpc.alloc_device_outputs(B, BENCH_FIELDS)
    gathered = torch.empty(world * packed.numel(), dtype=torch.uint8, device=dev) if world > 1 else None
    # the one collective of the path: the packed results of every rank end up on rank 0 ("gather", the default:
    # rank 0 is where the caller lives) or on every rank ("all_gather")
    slots = list(gathered.view(world, -1).unbind(0)) if (world > 1 and rank == 0) else None

    def collect():
        if args.collective == "gather":
            dist.gather(packed, slots, dst=0)
        else:
            dist.all_gather_into_tensor(gathered, packed)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step():
        mpc.solve_device(d_paths, None, d_vmax, False, out=views)
        if world > 1:   # the one collective of the path: final gather of the packed results over NVLink
            collect()

    for _ in range(max(W, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    mpc.set_profiling(True)        # events around each of the two kernels, on the launching stream
    pipelined = world > 1 and args.gather == "pipelined"
    if pipelined:
        # The gather of step i runs on a side stream WHILE step i + 1 computes (double-buffered outputs): every gather
        # starts after the start event of the step it overlaps and ends before that step's end event, so nothing
        # timed hides in an L2 flush; the gather of the last step gets a timed region of its own (ev[K]).
        main, side = torch.cuda.current_stream(), torch.cuda.Stream(device=dev)
        bufs = [(packed, views), mpc.alloc_device_outputs(B, BENCH_FIELDS)]
        solved = [torch.cuda.Event() for _ in range(2)]

        def gather_async(buf, after):
            side.wait_event(after)
            side.wait_event(solved[buf])
            with torch.cuda.stream(side):
                if args.collective == "gather":
                    dist.gather(bufs[buf][0], slots, dst=0)
                else:
                    dist.all_gather_into_tensor(gathered, bufs[buf][0])

        for i in range(K):
            flush.fill_(i & 0xFF)                  # L2 flush between timed iterations (outside the events)
            ev[i][0].record()
            if i > 0:
                gather_async((i - 1) & 1, ev[i][0])
            kev[i][0].record()
            mpc.solve_device(d_paths, None, d_vmax, False, out=bufs[i & 1][1])
            kev[i][1].record()
            solved[i & 1].record()
            main.wait_stream(side)                 # the step ends when its kernels AND the overlapped gather are done
            ev[i][1].record()
        ev[K][0].record()
        gather_async((K - 1) & 1, ev[K][0])
        main.wait_stream(side)
        ev[K][1].record()
        views = bufs[(K - 1) & 1][1]
    else:
        for i in range(K):
            flush.fill_(i & 0xFF)                  # L2 flush between timed iterations (outside the events)
            ev[i][0].record()
            kev[i][0].record()
            mpc.solve_device(d_paths, None, d_vmax, False, out=views)
            kev[i][1].record()
            if world > 1:
                collect()
            ev[i][1].record()
        ev.pop()
    torch.cuda.synchronize()
    per_kernel = mpc.collect_kernel_ms()
    mpc.set_profiling(False)
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / K
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the reference-facing API with pinned host buffers --------------------
    veh = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                         "max_steering_angle": lambda self: 0.30})()
    api = build_mpc(tracks.racing_config(args.track, H), veh, device=local_rank)
    h_paths = torch.from_numpy(paths).pin_memory()
    h_vmax = torch.from_numpy(vmax).pin_memory()
    h_out = api._batched().alloc_host_outputs(B, BENCH_FIELDS, pinned=True)
    np_paths, np_vmax = h_paths.numpy(), h_vmax.numpy()
    for _ in range(max(W, 3)):
        api.get_control_batch(np_paths, None, np_vmax, False, out=h_out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        api.get_control_batch(np_paths, None, np_vmax, False, out=h_out)     # synchronous: H2D + kernel + D2H
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    bin_, bout = bytes_per_solve(H, BENCH_FIELDS)

    # parity spot check of what was just timed (not in the timed region)
    solved = float((h_out["status"] == 1).mean())
    same = all(np.array_equal(views[k].cpu().numpy(), h_out[k]) for k in ("controls", "status", "iters"))
    if rank != 0:
        return
    iters, rhou = h_out["iters"], h_out["rho_updates"]
    flops_speed, flops_control = flops_per_batch(H, iters, rhou)
    flops = flops_speed + flops_control
    peaks, how = measured_peaks()
    fp64_peak = fp64_peak_tflops(local_rank)
    k_s = kernel_ms * 1e-3
    nl = max(per_kernel["launches"], 1)
    control_ms, speed_ms = per_kernel["control_ms"] / nl, per_kernel["speed_ms"] / nl
    # roofline of the dominant kernel (the control kernel): its algorithmic flops / its own device time
    achieved_tf = flops_control / (control_ms * 1e-3) / 1e12
    hbm_gbs = B * (bin_ + bout) / k_s / 1e9
    traffic = recorded_traffic()

    # batch-1 latency (BASELINE configs[0]: Monza, waypoint 0, no perturbation), host buffers in and out.
    # cold = a fresh solve per call (what `value` measures per instance); warm = the reference's real call
    # pattern, get_control on ONE object whose OSQP state persists (here: the handle's warm-start record).
    cl = tracks.synthetic_centreline(args.track)
    p1 = tracks.make_instances(cl, [0], H)[0]
    seq = tracks.make_instances(cl, (np.arange(max(args.latency_reps, 40) + 20) * 4) % cl.shape[0], H)   # 2 m per step

    def p50(fn, n):
        ts = []
        for i in range(n + 20):
            t0 = time.perf_counter()
            fn(i)
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts[20:]) * 1e3)

    o1 = api._batched().alloc_host_outputs(1, None, pinned=True)
    lat_ms = p50(lambda i: api.get_control_batch(p1[None], None, None, False, out=o1), args.latency_reps)
    lat_warm_ms = p50(lambda i: api.get_control(seq[i]), args.latency_reps)

    # CPU baseline: oracle port on the box's host cores, bounded sample of the same workload
    cores = len(os.sched_getaffinity(0))
    cpu_rate, cpu_done, cpu_dt = cpu_port_rate(kw, paths, vmax, cores, args.cpu_seconds, 16)
    from oracle import port

    pc = port.PortMPC(port.default_config(**kw))
    cpu_lat_ms = p50(lambda i: pc.step(p1, 0.0, None, False, warm=False), 40)
    pw = port.PortMPC(port.default_config(**kw))
    cpu_lat_warm_ms = p50(lambda i: pw.step(seq[i], 0.0, None, False, warm=True), 40)

    launch_info = mpc.launch_info()
    # start-up path (scope row 8f-1): the whole-track speed profile of the same track, one cooperative launch,
    # next to the oracle's C OSQP port on one host core (solve only, its setup excluded).  Outside the timed steps.
    map_line = None
    if not args.no_map_profile:
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

    # track sweep (scope row 8f-4): the centre line stays resident, a step uploads (index, lateral offset, heading
    # offset, v_max) = 28 bytes per instance, the paths are built on the device (acmpc_extract_paths_device) and only
    # controls + status come back.  Same instances as the timed steps (perturbed_batch's draws).  Outside the timed steps.
    sweep_line = None
    if not args.no_sweep:
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
        "warmup": max(W, 3), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.track} racing block H={H}, {B} perturbed initial states per GPU "
                               "(BASELINE configs[1]), cold start per instance",
                   "track": args.track, "horizon": H, "batch_per_gpu": B, "global_batch": world * B,
                   "l2": "flushed between timed steps (256 MiB fill outside the events)",
                   "parallelism": f"instances sharded over {world} GPU(s), one final NCCL {args.collective} per step ({args.gather})" if world > 1
                   else "single GPU", "osqp": "eps_abs=eps_rel=1e-3, check 25, adaptive rho interval 50"},
        "e2e": {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * bin_,
                "d2h_bytes_per_step": B * bout, "api": "SpatialMPC.get_control_batch -> acmpc_solve_batch_host"},
        "gpu_launches": K * mpc.launch_info()["launches"],
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
        "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cpu_done} solves of the same batch in {cpu_dt:.1f} s, cold start, "
                                   f"oracle/acmpc_port.c on {cores} pthreads"},
        "latency_b1_p50_ms": lat_ms, "cpu_latency_b1_p50_ms": cpu_lat_ms,
        "latency_b1_warm_p50_ms": lat_warm_ms, "cpu_latency_b1_warm_p50_ms": cpu_lat_warm_ms,
        "latency_note": "cold: one fresh solve per call; warm: consecutive get_control calls on one object, the car "
                        "advancing 2 m per call (OSQP warm start + carried rho, as the reference runs); cpu = the "
                        "oracle's C port without the reference's ~6 ms of Python glue per call",
        "iters_mean": [float(iters[:, 0].mean()), float(iters[:, 1].mean())],
        "solved_frac": solved, "device_equals_host_path": bool(same),
        "launch": launch_info, "clocks": clocks, "map_speed_profile": map_line, "track_sweep": sweep_line,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU per step")
    ap.add_argument("--gather", default="pipelined", choices=["pipelined", "in_step"],
                    help="N > 1: the final collective of step i overlaps the kernels of step i + 1 (side stream, double-"
                         "buffered outputs; the last one is timed on its own) / runs inside its own step")
    ap.add_argument("--collective", default="gather", choices=["gather", "all_gather"],
                    help="N > 1: where the packed results go at the end of a step (rank 0 / every rank)")
    ap.add_argument("--track", default="monza")
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--no-sweep", action="store_true", help="skip the device-resident track-sweep leg")
    ap.add_argument("--no-map-profile", action="store_true", help="skip the whole-track speed-profile leg")
    ap.add_argument("--cpu-seconds", type=float, default=4.0, help="wall-clock budget of the cpu_baseline leg")
    ap.add_argument("--ref-sample", type=int, default=1024, help="instances per step of --impl reference")
    ap.add_argument("--latency-reps", type=int, default=200)
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
