#!/usr/bin/env python
"""Where does a control-kernel warp spend its cycles?  Builds an instrumented copy of the library
(-DACMPC_PHASE_TIMING: lane 0 adds clock64() deltas per phase to a device array), runs the bench workload and prints the
share of each phase.  Experiment tooling: the product build has no clocks in it.

    python tools/phase_timing.py build        # CPU container: nvcc -> tools/_exp/libacmpc_phase.so
    python tools/phase_timing.py run [B]      # GPU box
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "_exp", "libacmpc_phase.so")
NAMES = ["waypoints", "setup (assembly + Ruiz)", "first factorisation", "ADMM iterations", "checks (+ refactor)", "-",
         "solve tail (obj, warm)", "outputs",
         "SPEED: staging + waypoints", "SPEED: assembly + Ruiz", "SPEED: first factorisation", "SPEED: ADMM iterations",
         "SPEED: checks (+ refactor)", "SPEED: solve tail", "SPEED: outputs"]

if sys.argv[1] == "build":
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
                    "-fPIC", "-shared", "-DACMPC_PHASE_TIMING", "-o", SO,
                    os.path.join(ROOT, "ac_mpc_b200", "csrc", "acmpc_b200.cu")], check=True)
    print("built", SO)
else:
    os.environ["ACMPC_B200_LIB"] = SO
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    from ac_mpc_b200 import BatchedMPC, _capi, tracks

    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    paths, vmax = tracks.perturbed_batch("monza", B, horizon=H, seed=1)
    mpc = BatchedMPC(_capi.default_config(horizon=H), device=0)
    dp, dv = torch.from_numpy(paths).cuda(), torch.from_numpy(vmax).cuda()
    _, views = mpc.alloc_device_outputs(B, ["controls", "status", "iters"])
    lib = _capi.load()
    buf = (C.c_ulonglong * 16)()
    for rep in range(3):
        mpc.solve_device(dp, None, dv, False, out=views)
        torch.cuda.synchronize()
        lib.acmpc_exp_phase_cycles(buf)
    cyc = np.array(list(buf), dtype=np.float64)
    for lo, hi, what in ((0, 8, "control kernel"), (8, 16, "speed kernel")):
        tot = cyc[lo:hi].sum()
        print(f"B={B} H={H} {what}: {tot / B:.0f} cycles per instance (lane-0 clock64 deltas, summed over phases)")
        for k in range(lo, min(hi, len(NAMES))):
            if cyc[k] > 0:
                print(f"  {NAMES[k]:30s} {cyc[k] / B:10.0f} cycles  {100 * cyc[k] / tot:5.1f} %")
    it = views["iters"].cpu().numpy()
    print("  mean iterations (speed, control)", it[:, 0].mean(), it[:, 1].mean())
