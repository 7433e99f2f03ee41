#!/usr/bin/env python
"""tools/track_prep_bench.py -- the track-side kernels (track_prep.cuh) at sizes where HBM matters.  Prints one JSON
line with the host-call times; run it under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
for the per-kernel device times and traffic (profiles/r1o_track_launches.csv), or plainly for the extraction kernel's
CUDA-event time (the only one with a device entry point)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ac_mpc_b200 import BatchedMPC, _capi, tracks  # noqa: E402


def main():
    import torch

    s = BatchedMPC(_capi.default_config(), device=0)
    rng = np.random.default_rng(0)
    out = {}

    M = 4_000_000
    line = np.cumsum(rng.normal(0, 0.3, (M, 2)), axis=0)
    dup = rng.random(M) < 0.2
    dup[0] = False
    line[dup] = line[np.flatnonzero(dup) - 1]
    t0 = time.perf_counter()
    kept = s.remove_near_duplicate_points(line)
    out["dedup"] = {"points": M, "kept": int(kept.shape[0]), "host_call_ms": 1e3 * (time.perf_counter() - t0),
                    "algorithmic_bytes": 16 * M + 16 * int(kept.shape[0])}

    B, npts = 65536, 100
    counts = rng.integers(40, 260, B)
    y = rng.uniform(2, 140, counts.sum())
    x = 2e-4 * y * y + 0.02 * y + 5 + rng.normal(0, 0.15, y.shape[0])
    pts = np.stack([x, y], axis=1)
    offs = np.concatenate([[0], np.cumsum(counts)])
    trs = [pts[offs[b]:offs[b + 1]] for b in range(B)]
    t0 = time.perf_counter()
    sm = s.smooth_tracks_with_polyfit(trs, npts, 2)
    out["polyfit"] = {"tracks": B, "points": int(counts.sum()), "num_points": npts,
                      "host_call_ms": 1e3 * (time.perf_counter() - t0),
                      "algorithmic_bytes": 16 * int(counts.sum()) + 16 * B * npts}
    chk = rng.integers(0, B, 16)
    ref = np.stack([np.polyval(np.polyfit(trs[b][:, 1], trs[b][:, 0], 2), sm[b][:, 1]) for b in chk])
    out["polyfit"]["max_abs_diff_vs_numpy_m"] = float(np.abs(ref - sm[chk][:, :, 0]).max())

    cl = tracks.synthetic_centreline("nordschleife")
    Bx = 262144
    dev = torch.device("cuda:0")
    d_cl = torch.from_numpy(cl).to(dev)
    d_idx = torch.from_numpy(rng.integers(0, cl.shape[0], Bx).astype(np.int32)).to(dev)
    d_lat = torch.from_numpy(rng.uniform(-2, 2, Bx)).to(dev)
    d_psi = torch.from_numpy(rng.uniform(-0.1, 0.1, Bx)).to(dev)
    d_out = torch.empty((Bx, s.H, 3), dtype=torch.float64, device=dev)
    for _ in range(3):
        s.extract_paths_device(d_cl, d_idx, d_lat, d_psi, out=d_out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        s.extract_paths_device(d_cl, d_idx, d_lat, d_psi, out=d_out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = Bx * (s.H * 24 + 20)
    out["extract_paths"] = {"instances": Bx, "horizon": s.H, "kernel_ms": ms, "algorithmic_bytes": nbytes,
                            "GB_per_s": nbytes / ms / 1e6, "instances_per_s": Bx / ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
