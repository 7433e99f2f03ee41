"""Whole-track speed profile on every named track: device time of the cooperative
kernel (CUDA events inside the library), ADMM iterations, microseconds per iteration, end-to-end wall time of the
host call, and -- for --cpu tracks -- the oracle's C OSQP port on one host core beside it.

    python tools/map_profile_bench.py [--cpu monza,vallelunga] [--tracks monza,...] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ac_mpc_b200 import BatchedMPC, _capi, tracks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", default=",".join(tracks.TRACK_ORDER))
    ap.add_argument("--cpu", default="monza")
    ap.add_argument("--json", default=None)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    cpu = set(filter(None, a.cpu.split(",")))
    rows = []
    for tr in a.tracks.split(","):
        trk = tracks.map_track(tracks.synthetic_centreline(tr))
        c, mp = tracks.racing_config(tr)["speed_profile_constraints"], tracks.MAP_PROFILE[tr]
        mpc = BatchedMPC(_capi.default_config(v_min=c["v_min"], a_max=c["a_max"], ki_min=c["ki_min"]), device=0)
        best, wall = None, None
        for _ in range(a.repeat):
            t0 = time.perf_counter()
            way, x, info = mpc.track_speed_profile(trk, c["v_max"], mp["ay_max"], mp["a_min"])
            t1 = time.perf_counter()
            if best is None or info["kernel_ms"] < best["kernel_ms"]:
                best, wall = info, (t1 - t0) * 1e3
        row = dict(track=tr, n=len(x), ctas=best["ctas"], status=best["status_str"], iters=best["iters"],
                   rho_updates=best["rho_updates"], kernel_ms=round(best["kernel_ms"], 3),
                   us_per_iter=round(1e3 * best["kernel_ms"] / best["iters"], 3), host_call_ms=round(wall, 3))
        t0 = time.perf_counter()
        sm, wm = mpc.reference_speeds(way[6])
        row["reference_speeds_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        if tr in cpu:
            from oracle import port

            w0 = port.construct_waypoints(trk)
            t0 = time.perf_counter()
            xo, io = port.map_speed_profile(w0, c, mp["ay_max"], mp["a_min"])
            row["cpu_port_s"] = round(time.perf_counter() - t0, 3)
            row["cpu_iters"] = io.iter
            row["max_abs_diff"] = float(np.abs(xo - x).max())
        rows.append(row)
        print(json.dumps(row), flush=True)
        mpc.close()
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
