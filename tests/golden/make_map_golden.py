"""Generate tests/golden/map_golden.npz: the whole-track speed profile computed by
the UNMODIFIED reference Python -- SpatialMPC.construct_waypoints + compute_map_speed_profile
(/root/reference/src/acmpc/control/spatial_mpc.py:60-87,125-154), then agent.py:300 / :137-143 restated with the
same scipy / numpy calls -- on top of the `osqp` stand-in of oracle/shim (PARITY UNPINNED, see oracle/osqp_port.h).

    python tests/golden/make_map_golden.py        # needs /root/reference, so it cannot run on the GPU box
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shim"), "/root/reference/src"]

import osqp  # noqa: E402  (the shim)
from ace.steering import SteeringGeometry  # noqa: E402
from acmpc.control.controller import build_mpc  # noqa: E402
from scipy.signal import savgol_filter  # noqa: E402

from ac_mpc_b200 import tracks  # noqa: E402

# (group, track block, spacing of the synthetic centre line in metres, points kept)
CASES = [("vallelunga_1m", "vallelunga", 1.0, None), ("monza_2m", "monza", 2.0, None),
         ("silverstone_arc", "silverstone", 0.5, 700), ("spa_arc", "spa", 0.5, 300)]


def main():
    out = {}
    veh = SteeringGeometry()
    for group, tr, ds, keep in CASES:
        cl = tracks.synthetic_centreline(tr, ds=ds)
        if keep:
            cl = cl[:keep]
        track = tracks.map_track(cl)
        mpc = build_mpc(tracks.racing_config(tr, 50), veh)
        mp = tracks.MAP_PROFILE[tr]
        del osqp._RECORD[:]
        # controller.py:49-57
        waypoints = mpc.construct_waypoints(track)
        before = np.array(waypoints._reference_path)
        profile = mpc.compute_map_speed_profile(waypoints, ay_max=mp["ay_max"], a_min=mp["a_min"])
        rec = [r for k, r in osqp._RECORD if k == "solve"][0]
        v = np.array(profile.velocities)
        out[f"{group}/track"] = track
        out[f"{group}/waypoints_in"] = before
        out[f"{group}/waypoints"] = np.array(profile._reference_path)
        out[f"{group}/dec_x"] = rec["x"]
        out[f"{group}/status"] = np.int32(rec["status_val"])
        out[f"{group}/iters"] = np.int32(rec["iter"])
        out[f"{group}/rho_updates"] = np.int32(rec["rho_updates"])
        out[f"{group}/obj_val"] = np.float64(rec["obj_val"])
        out[f"{group}/pri_res"] = np.float64(rec["pri_res"])
        out[f"{group}/dua_res"] = np.float64(rec["dua_res"])
        out[f"{group}/constraints"] = np.array([mpc.speed_profile_constraints["v_max"], mp["ay_max"], mp["a_min"]])
        # agent.py:300 and agent.py:137-143 (REFERENCE_SPEED_WINDOW_BEHIND / AHEAD = 25 / 75) for every map index
        sm = savgol_filter(v, 21, 3)
        idx = np.arange(len(v))[:, None] + np.arange(-25, 75)[None, :]
        out[f"{group}/reference_speeds"] = sm
        out[f"{group}/window_mean"] = np.array([np.mean(sm.take(i, mode="wrap")) for i in idx])
        print(group, "n =", len(v), "status", rec["status_val"], "iters", rec["iter"], "rho updates", rec["rho_updates"],
              "v in [%.2f, %.2f]" % (v.min(), v.max()))
    np.savez_compressed(os.path.join(HERE, "map_golden.npz"), **out)


if __name__ == "__main__":
    main()
