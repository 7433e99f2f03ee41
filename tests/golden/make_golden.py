"""Generate tests/golden/mpc_golden.npz by running the UNMODIFIED reference Python
(/root/reference/src/acmpc/control, entry point build_mpc -> SpatialMPC.get_control) in this
container, with three stand-ins for packages that are absent here (oracle/shim): `osqp` (-> the C
restatement oracle/osqp_port.c; PARITY UNPINNED, see oracle/osqp_port.h), `ace.steering` (synthetic
vehicle constants) and `aci.utils.system_monitor` (no-op decorator).

    python tests/golden/make_golden.py        # needs /root/reference, so it cannot run on the GPU box

What is pinned by these vectors: the reference's own waypoint / speed-bound / linearisation /
QP-assembly / unpack / rollout code (its numpy + scipy.sparse arithmetic), composed with the oracle's
OSQP restatement.  The C port of the whole step (oracle/acmpc_port.c) and the CUDA path are both
tested against this file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE_SRC = os.environ.get("ACMPC_REFERENCE_SRC", "/root/reference/src")
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shim"), REFERENCE_SRC]

from oracle import osqp_select  # noqa: E402

# The committed fixtures are made on the oracle's OSQP restatement (reproducible bit for bit in this container).
# tools/pin_osqp.py imports this module with ACMPC_GOLDEN_SOLVER=real to run the SAME cases on a real `osqp` wheel
# (explicit adaptive_rho_interval=50, see oracle/osqp_select.py) and writes tests/golden/osqp_pin.npz next to them.
osqp, SOLVER_LABEL = osqp_select.select(prefer_real=os.environ.get("ACMPC_GOLDEN_SOLVER") == "real",
                                        **({"check_dualgap": False} if os.environ.get("ACMPC_PIN_NO_DUALGAP") else {}))
osqp_select.install(osqp)
from ace.steering import SteeringGeometry  # noqa: E402
from acmpc.control.controller import build_mpc  # noqa: E402
from acmpc.control.utils import (  # noqa: E402
    get_chicane_track, get_curved_track, get_hairpin_track, get_straight_track)

from ac_mpc_b200 import tracks  # noqa: E402

STATUS = {v: k for k, v in __import__("oracle.port", fromlist=["x"]).STATUS_STRINGS.items()}

# /root/reference/src/acmpc/tests/test_spatial_mpc.py:16-31
FIXTURE_CONFIG = {
    "horizon": 100,
    "unlocalised_max_speed": 28,
    "speed_profile_constraints": {"v_min": 12.0, "v_max": 84.0, "a_min": -1.0, "a_max": 1.0,
                                  "ay_max": 5.5, "ki_min": 0.005, "end_velocity": 14.0},
    "step_cost": [2.0e-3, 5.0e-2, 0.0],
    "r_term": [1.0e-2, 10.0],
    "final_cost": [1.0, 0.0, 0.1],
}


def fixture_paths(N=100, road_width=100.0):
    """The 4 x 7 paths of test_spatial_mpc.py:36-83."""
    experiments, angle = 7, 0.1
    coeff = np.linspace(-0.02, 0.02, experiments)
    radii = np.linspace(10, 100, experiments)
    dist = np.linspace(40, 100, experiments)
    los = np.linspace(40, 200, experiments)
    out = []
    for kind in ["hairpin", "chicane", "curve", "straight"]:
        for i in range(experiments):
            if kind == "hairpin":
                x, y = get_hairpin_track(radii[i], N, -np.pi / 6)
            elif kind == "chicane":
                x, y = get_chicane_track(dist[i], 40, N, angle)
            elif kind == "curve":
                x, y = get_curved_track(coeff[i], N, angle)
            else:
                x, y = get_straight_track(los[i], N, angle)
            out.append(np.stack([x, y, np.ones(N) * road_width]).T)
    return np.array(out)


def run_case(mpc, path, is_localised, offset, v_max):
    """One get_control; returns everything observable, including the two OSQP results."""
    H = mpc.MPC_horizon
    n = H - 1
    if v_max is not None:
        mpc.speed_profile_constraints["v_max"] = float(v_max)   # controller.py:241-243
    del osqp._RECORD[:]
    before = mpc.infeasibility_counter
    mpc.get_control(path, is_localised, offset)
    solves = [r for k, r in osqp._RECORD if k == "solve"]
    sp, ct = solves[0], solves[1]
    assert sp["n"] == n and ct["n"] == 5 * H - 2
    solved = ct["status_val"] == 1
    rec = dict(
        status=ct["status_val"], status_speed=sp["status_val"], iters=[sp["iter"], ct["iter"]],
        rho_updates=[sp["rho_updates"], ct["rho_updates"]], cost=ct["obj_val"], pri_res=ct["pri_res"],
        dua_res=ct["dua_res"], dec_x=ct["x"], speed_x=sp["x"],
        controls=np.array(mpc.projected_control) if solved else np.full((2, n), np.nan),
        prediction=np.array(mpc.current_prediction) if solved else np.full((n, 2), np.nan),
        cum_time=np.array(mpc.cum_time) if solved else np.full(n, np.nan),
        v_ref=np.array(mpc.reference_path.velocities) if solved else np.full(n, np.nan),
        infeasibility_delta=mpc.infeasibility_counter - before,
        # spatial_mpc.py:208-211 (accelerations: sic, diff of the e_y column)
        times=np.array(mpc.times) if solved else np.full(n - 1, np.nan),
        accelerations=np.array(mpc.accelerations) if solved else np.full(n - 1, np.nan),
        steer_rates=np.array(mpc.steer_rates) if solved else np.full(n - 1, np.nan),
    )
    if solved:
        rec["waypoints"] = np.array(mpc.reference_path._reference_path)
    else:
        rec["waypoints"] = np.full((7, n), np.nan)
    return rec


def stack(recs):
    return {k: np.array([r[k] for r in recs]) for k in recs[0]}


def generate():
    """Every case group -> dict of arrays ("group/name").  Runs on whichever `osqp` module was selected above."""
    out = {"meta/solver": np.array(SOLVER_LABEL)}
    veh = SteeringGeometry()
    # (a) the reference's own fixture grid, cold start (fresh objects per case)
    paths = fixture_paths()
    recs = [run_case(build_mpc(dict(FIXTURE_CONFIG, speed_profile_constraints=dict(FIXTURE_CONFIG["speed_profile_constraints"])), veh),
                     p, False, 0.0, None) for p in paths]
    out.update({f"fixture_cold/{k}": v for k, v in stack(recs).items()})
    out["fixture_cold/paths"] = paths
    # (b) the same grid on ONE object, as the reference test runs it (warm starts, carried rho)
    mpc = build_mpc(dict(FIXTURE_CONFIG, speed_profile_constraints=dict(FIXTURE_CONFIG["speed_profile_constraints"])), veh)
    recs = [run_case(mpc, p, False, 0.0, None) for p in paths]
    out.update({f"fixture_warm/{k}": v for k, v in stack(recs).items()})
    # (c) racing blocks of all 7 tracks at H=50 on synthetic centrelines, perturbed, cold
    for ti, tr in enumerate(tracks.TRACK_ORDER):
        cl = tracks.synthetic_centreline(tr)
        rng = np.random.default_rng(100 + ti)
        B = 6
        idx = rng.integers(0, cl.shape[0], B)
        lat, psi = rng.uniform(-2, 2, B), rng.uniform(-0.1, 0.1, B)
        vmax = rng.uniform(20, 84, B)
        offs = np.array([0.0, 0.0, 0.0, 0.3, -0.4, 0.0])
        loc = np.array([0, 0, 0, 0, 1, 1])
        p = tracks.make_instances(cl, idx, 50, lat, psi)
        recs = [run_case(build_mpc(tracks.racing_config(tr, 50), veh), p[b], bool(loc[b]), offs[b], vmax[b])
                for b in range(B)]
        out.update({f"racing_{tr}/{k}": v for k, v in stack(recs).items()})
        out[f"racing_{tr}/paths"] = p
        out[f"racing_{tr}/vmax"] = vmax
        out[f"racing_{tr}/offsets"] = offs
        out[f"racing_{tr}/localised"] = loc
    # (d) horizon sweep on the Spa block (BASELINE.json configs[3]), cold
    cl = tracks.synthetic_centreline("spa")
    for H in (20, 40, 80):
        rng = np.random.default_rng(200 + H)
        B = 4
        idx = rng.integers(0, cl.shape[0], B)
        lat, psi = rng.uniform(-2, 2, B), rng.uniform(-0.1, 0.1, B)
        vmax = rng.uniform(20, 84, B)
        p = tracks.make_instances(cl, idx, H, lat, psi)
        recs = [run_case(build_mpc(tracks.racing_config("spa", H), veh), p[b], False, 0.0, vmax[b]) for b in range(B)]
        out.update({f"spa_h{H}/{k}": v for k, v in stack(recs).items()})
        out[f"spa_h{H}/paths"] = p
        out[f"spa_h{H}/vmax"] = vmax
    # (e) the full QP data the reference hands to osqp.setup for one Monza instance
    mpc = build_mpc(tracks.racing_config("monza", 50), veh)
    p = out["racing_monza/paths"][0]
    mpc.speed_profile_constraints["v_max"] = float(out["racing_monza/vmax"][0])
    del osqp._RECORD[:]
    mpc.get_control(p, False, 0.0)
    setups = [r for k, r in osqp._RECORD if k == "setup"]
    for name, s in zip(("speed", "control"), setups):
        A = s["A"].tocoo()
        out[f"qp_{name}/A_row"], out[f"qp_{name}/A_col"], out[f"qp_{name}/A_val"] = A.row, A.col, A.data
        out[f"qp_{name}/shape"] = np.array(A.shape)
        out[f"qp_{name}/P_diag"] = s["P"].diagonal()
        assert abs(s["P"] - __import__("scipy.sparse", fromlist=["x"]).diags(s["P"].diagonal())).sum() == 0
        for k in ("q", "l", "u"):
            out[f"qp_{name}/{k}"] = s[k]
    # (f) cut A of the boundary called stand-alone: compute_speed_profile on ONE object through a
    # sequence that mixes it with get_control (the speed solvers are shared, spatial_mpc.py:101-105), and the
    # SpatialBicycleModel methods t2s / s2t / linearise + update_prediction on the resulting paths
    mpc = build_mpc(tracks.racing_config("monza", 50), veh)
    P = out["racing_monza/paths"]
    seq = [("speed", 0, False, 14.0), ("speed", 1, False, None), ("step", 2, False, None), ("speed", 3, True, 9.0),
           ("speed", 4, False, 11.5), ("step", 5, True, None), ("speed", 0, True, None), ("speed", 2, False, 14.0)]
    rng = np.random.default_rng(77)
    rec = {k: [] for k in ("kind", "path_index", "localised", "end_vel", "vmax", "way_in", "way_out", "x", "status", "iters",
                           "rho_updates", "t2s_state", "t2s_out", "s2t_states", "s2t_out", "pred_out", "lin_f", "lin_A",
                           "lin_B")}
    for kind, pi, loc, ev in seq:
        vm = float(rng.uniform(30, 84))
        mpc.speed_profile_constraints["v_max"] = vm
        del osqp._RECORD[:]
        rec["kind"].append(0 if kind == "speed" else 1), rec["path_index"].append(pi), rec["localised"].append(int(loc))
        rec["end_vel"].append(np.nan if ev is None else ev), rec["vmax"].append(vm)
        if kind == "step":
            mpc.get_control(P[pi], loc, 0.0)
            sp = [r for k, r in osqp._RECORD if k == "solve"][0]
            z = np.zeros((7, 49))
            for k in ("way_in", "way_out"):
                rec[k].append(z)
            rec["x"].append(sp["x"]), rec["status"].append(sp["status_val"]), rec["iters"].append(sp["iter"])
            rec["rho_updates"].append(sp["rho_updates"])
            for k in ("t2s_state", "t2s_out"):
                rec[k].append(np.zeros(3))
            rec["s2t_states"].append(np.zeros((49, 3))), rec["s2t_out"].append(np.zeros((3, 49)))
            rec["pred_out"].append(np.zeros((49, 2))), rec["lin_f"].append(np.zeros((49, 3)))
            rec["lin_A"].append(np.zeros((49, 3, 3))), rec["lin_B"].append(np.zeros((49, 3, 2)))
            continue
        rp = mpc.construct_waypoints(P[pi])
        rp.velocities = np.full(49, 5.0 + pi)          # must survive a failed solve untouched
        rec["way_in"].append(np.array(rp._reference_path))
        rp = mpc.compute_speed_profile(rp, loc, end_vel=ev)
        sp = [r for k, r in osqp._RECORD if k == "solve"][0]
        rec["way_out"].append(np.array(rp._reference_path))
        rec["x"].append(sp["x"]), rec["status"].append(sp["status_val"]), rec["iters"].append(sp["iter"])
        rec["rho_updates"].append(sp["rho_updates"])
        state = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), np.pi / 2 + rng.uniform(-3.5, 3.5)])
        rec["t2s_state"].append(state), rec["t2s_out"].append(mpc.model.t2s(rp.get_state(0), state))
        xs = np.column_stack([rng.uniform(-2, 2, 49), rng.uniform(-0.3, 0.3, 49), np.cumsum(rng.uniform(0.01, 0.1, 49))])
        rec["s2t_states"].append(xs), rec["s2t_out"].append(mpc.model.s2t(rp, xs))
        rec["pred_out"].append(mpc.update_prediction(xs, rp))
        f, A, Bm = mpc.model.linearise(rp)
        rec["lin_f"].append(f), rec["lin_A"].append(A), rec["lin_B"].append(Bm)
    out.update({f"cut_a/{k}": np.array(v) for k, v in rec.items()})
    return out


def main():
    out = generate()
    if SOLVER_LABEL != osqp_select.PORT_LABEL:
        sys.exit("make_golden.py writes the port-generated fixtures; use tools/pin_osqp.py to run the cases on " + SOLVER_LABEL)
    np.savez_compressed(os.path.join(HERE, "mpc_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "mpc_golden.npz"), len(out), "arrays, solver:", SOLVER_LABEL)
    for g in sorted({k.split("/")[0] for k in out}):
        if f"{g}/status" in out:
            print(g, "status", out[f"{g}/status"].tolist(), "iters", out[f"{g}/iters"].tolist())


if __name__ == "__main__":
    main()
