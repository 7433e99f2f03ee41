#!/usr/bin/env python
"""Pin the OSQP restatement (oracle/osqp_port.c) -- and through it the CUDA kernels -- against a REAL `osqp` wheel.

Run ONCE by whoever has the wheel (this image has none and no network, so the repo ships "parity unpinned"):

    pip install osqp            # any version: 0.6.x or 1.x
    python tools/pin_osqp.py [--reference /path/to/ac-mpc/src] [--out tests/golden/osqp_pin.npz]

What it does
  1. imports the wheel (never the stand-in) and wraps it so that `setup` receives an explicit
     `adaptive_rho_interval=50` (oracle/osqp_select.py: the 0.6.x default is wall-clock based);
  2. runs the UNMODIFIED reference Python (build_mpc -> SpatialMPC.get_control / compute_speed_profile, the case
     groups of tests/golden/make_golden.py) on it and writes every input and output to --out, tagged
     meta/solver = "osqp <version>";
  3. compares with the committed port-generated fixtures (tests/golden/mpc_golden.npz: identical inputs) and prints a
     parity report: status / iteration / rho-update agreement and max |dv| (m/s), |ddelta| (rad) per group.  On a 1.x
     wheel the cases are run twice, with check_dualgap on (the 1.x default) and off, so the report says which
     `acmpc_config.check_dualgap` value the maintainer's wheel corresponds to.
Afterwards `pytest tests/test_osqp_pin.py` (CPU: oracle port; -m gpu: CUDA path) consumes the file: controls within the
north-star bar of 1e-3 and identical iteration counts flip "parity unpinned" to pinned with no code change.

--self-test runs the whole flow on the port disguised as a wheel (used by tests/test_osqp_pin.py; the file it writes is
tagged "oracle-port" and does NOT count as a pin).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def compare(pin: dict, gold: dict) -> dict:
    """Per-group agreement of two runs of the same cases (`group/name` -> array)."""
    groups = sorted({k.split("/")[0] for k in pin if "/status" in k and not k.startswith(("meta", "qp_", "cut_a"))})
    rep = {}
    for g in groups:
        st_p, st_g = pin[f"{g}/status"], gold[f"{g}/status"]
        ok = (st_p == 1) & (st_g == 1)
        r = {"cases": int(st_p.shape[0]), "status_equal": bool(np.array_equal(st_p, st_g)),
             "status_speed_equal": bool(np.array_equal(pin[f"{g}/status_speed"], gold[f"{g}/status_speed"])),
             "iters_equal": bool(np.array_equal(pin[f"{g}/iters"], gold[f"{g}/iters"])),
             "iters_mismatches": int((pin[f"{g}/iters"] != gold[f"{g}/iters"]).any(axis=1).sum())}
        if ok.any():
            dc = np.abs(pin[f"{g}/controls"][ok] - gold[f"{g}/controls"][ok])
            r["max_abs_dv"] = float(dc[:, 0].max())
            r["max_abs_ddelta"] = float(dc[:, 1].max())
            r["max_abs_dv_ref"] = float(np.abs(pin[f"{g}/v_ref"][ok] - gold[f"{g}/v_ref"][ok]).max())
            r["max_abs_dprediction"] = float(np.abs(pin[f"{g}/prediction"][ok] - gold[f"{g}/prediction"][ok]).max())
        rep[g] = r
    if "cut_a/iters" in pin and "cut_a/iters" in gold:
        rep["cut_a"] = {"iters_equal": bool(np.array_equal(pin["cut_a/iters"], gold["cut_a/iters"])),
                        "max_abs_dx": float(np.abs(pin["cut_a/x"] - gold["cut_a/x"]).max())}
    vals = [r for r in rep.values() if "max_abs_dv" in r]
    rep["_summary"] = {
        "all_status_equal": all(r.get("status_equal", True) for r in rep.values()),
        "all_iters_equal": all(r.get("iters_equal", True) for r in rep.values()),
        "max_abs_dv": max((r["max_abs_dv"] for r in vals), default=None),
        "max_abs_ddelta": max((r["max_abs_ddelta"] for r in vals), default=None),
        "within_north_star_1e-3": bool(vals) and all(r["max_abs_dv"] < 1e-3 and r["max_abs_ddelta"] < 1e-3 for r in vals),
    }
    return rep


def run_cases(env: dict) -> dict:
    """tests/golden/make_golden.generate() in a fresh import under `env` (it selects the solver at import time)."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: v for k, v in env.items() if v is not None})
    for k, v in env.items():
        if v is None:
            os.environ.pop(k, None)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "osqp" or k.startswith(("acmpc", "ace", "aci"))}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        sys.modules.pop("make_golden", None)
        mg = importlib.import_module("make_golden")
        return mg.generate(), mg.SOLVER_LABEL
    finally:
        sys.path.remove(os.path.join(ROOT, "tests", "golden"))
        sys.modules.pop("make_golden", None)
        for k in [k for k in sys.modules if k == "osqp" or k.startswith(("acmpc", "ace", "aci"))]:
            sys.modules.pop(k)
        sys.modules.update(saved)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=os.environ.get("ACMPC_REFERENCE_SRC", "/root/reference/src"),
                    help="the `src` directory of an ac-mpc checkout (the unmodified reference Python)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "osqp_pin.npz"))
    ap.add_argument("--report", default=os.path.join(ROOT, "profiles", "osqp_pin_report.json"))
    ap.add_argument("--self-test", action="store_true", help="run the flow on the port disguised as a wheel")
    args = ap.parse_args(argv)

    from oracle import osqp_select

    if not os.path.isdir(os.path.join(args.reference, "acmpc", "control")):
        print(f"pin_osqp: no ac-mpc sources under {args.reference} (pass --reference)", file=sys.stderr)
        return 2
    wheel = None if args.self_test else osqp_select.real_osqp()
    if wheel is None and not args.self_test:
        print("pin_osqp: no `osqp` wheel importable in this interpreter -- parity stays UNPINNED.\n"
              "          pip install osqp (0.6.x or 1.x) where a network exists and run this script again.", file=sys.stderr)
        return 3
    env = {"ACMPC_REFERENCE_SRC": args.reference, "ACMPC_GOLDEN_SOLVER": None if args.self_test else "real",
           "ACMPC_PIN_NO_DUALGAP": None}
    gold_path = os.path.join(ROOT, "tests", "golden", "mpc_golden.npz")
    with np.load(gold_path) as z:
        gold = {k: z[k] for k in z.files}
    pin, label = run_cases(env)
    report = {"solver": label, "fixtures": "tests/golden/mpc_golden.npz (" + str(gold.get("meta/solver", "?")) + ")",
              "adaptive_rho_interval": 50, "as_installed": compare(pin, gold)}
    major = int(str(getattr(wheel, "__version__", "0")).split(".")[0] or 0) if wheel is not None else 0
    if major >= 1:
        # 1.x: the default run above had check_dualgap on; a second run with it off tells the two semantics apart
        off, _ = run_cases(dict(env, ACMPC_PIN_NO_DUALGAP="1"))
        report["check_dualgap_off"] = compare(off, gold)
        report["hint"] = ("fixtures are generated with acmpc_config.check_dualgap = 0 (OSQP 0.6.x termination): "
                          "`check_dualgap_off` is the like-for-like comparison; if only `as_installed` disagrees, run the "
                          "product with check_dualgap = 1 to match this wheel's defaults")
        for k, v in off.items():
            pin["nodualgap__" + k] = v
    pin["meta/solver"] = np.array(label)
    pin["meta/report"] = np.array(json.dumps(report))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    np.savez_compressed(args.out, **pin)
    os.makedirs(os.path.dirname(os.path.abspath(args.report)), exist_ok=True)
    with open(args.report, "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report["as_installed"]["_summary"], indent=1))
    print(f"wrote {args.out} ({len(pin)} arrays, solver: {label}) and {args.report}")
    pinned = label.startswith("osqp ")
    print("parity:", "PINNED against " + label if pinned else "still UNPINNED (self-test on the port)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
