"""The OSQP pin: tests/golden/osqp_pin.npz is written by tools/pin_osqp.py on a machine that
has a real `osqp` wheel -- the same cases as tests/golden/mpc_golden.npz, run by the unmodified reference Python on the
wheel with an explicit adaptive_rho_interval=50.

* file present and tagged "osqp <version>": the oracle port (CPU) and the CUDA path (-m gpu) are held to it --
  identical statuses / iteration counts, controls within the north-star bar of 1e-3 (m/s, rad).  PARITY PINNED.
* file absent: these tests SKIP with the reason "parity unpinned" -- nothing in this image can produce it
  (no wheel, no network).  The harness itself is exercised end to end on a disguised port (self-test)."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIN = os.path.join(ROOT, "tests", "golden", "osqp_pin.npz")
BAR = 1e-3      # north star: controls within 1e-3 absolute (steer rad, speed m/s) of the reference OSQP solve


def _load_pin(path=PIN):
    if not os.path.exists(path):
        pytest.skip("parity unpinned: tests/golden/osqp_pin.npz absent (no `osqp` wheel has met this tree; "
                    "run tools/pin_osqp.py where one is installed)")
    with np.load(path) as z:
        d = {k: z[k] for k in z.files}
    label = str(d["meta/solver"])
    if not label.startswith("osqp "):
        pytest.skip(f"pin file was written by {label!r}, not by a real wheel")
    # like-for-like: fixtures and kernels run OSQP 0.6.x termination; on a 1.x wheel use the check_dualgap-off run
    pre = "nodualgap__" if any(k.startswith("nodualgap__") for k in d) else ""
    return d, pre, label


def _groups(d, pre):
    import _golden

    for group, kw in _golden.groups():
        if f"{pre}{group}/paths" in d:
            yield group, kw, {k[len(pre) + len(group) + 1:]: v for k, v in d.items() if k.startswith(f"{pre}{group}/")}


def _check(got, g, label):
    assert np.array_equal(got["status"], g["status"]), label
    assert np.array_equal(got["status_speed"], g["status_speed"]), label
    assert np.array_equal(got["iters"], g["iters"]), f"ADMM iteration counts differ from {label}"
    ok = g["status"] == 1
    d = np.abs(got["controls"][ok] - g["controls"][ok])
    assert d[:, 0].max() < BAR and d[:, 1].max() < BAR, (label, d[:, 0].max(), d[:, 1].max())


def test_oracle_port_matches_the_real_wheel():
    from oracle import port

    d, pre, label = _load_pin()
    for group, kw, g in _groups(d, pre):
        B = g["paths"].shape[0]
        loc = g.get("localised", np.zeros(B, int))
        for flag in (0, 1):
            m = loc == flag
            if m.any():
                got = port.solve_batch(port.default_config(**kw), g["paths"][m], g.get("offsets", np.zeros(B))[m],
                                       g.get("vmax", np.full(B, kw["v_max"]))[m], bool(flag), nthreads=4)
                _check(got, {k: v[m] for k, v in g.items() if v.shape[:1] == (B,)}, f"{label} / {group}")


@pytest.mark.gpu
def test_cuda_path_matches_the_real_wheel():
    from ac_mpc_b200 import BatchedMPC, _capi

    d, pre, label = _load_pin()
    for group, kw, g in _groups(d, pre):
        B = g["paths"].shape[0]
        loc = g.get("localised", np.zeros(B, int))
        for flag in (0, 1):
            m = loc == flag
            if m.any():
                got = BatchedMPC(_capi.default_config(**kw), device=0).solve_host(
                    g["paths"][m], g.get("offsets", np.zeros(B))[m], g.get("vmax", np.full(B, kw["v_max"]))[m], bool(flag))
                _check(got, {k: v[m] for k, v in g.items() if v.shape[:1] == (B,)}, f"{label} / {group}")


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/acmpc"), reason="needs the reference sources (container only)")
def test_pin_harness_runs_end_to_end_on_a_disguised_port(tmp_path):
    """tools/pin_osqp.py --self-test: select -> run the unmodified reference -> write the pin file -> compare with the
    committed fixtures.  On the port the comparison must be exact (same solver, same cases); the file is tagged as NOT
    a pin and the consuming tests skip it."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("pin_osqp", os.path.join(ROOT, "tools", "pin_osqp.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    out, rep = str(tmp_path / "pin.npz"), str(tmp_path / "report.json")
    assert tool.main(["--self-test", "--out", out, "--report", rep]) == 0
    report = json.load(open(rep))
    s = report["as_installed"]["_summary"]
    assert s["all_status_equal"] and s["all_iters_equal"] and s["max_abs_dv"] == 0.0 and s["within_north_star_1e-3"]
    with np.load(out) as z:
        assert str(z["meta/solver"]).startswith("oracle-port")
    with pytest.raises(pytest.skip.Exception):
        _load_pin(out)


def test_real_wheel_is_preferred_and_gets_an_explicit_adaptive_rho_interval(tmp_path, monkeypatch):
    """oracle.osqp_select: a module named `osqp` on sys.path that is not the stand-in is taken first, and its setup()
    receives adaptive_rho_interval=50 (plus forced settings) unless the caller passed one."""
    import sys

    from oracle import osqp_select

    pkg = tmp_path / "osqp"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(
        "__version__ = '9.9.9'\n"
        "CALLS = []\n"
        "class OSQP:\n"
        "    def setup(self, **kw):\n"
        "        CALLS.append(kw)\n"
        "    def solve(self):\n"
        "        import types\n"
        "        return types.SimpleNamespace(x=[0.0], info=types.SimpleNamespace(status_val=1, iter=25, obj_val=0.0,\n"
        "                                     pri_res=0.0, dua_res=0.0, rho_updates=0))\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.delitem(sys.modules, "osqp", raising=False)
    assert osqp_select.info().label == "osqp 9.9.9"
    mod, label = osqp_select.select(check_dualgap=False)
    assert label == "osqp 9.9.9" and mod.__wrapped_real__.__version__ == "9.9.9"
    from scipy import sparse

    s = mod.OSQP()
    s.setup(P=sparse.eye(1, format="csc"), q=np.zeros(1), A=sparse.eye(1, format="csc"), l=np.zeros(1), u=np.ones(1),
            verbose=False, max_iter=4000)
    kw = mod.__wrapped_real__.CALLS[-1]
    assert kw["adaptive_rho_interval"] == 50 and kw["check_dualgap"] is False and kw["max_iter"] == 4000
    s.solve()
    assert [k for k, _ in mod._RECORD] == ["setup", "solve"]
    monkeypatch.setenv("ACMPC_ORACLE_FORCE_PORT", "1")
    assert osqp_select.select()[1] == osqp_select.PORT_LABEL
