"""CPU-only: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/acmpc_b200.h declares; without a CUDA device the product refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ac_mpc_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_builds_and_exports_every_declared_symbol():
    _capi.build()
    lib = C.CDLL(_capi.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "acmpc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(acmpc_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert set(declared) == set(_capi.EXPORTED)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.acmpc_abi_version() == 2


def test_config_struct_layout_matches_the_header_defaults():
    cfg = _capi.default_config()
    assert (cfg.horizon, cfg.max_iter, cfg.scaling, cfg.check_termination) == (50, 4000, 10, 25)
    assert (cfg.rho, cfg.sigma, cfg.alpha, cfg.eps_abs, cfg.eps_rel) == (0.1, 1e-6, 1.6, 1e-3, 1e-3)
    assert (cfg.adaptive_rho, cfg.adaptive_rho_interval, cfg.adaptive_rho_tolerance) == (1, 50, 5.0)
    assert list(cfg.step_cost) == [4e-3, 5e-2, 0.0] and list(cfg.final_cost) == [1.0, 0.0, 0.1]
    assert (cfg.wheelbase, cfg.width, cfg.delta_max) == (2.65, 1.99, 0.30)
    # the oracle mirrors the same struct: same size, same defaults
    from oracle import port

    assert C.sizeof(port.Config) == C.sizeof(_capi.Config)
    assert bytes(port.default_config()) == bytes(cfg)


def test_create_rejects_bad_configs():
    lib = _capi.load()
    h = C.c_void_p()
    for kw in (dict(horizon=3), dict(horizon=129), dict(max_iter=0), dict(alpha=2.5), dict(wheelbase=0.0)):
        assert lib.acmpc_create(C.byref(_capi.default_config(**kw)), 0, C.byref(h)) == 1
        assert not h.value


@pytest.mark.skipif(_has_cuda(), reason="checks the behaviour WITHOUT a device")
def test_no_device_means_no_solver():
    from ac_mpc_b200 import BatchedMPC, tracks
    from ac_mpc_b200.control import build_mpc

    lib = _capi.load()
    h = C.c_void_p()
    assert lib.acmpc_create(C.byref(_capi.default_config()), 0, C.byref(h)) == 3   # ACMPC_ERR_NO_DEVICE
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedMPC(_capi.default_config()).solve_host(np.zeros((1, 50, 3)))
    veh = type("V", (), {"vehicle_data": type("D", (), {"wheelbase": 2.65, "width": 1.99})(),
                         "max_steering_angle": lambda self: 0.3})()
    mpc = build_mpc(tracks.racing_config("monza"), veh)            # construction is lazy (fork safety)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mpc.get_control(np.zeros((50, 3)))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="is missing"):
        _capi.load()


def test_host_wrapper_validates_shapes():
    from ac_mpc_b200 import BatchedMPC

    mpc = BatchedMPC(_capi.default_config())
    with pytest.raises(ValueError):
        mpc.solve_host(np.zeros((2, 49, 3)))
    with pytest.raises(ValueError):
        mpc.solve_host(np.zeros((2, 50, 3)), offsets=np.zeros(3))
