"""Which OSQP implementation the reference runs on here.  TEST INFRASTRUCTURE (see oracle/port.py).

The reference's numerical core is the third-party `osqp` wheel (requirements.txt:2, un-pinned; call sites
solvers/control.py:88-106, solvers/speed_profile.py:68-86,146).  This image has no such wheel and no network, so
everything was built against the C restatement (oracle/osqp_port.c, "PARITY UNPINNED").  The day a wheel is
importable, every consumer goes to it FIRST without code changes:

    select(prefer_real=True) -> (module, label)     label = "osqp <version>" or "oracle-port 0.6.x"

* `module` has the `osqp.OSQP` object API the reference calls (setup / update / solve).
* a real wheel is wrapped so that `setup` receives an EXPLICIT `adaptive_rho_interval` (default 50): OSQP <= 0.6 picks
  the interval from wall-clock timing when it is 0, which no test can reproduce; the port and the CUDA
  kernels take the same explicit value (`acmpc_config.adaptive_rho_interval`).
* `install(module)` puts the chosen module into sys.modules["osqp"] so the UNMODIFIED reference
  (`import osqp` at the top of its solver files) picks it up.
"""
from __future__ import annotations

import importlib
import os
import sys
from types import ModuleType, SimpleNamespace

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(_HERE, "shim")
PORT_LABEL = "oracle-port 0.6.x"


def real_osqp():
    """The installed third-party wheel, or None.  Never returns the shim."""
    saved = sys.modules.pop("osqp", None)
    path = [p for p in sys.path if os.path.abspath(p or ".") != SHIM_DIR]
    old_path, sys.path = sys.path, path
    try:
        mod = importlib.import_module("osqp")
        if getattr(mod, "__version__", "").endswith("-port"):
            return None
        return mod
    except Exception:
        return None
    finally:
        sys.path = old_path
        sys.modules.pop("osqp", None)
        if saved is not None:
            sys.modules["osqp"] = saved


def _shim():
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    saved = sys.modules.pop("osqp", None)
    try:
        return importlib.import_module("osqp")
    finally:
        if saved is not None and not getattr(saved, "__version__", "").endswith("-port"):
            pass


def wrap_real(mod, adaptive_rho_interval: int = 50, record=None, **forced):
    """A module-shaped wrapper around a real wheel: OSQP().setup gets `adaptive_rho_interval` (and any other
    `forced` setting, e.g. check_dualgap=False on 1.x) unless the caller passed it; solves are logged into
    `record` in the shim's format so the golden generators work on either."""
    import numpy as np

    major = int(str(getattr(mod, "__version__", "0")).split(".")[0] or 0)

    class OSQP:
        def __init__(self, *a, **k):
            self._s = mod.OSQP(*a, **k)
            self._n = self._m = 0

        def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
            settings.setdefault("adaptive_rho_interval", adaptive_rho_interval)
            for k, v in forced.items():
                settings.setdefault(k, v)
            self._n, self._m = A.shape[1], A.shape[0]
            if record is not None:
                record.append(("setup", dict(P=P.copy(), q=np.array(q), A=A.copy(), l=np.array(l), u=np.array(u))))
            return self._s.setup(P=P, q=q, A=A, l=l, u=u, **settings)

        def update(self, **kw):
            if record is not None:
                record.append(("update", {k: (None if v is None else np.array(v)) for k, v in kw.items()}))
            return self._s.update(**kw)

        def warm_start(self, **kw):
            return self._s.warm_start(**kw)

        def update_settings(self, **kw):
            return self._s.update_settings(**kw)

        def solve(self, *a, **k):
            res = self._s.solve(*a, **k)
            if record is not None:
                info = res.info
                record.append(("solve", dict(x=np.array(res.x, dtype=float), n=self._n, m=self._m,
                                             status_val=int(info.status_val), iter=int(info.iter),
                                             obj_val=float(info.obj_val),
                                             pri_res=float(getattr(info, "pri_res", getattr(info, "prim_res", float("nan")))),
                                             dua_res=float(getattr(info, "dua_res", getattr(info, "dual_res", float("nan")))),
                                             rho_updates=int(getattr(info, "rho_updates", 0)))))
            return res

    w = ModuleType("osqp")
    w.OSQP = OSQP
    w.__version__ = getattr(mod, "__version__", "unknown")
    w.__wrapped_real__ = mod
    w._RECORD = record if record is not None else []
    w.major = major
    for name in ("constant", "algebras_available", "algebra_available", "default_algebra"):
        if hasattr(mod, name):
            setattr(w, name, getattr(mod, name))
    return w


def select(prefer_real: bool = True, adaptive_rho_interval: int = 50, **forced):
    """(module, label): the real wheel (wrapped, see wrap_real) when importable and wanted, else the port's shim."""
    if prefer_real and os.environ.get("ACMPC_ORACLE_FORCE_PORT") != "1":
        mod = real_osqp()
        if mod is not None:
            rec = []
            return wrap_real(mod, adaptive_rho_interval, rec, **forced), f"osqp {mod.__version__}"
    return _shim(), PORT_LABEL


def install(module) -> None:
    sys.modules["osqp"] = module


def info() -> SimpleNamespace:
    mod = real_osqp()
    return SimpleNamespace(real=mod is not None, label=(f"osqp {mod.__version__}" if mod is not None else PORT_LABEL))
