"""`osqp`-shaped module over the C restatement (oracle/osqp_port.c).

TEST INFRASTRUCTURE.  Lets the UNMODIFIED reference Python
(/root/reference/src/acmpc/control/solvers/control.py:88-106, speed_profile.py:68-86,146) run in
this container, where the real `osqp` wheel is absent.  Only the calls the reference makes are
provided: OSQP().setup(P,q,A,l,u,verbose,max_iter,...), .update(Ax,q,l,u), .solve() -> .x/.y/.info.
"""
from types import SimpleNamespace

import numpy as np

from oracle import port as _port

__version__ = "0.6-port"
_RECORD = []  # (kind, kwargs) log, used by the assembly tests


class OSQP:
    def __init__(self):
        self._solver = None

    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        settings.pop("verbose", None)
        _RECORD.append(("setup", dict(P=P.copy(), q=np.array(q), A=A.copy(), l=np.array(l), u=np.array(u))))
        self._solver = _port.PortOSQP(P, q, A, l, u, **settings)

    def update(self, q=None, l=None, u=None, Ax=None, Px=None, Ax_idx=None, Px_idx=None):
        if Px is not None or Ax_idx is not None or Px_idx is not None:
            raise NotImplementedError("shim supports update(q, l, u, Ax) only")
        _RECORD.append(("update", dict(q=q, l=l, u=u, Ax=None if Ax is None else np.array(Ax))))
        self._solver.update(q=q, l=l, u=u, Ax=Ax)

    def warm_start(self, x=None, y=None):
        self._solver.warm_start(x, y)

    def solve(self):
        x, y, info = self._solver.solve()
        status = _port.STATUS_STRINGS[info.status]
        _RECORD.append(("solve", dict(x=x.copy(), n=self._solver.n, m=self._solver.m, status_val=info.status,
                                      iter=info.iter, obj_val=info.obj_val, pri_res=info.pri_res,
                                      dua_res=info.dua_res, rho_updates=info.rho_updates)))
        if info.status in (-3, 3, -4, 4, -7):
            x = np.full_like(x, np.nan)
            y = np.full_like(y, np.nan)
        return SimpleNamespace(
            x=x, y=y,
            info=SimpleNamespace(status=status, status_val=info.status, iter=info.iter,
                                 obj_val=info.obj_val, pri_res=info.pri_res, dua_res=info.dua_res,
                                 rho_updates=info.rho_updates, rho_estimate=info.rho_estimate),
        )
