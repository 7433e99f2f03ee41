"""Generate tests/golden/publish_golden.npz: the caller side of the MPC step computed by
the UNMODIFIED reference classes -- control/commands.py (TemporalCommandSelector / TemporalCommandInterpolator) driven
through a stand-in controller object, and the expressions of ControlProcess._reference_path / _update_shared_memory
(control/controller.py:257-280; the class itself needs the simulator stack to construct) evaluated with the same numpy
calls.  Also stores the vectors of the reference's own test (tests/test_commands.py:15-58).

    python tests/golden/make_publish_golden.py        # needs /root/reference, so it cannot run on the GPU box
"""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = ["/root/reference/src"]

from acmpc.control.commands import TemporalCommandInterpolator, TemporalCommandSelector  # noqa: E402


def lookups(dtype, rng, B, n):
    """B random published plans + elapsed times covering: before the first stamp, exact hits, mid-points (ties),
    between stamps, beyond the last stamp."""
    cum = np.cumsum(rng.uniform(0.01, 0.08, (B, n)), axis=1).astype(dtype)
    cum[:, 0] = 0.0
    cmd = np.stack([rng.uniform(5, 80, (B, n)), rng.uniform(-0.3, 0.3, (B, n))], axis=2).astype(dtype)
    kinds = rng.integers(0, 5, B)
    k = rng.integers(0, n - 1, B)
    rows = np.arange(B)
    el = np.where(kinds == 0, -rng.uniform(0, 0.2, B),
         np.where(kinds == 1, cum[rows, k].astype(np.float64),
         np.where(kinds == 2, 0.5 * (cum[rows, k].astype(np.float64) + cum[rows, k + 1].astype(np.float64)),
         np.where(kinds == 3, rng.uniform(0, 1, B) * cum[:, -1], cum[:, -1] + rng.uniform(0, 0.5, B)))))
    sel, itp = np.zeros((B, 2), dtype), np.zeros((B, 2), dtype)
    for b in range(B):
        ctl = SimpleNamespace(control_cumtime=cum[b], control_inputs=cmd[b])
        sel[b] = TemporalCommandSelector(ctl)(float(el[b]))
        ctl_t = SimpleNamespace(control_cumtime=cum[b], control_inputs=cmd[b].T)   # the interpolator reads .T
        itp[b] = TemporalCommandInterpolator(ctl_t)(float(el[b]))
    return dict(cum_time=cum, commands=cmd, elapsed=el, selected=sel, interpolated=itp)


def main():
    out = {}
    rng = np.random.default_rng(7)
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        for k, v in lookups(dt, rng, 512, 49).items():
            out[f"lookup_{name}/{k}"] = v
    # tests/test_commands.py:15-24
    cum = np.round(np.linspace(0, 1, 10), 1)
    out["ref_test_index/cum_time"] = cum
    out["ref_test_index/elapsed"] = np.array([0, 0.22, 1.0, 0.95, 0.77])
    out["ref_test_index/expected_index"] = np.array([0, 2, 9, 8, 7])
    out["ref_test_index/expected_distance"] = np.array([0.0, -0.02, 0.0, -0.05, 0.03])
    # tests/test_commands.py:26-58
    out["ref_test_interp/cum_time"] = np.linspace(0, 1, 11)
    out["ref_test_interp/commands"] = np.array([[17.0, -0.03], [0.0, 0.0], [5.0, 0.15], [1.0, 0.0], [0.0, 0.0], [0.0, 0.0],
                                                [0.0, 0.0], [-5, -0.06], [12.0, 0.04], [1.0, 0.4], [-2.0, 0.02]])
    out["ref_test_interp/elapsed"] = np.array([-0.1, 0.22, 1.0, 0.95, 0.77, 1.1])
    out["ref_test_interp/expected"] = np.array([[17, -0.03], [4.2, 0.12], [-2.0, 0.02], [-0.5, 0.21], [6.9, 0.01], [-2.0, 0.02]])
    # the reference's own classes on its own vectors (float64), for the record
    got = []
    for t in out["ref_test_interp/elapsed"]:
        ctl = SimpleNamespace(control_cumtime=out["ref_test_interp/cum_time"], control_inputs=out["ref_test_interp/commands"].T)
        got.append(TemporalCommandInterpolator(ctl)(float(t)))
    assert np.allclose(np.array(got), out["ref_test_interp/expected"], atol=1e-7)
    # controller.py:257-267 on perceived centre lines (float32, 500 points) for H = 50, and :274-280 into float32
    B, P, H = 16, 500, 50
    c = np.cumsum(rng.normal(0, 0.2, (B, P, 2)), axis=1).astype(np.float32)
    paths = []
    for b in range(B):
        centreline = c[b]
        ds = int(len(centreline) / H)
        paths.append(np.stack([centreline[0::ds, 0], centreline[0::ds, 1], np.linspace(10.0, 6.0, H)]).T)
    out["reference_path/centrelines"] = c
    out["reference_path/paths"] = np.array(paths)
    n = H - 1
    ctrl, ct, pred = rng.uniform(-50, 80, (B, 2, n)), np.cumsum(rng.uniform(0.01, 0.1, (B, n)), axis=1), rng.normal(0, 30, (B, n, 2))
    ci, cc, pl = np.zeros((B, n, 2), np.float32), np.zeros((B, n), np.float32), np.zeros((B, n, 2), np.float32)
    for b in range(B):
        ci[b][:] = ctrl[b].T        # SharedPoints.points setter: np_array[:] = points (shared_memory.py:90-94)
        cc[b][:] = ct[b]
        pl[b][:] = pred[b]
    out.update({"publish/controls": ctrl, "publish/cum_time": ct, "publish/prediction": pred,
                "publish/control_inputs": ci, "publish/control_cumtime": cc, "publish/predicted_locations": pl})
    np.savez_compressed(os.path.join(HERE, "publish_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
