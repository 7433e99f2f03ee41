"""CPU-only: the arithmetic of the CUDA kernel body (ac_mpc_b200/csrc/mpc_warp.cuh compiled by g++ with
32 lock-step lanes, tests/_emul) against the golden vectors of the reference Python.  This does not exercise the
32-lane execution -- tests/test_gpu_parity.py does, on the B200."""
import numpy as np
import pytest

import _cases
import _emul
from oracle import port

CASES = list(_cases.golden_batches())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_one_lane_body_matches_golden(case):
    _, kw, paths, offs, vmax, loc, want = case
    cfg = port.default_config(**kw)
    got = _emul.solve_batch(cfg, paths, offs, vmax, loc)
    _cases.assert_matches_golden(got, want, cfg.horizon, atol=1e-8)


def test_one_lane_body_matches_oracle_on_a_perturbed_batch():
    from ac_mpc_b200 import tracks

    paths, vmax = tracks.perturbed_batch("monza", 96, seed=11)
    cfg = port.default_config()
    a = _emul.solve_batch(cfg, paths, None, vmax, False)
    b = port.solve_batch(cfg, paths, None, vmax, False, nthreads=4)
    assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["rho_updates"], b["rho_updates"])
    assert np.abs(a["controls"] - b["controls"]).max() < 1e-8
