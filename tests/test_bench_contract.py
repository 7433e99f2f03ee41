"""CPU: the reference arm of bench.py prints ONE JSON line with the contract's keys, the same `config` object as the GPU
arm would, and a cpu_baseline that says which solver ran (the oracle's C port here: no osqp wheel, and the probe says so)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "64", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "batched MPC solves/sec" and line["unit"] == "solves/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert cb["solver"].startswith(("oracle-port", "osqp "))
    # the same `config` the GPU arm prints for this invocation (bench.bench_config is shared by both arms)
    sys.path.insert(0, ROOT)
    import bench

    class A:
        track, horizon, batch = "monza", 50, 64
    assert line["config"] == bench.bench_config(A, 1)
    assert "model" not in line["config"]
