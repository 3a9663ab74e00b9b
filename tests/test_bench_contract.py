"""bench.py's CPU-side contract: the `--impl reference` arm (no GPU needed) prints exactly one JSON line with the keys the driver reads,
its parts are the compiled C++/OpenMP ports, and under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_json_line():
    lines = run(["--impl", "reference", "--res", "T42", "--levels", "25", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1                                                     # ONE JSON line on stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "model_days_per_sec" and d["unit"] == "model-days/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "C++/OpenMP" in cb["sample"]
    assert set(cb["parts_s"]) == {"dynamics_cpp_s", "physics_cpp_s", "physics_cpp_kernels_ms", "rrtmg_cpp_scaled_per_call_s"}
    assert all(v > 0 for v in cb["parts_s"]["physics_cpp_kernels_ms"].values())
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["value"] - 720.0 / 86400.0 / (d["ms_per_step"] * 1e-3)) < 1e-9 * d["value"]      # MiMA dt at T42 = 720 s


def test_reference_arm_other_ranks_exit_without_work():
    assert run(["--impl", "reference", "--res", "T42", "--levels", "25", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_dry_reference_arm():
    d = json.loads(run(["--impl", "reference", "--workload", "hs", "--res", "T42", "--levels", "25", "--steps", "1", "--warmup", "0"])[0])
    assert set(d["cpu_baseline"]["parts_s"]) == {"dynamics_s"} and d["cpu_baseline"]["kind"] == "port"
