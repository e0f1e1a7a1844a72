"""bench.py's reference arm runs without a GPU: check the JSON line the driver parses (keys, units, the
`impl`, `cpu_baseline` and `e2e` objects of the tier's contract) on a small mesh."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--factor", "1", "--scenarios", "3", "--solve-scenarios", "0"] + extra, cwd=ROOT, env=e, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_prints_the_contract_line():
    lines = _run([])
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["metric"] == "fd_jacobian_plus_residual_evals_per_s"
    assert b["unit"] == "evals/s" and b["higher_is_better"] is True and b["scaling"] == "weak"
    # the arm honours the driver's --steps / --warmup / --scenarios
    assert b["n_gpus"] == 1 and b["steps"] == 2 and b["warmup"] == 1 and b["dtype"] == "f64" and b["data"] == "synthetic"
    assert b["config"]["scenarios_per_gpu"] == 3
    assert b["value"] > 0 and b["ms_per_step"] > 0 and b["vs_baseline"] is None and b["gpu_launches"] == 0
    assert "workload" in b["config"] and "model" not in b["config"]
    assert b["solves"] is None and b["solves_per_hour"] is None  # the solve leg (minutes of CPU work) is switched off here
    cb = b["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == b["value"] and cb["sample"]
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value = evaluations of the step / time of the step
    evals = b["config"]["scenarios_per_gpu"] * b["config"]["evals_per_scenario_step"]
    assert abs(b["pairs_per_s"] - b["config"]["scenarios_per_gpu"] / (b["ms_per_step"] * 1e-3)) <= 1e-6 * b["pairs_per_s"]
    assert abs(b["value"] - evals / (b["ms_per_step"] * 1e-3)) <= 1e-6 * b["value"]


def test_reference_arm_other_ranks_stay_silent():
    """Under torchrun only rank 0 runs and prints the reference arm; the others exit 0 without work."""
    assert _run(["--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
