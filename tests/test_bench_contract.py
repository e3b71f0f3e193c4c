"""CPU: the bench line committed under profiles/ (the last default `python bench.py` run on a B200 this round) carries
every key of the measurement contract, and bench.py's reference arm prints a contract-shaped line without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_common(d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "images/sec" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k


def test_committed_bench_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r01_bench20_default.json")) as f:
        d = json.loads(f.readline())
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * 3 * 513 * 513 * 4 + 16 * 513 * 513 * 4 and d["e2e"]["d2h_bytes_per_step"] == 4
    assert 0 < d["e2e"]["value"] <= d["value"] * 1.02
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["step2"]["unit"] == "images/sec" and d["step2"]["generator_updates_per_step"] > 0


def test_reference_arm_line_shape():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
