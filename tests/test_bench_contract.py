"""CPU: the bench line committed under profiles/ (the last default `python bench.py` run on a B200 this round) carries
every key of the measurement contract, and bench.py's reference arm prints a contract-shaped line without a GPU."""
import json
import os
import subprocess
import sys

import pytest

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


@pytest.mark.parametrize("name", ["r01_bench20_default.json", "r02_bench_final.json"])
def test_committed_bench_line_has_the_contract_keys(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        d = json.loads(f.readline())
    _check_common(d)
    if name.startswith("r02"):
        _check_round2_blocks(d)
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


def _check_round2_blocks(d):
    """blocks added in round 2: the reference's GPU path on the same B200, the tolerance-meeting training mode with its
    distance to fp64, step 2 with e2e + baselines, config 5, the device input transforms"""
    lib = d["library_baseline"]
    for arm in ("tf32_reference_defaults", "tf32_cudnn_benchmark_channels_last", "bf16_autocast_channels_last"):
        assert lib[arm]["value"] > 0 and lib[arm]["unit"] == "images/sec"
        assert d["value"] > lib[arm]["value"]                       # the bf16 path beats every stock-PyTorch arm
    par = d["parity_mode"]
    assert par["split2"]["value"] > lib["tf32_reference_defaults"]["value"]   # the tolerance-meeting mode beats cuDNN TF32
    arms = d["numerics_vs_fp64"]["arms"]
    assert arms["zs3_b200_split3"]["logits_eval_configs0"] < 1e-3 and arms["zs3_b200_split2"]["logits_eval_configs0"] < 1e-3
    assert arms["zs3_b200_split3"]["grads_train"] < arms["torch_fp32_no_tf32"]["grads_train"] * 1.5
    s2 = d["step2"]
    for k in ("e2e", "e2e_label_table_api", "library_baseline", "cpu_baseline", "roofline", "segments_ms", "unchanged_trainer_loop"):
        assert k in s2, k
    assert s2["e2e_label_table_api"]["h2d_bytes_per_step"] == 16 * 3 * 513 * 513 * 4 + 16 * 513 * 513 * 4
    assert s2["value"] > 20 * s2["library_baseline"]["value"]
    assert s2["unchanged_trainer_loop"]["host_noise_draw_ms_per_step"] > 0.5 * s2["unchanged_trainer_loop"]["ms_per_step"]
    c5 = d["config5"]
    assert c5["value"] > 0 and c5["graph_generator_updates_per_step"] > 0 and "segments_ms" in c5
    tf = d["input_transforms"]
    assert tf["bit_exact_vs_pillow"] is True and tf["unit"] == "pictures/sec"
    for k in ("bound", "achieved", "peak", "unit", "frac"):
        assert k in tf["roofline"], k
    assert tf["roofline"]["bound"] == "hbm" and tf["cpu_baseline"]["kind"] == "reference"
    assert tf["e2e"]["h2d_bytes_per_step"] > 0 and tf["value"] > 100 * tf["cpu_baseline"]["value"]
    assert d["roofline"]["traffic"] is not None and "r02_conv_traffic" in d["roofline"]["traffic_source"]
    fo = d["forward_only"]
    assert fo["eval_mode_bn_fused_epilogue_cuda_graph"]["frac_of_measured_bf16_peak"] > 0.37


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
