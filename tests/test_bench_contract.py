"""CPU: the JSON line of bench.py's reference arm (the reference's multiprocessing CPU path as restated by the
oracle port) carries every key of the measurement contract; argument handling of the strong-scaling option."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    return out


def test_reference_arm_line_has_the_contract_keys():
    out = _bench("--impl", "reference", "--workload", "sunspot", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "replica_mcmc_steps_per_sec" and line["unit"] == "replica-steps/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["dtype"] == "f64" and "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_strong_scaling_option_must_split_evenly():
    out = _bench("--impl", "reference", "--gpus", "3", "--ladder-total", "1024", "--steps", "1", "--warmup", "0")
    assert out.returncode != 0 and "does not split evenly" in (out.stderr + out.stdout)


import pytest


@pytest.mark.gpu
def test_b200_arm_line_has_the_contract_keys():
    """The B200 arm on the reference's own Sunspot configuration (BASELINE configs[0]; seconds)."""
    out = _bench("--workload", "sunspot", "--steps", "3", "--warmup", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "result_pipeline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["gpu_launches"] == 3
    assert line["value"] > 1e4 and 0 < line["e2e"]["value"] <= line["value"] * 1.05
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    rf = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and line["value"] > 20 * cb["value"]
    assert line["clocks"]["sm_mhz"] > 0 and not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    rp = line["result_pipeline"]
    assert rp["bound"] == "hbm" and rp["achieved"] > 0 and rp["peak"] > 0
