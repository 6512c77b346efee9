"""bench.py prints exactly ONE JSON line on stdout with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # library chatter goes to stderr
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the oracle port timed on the host cores, same metric / unit / config as the B200 arm."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "lenseflow_batched_applies_per_sec" and d["unit"] == "applies/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["ranks"]["ran_on"] == "rank 0 only" and cb["fft"] in ("pocketfft", "mkl")


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU reference; every other rank prints nothing and exits 0."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--cg-iters", "2"], 900)
    assert BASE_KEYS <= set(d) and d.get("impl") != "reference"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["gpu_launches"] > 0 and d["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.05 < r["frac"] < 1.2
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] == 8 * 2 * 1024 * 1024 * 8 and 0 < e["synchronous"]["value"] <= e["value"] < 1.02 * d["value"]
    c = d["clocks"]
    assert c["sm_max_mhz"] and c["samples"] >= 1 and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["gpu_vs_oracle_rel_l2"] < 1e-11
    cg = d["cg"]
    assert cg["iters_timed"] >= 20 and cg["value"] > 0 and cg["res_last"] > 0 and cg["cpu_baseline"]["value"] > 0
    mj, hm = d["map_joint"], d["hmc"]
    assert mj["value"] > 0 and mj["higher_is_better"] is False and 0.05 < min(mj["alpha"]) and mj["corr_phi_map_vs_truth_rank0"] > 0.2
    assert hm["value"] > 0 and hm["abs_dH_max_rank0"] < 50
    ff = d["fft"]
    assert ff["rfft2_us"] > 0 and ff["irfft2_us"] > 0 and ff["round_trip_max_abs_err"] < 1e-10 and 0.05 < ff["rfft2_frac_of_measured_peak"] < 1.2
