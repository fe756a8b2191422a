"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the required keys
(run here on a tiny cube so it takes seconds), and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "8", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, check=True)
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] == 0 and len(d["step_seconds"]) == 2   # stock DoTimestep runs, from rest, no extrapolation
    assert abs(d["ms_per_step"] * d["steps"] - 1e3 * sum(d["step_seconds"])) < 1e-6 * d["ms_per_step"] * d["steps"] + 1e-9
    assert d["impl"] == "reference" and d["metric"] == "fem_steps_per_s" and d["unit"] == "steps/s" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--nx", "8"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_line_carries_the_whole_contract():
    """The last B200 line of the round (profiles/): every key the measurement contract names, internally consistent."""
    with open(os.path.join(ROOT, "profiles", "r02_bench_1gpu_10Mtets_v5.json")) as fh:
        d = json.load(fh)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "solver_variant", "config5_50M", "config4_batch",
                "config2_1M"):
        assert key in d, key
    assert d["metric"] == "fem_steps_per_s" and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "10110954 tets" in d["config"]["workload"] and "model" not in d["config"]
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
    assert abs(r["traffic"] - r["algorithmic_bytes_per_launch"]) < 0.01 * r["traffic"]       # ncu DRAM bytes = the byte model
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 1000
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] < 0.01
    v = d["solver_variant"]
    assert v["value"] > 10 * d["value"] and max(v["cg_iterations_per_step"]) < 30 and v["final_displacement_rel_diff_vs_parity_path"] < 1e-5
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
