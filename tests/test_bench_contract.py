"""bench.py's JSON line: the reference arm runs here (it needs no GPU: the oracle on the host cores, enumeration by
the oracle's own buildcachematrix restatement), and the recorded GPU lines under profiles/ are checked for the keys
the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _last_json(text):
    return json.loads([l for l in text.strip().splitlines() if l.startswith("{")][-1])


def test_reference_arm_prints_the_contract_line_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--rows", "3000", "--nodes", "9", "--batch", "256", "--cpu-sample", "32"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = _last_json(out.stdout)
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "score-cache GLM fits/sec" and line["unit"] == "fits/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["e2e"] == {"value": line["value"], "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "fits per step" in cb["sample"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_is_silent_on_the_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_recorded_gpu_lines_carry_the_driver_keys():
    for name, n in (("bench_r2_n1.json", 1), ("bench_r2_n2.json", 2)):
        line = _last_json(open(os.path.join(ROOT, "profiles", name)).read())
        assert BASE_KEYS | {"roofline", "roofline_fp64", "gpu_launches", "clocks"} <= set(line), name
        assert line["n_gpus"] == n and line["scaling"] == "strong" and line["dtype"] == "f64"
        r = line["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        f = line["roofline_fp64"]
        assert abs(f["frac"] - f["achieved"] / f["peak"]) < 1e-9 and 0.3 < f["frac"] < 1.0
        e = line["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != line["value"]
        assert line["gpu_launches"] > 0 and line["clocks"]["sm_mhz"] > 0
        assert {"exact_path", "default_path_same_batch", "bayes"} <= set(line)
        if n == 1:
            assert line["cpu_baseline"]["status_agree"] is True
            assert line["cpu_baseline"]["max_rel_diff_gpu_vs_oracle_on_sample"] < 1e-8
