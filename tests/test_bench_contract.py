"""bench.py output contract: the reference arm runs here (CPU only, bounded sample) and must print ONE JSON line with the
agreed keys; the committed GPU bench line (profiles/r1_bench_ours.json) is checked for the same schema plus the GPU-only keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline")


def _check_common(d):
    for k in COMMON:
        assert k in d, k
    assert d["unit"] == "QP/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no throughput for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["value"] > 0 and d["ms_per_step"] > 0


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]


@pytest.mark.parametrize("config", [2, 3, 4])
def test_reference_arm_other_configs(config):
    """--config 2/3/4 (BASELINE configs[2..4]): bounded CPU samples, same contract line, same `config` object as our arm."""
    sys.path.insert(0, ROOT)
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(config), "--steps", "1",
                          "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
    for k in COMMON:
        assert k in d, k
    assert d["impl"] == "reference" and d["scaling"] == bench.CONFIGS[config]["scaling"]
    assert d["config"] == bench.config_object(dict(bench.CONFIGS[config], id=config), 2)     # identical object in both arms


def test_committed_gpu_bench_line_schema():
    path = os.path.join(ROOT, "profiles", "r1_bench_ours.json")
    if not os.path.exists(path):
        pytest.skip("no committed GPU bench line")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    _check_common(d)
    assert d["gpu_launches"] >= d["steps"] and d["data"] == "synthetic"
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not any(x in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for x in c["reasons"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
