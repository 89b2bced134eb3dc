"""bench.py contract checks that need no GPU: the reference arm runs on host cores only, and the committed round-1
line carries every key the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_on_cpu_and_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"].startswith("EM link-updates/sec") and rec["unit"] == "link-updates/s"
    assert rec["higher_is_better"] is True and rec["value"] > 0 and rec["gpu_launches"] == 0
    assert rec["steps"] == 2 and rec["warmup"] == 1          # the arm does the warm-up and step counts it was asked for
    # oracle/_ref (the unmodified reference script) is there in the build container and wherever the tree travelled with it
    want = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "TrigenicInteractionPredictor.py")) else "port"
    assert rec["cpu_baseline"]["kind"] == want and rec["cpu_baseline"]["cores"] >= 1 and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


import pytest


@pytest.mark.parametrize("name", ["r1_bench_line.json", "r2_bench_line.json"])
def test_committed_bench_line_has_the_contract_keys(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        rec = json.loads(fh.read().strip().splitlines()[-1])
    if name.startswith("r2"):
        assert rec["cpu_baseline"]["kind"] == "reference" and rec["config"]["shape"] == "kuzmin"
        assert rec["roofline"]["kernel_ms"] <= rec["ms_per_step"] and rec["roofline"]["binding"]["frac"] < 1.0
        for key in ("cfg4_strong", "cfg3_samples", "value_uniform", "value_kuzmin"):
            assert key in rec, key
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in rec, key
    assert rec["dtype"] == "f64" and rec["vs_baseline"] is None and "workload" in rec["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in rec["roofline"], key
    assert abs(rec["roofline"]["frac"] - rec["roofline"]["achieved"] / rec["roofline"]["peak"]) < 1e-9
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in rec["cpu_baseline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in rec["e2e"], key
    assert rec["e2e"]["h2d_bytes_per_step"] > 0 and rec["e2e"]["value"] < rec["value"]
    assert rec["gpu_launches"] >= rec["steps"]
