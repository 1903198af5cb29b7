"""The benchmark contract on the CPU: `bench.py --impl reference` (the reference's own scalar path on
the host cores; no GPU involved) must print exactly ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(*extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.02",
           "--steps", "1", "--warmup", "1", "--ref-sample", "3000", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout                     # library banners and logs go to stderr
    return json.loads(lines[0])


@pytest.mark.parametrize("config,unit", [("c2", "Mtriangles/s"), ("c3", "Mpixels/s")])
def test_reference_arm_prints_one_contract_line(config, unit):
    d = run_reference("--config", config)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == unit and d["unit"] == unit
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    # timing rule of the contract: at least 3 warm-up steps, whatever --warmup says; the line reports what ran
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if ol.ref_available() else "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]


def test_reference_arm_other_ranks_stay_silent():
    """Under torchrun only rank 0 runs and prints the reference arm; the others exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(not ol.ref_available(), reason="the threaded port has no Phong / textured variant")
def test_reference_arm_textured_phong_variant():
    d = run_reference("--config", "c2", "--textured", "--phong")
    assert "Bitmap" in d["cpu_baseline"]["sample"] and "PhongShading" in d["cpu_baseline"]["sample"]
