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


def _archived(name):
    p = os.path.join(ROOT, "profiles", name)
    lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
    assert len(lines) == 1, name
    return json.loads(lines[0])


@pytest.mark.parametrize("name,n", [("r02_bench_default.json", 1), ("r02_bench_default_n2.json", 2), ("r02_bench_default_n8.json", 8)])
def test_archived_gpu_lines_carry_the_contract(name, n):
    """The lines archived under profiles/ (what the README tables are printed from) are ONE JSON line each with
    every key of the contract, consistent with each other, and every image hash equal to the oracle's."""
    d = _archived(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "gpu_launches", "e2e", "roofline", "clocks", "stage_ms", "legs"):
        assert key in d, key
    assert d["n_gpus"] == n and d["metric"] == d["unit"] == "Mtriangles/s" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = the units all ranks processed / the max-over-ranks time
    assert abs(d["value"] - n * d["config"]["triangles"] / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] and r["traffic"] > r["algorithmic_bytes_per_launch"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["image_ok"] is True and e["image_ok"] is True
    c4 = d["legs"]["c4_bands"]
    assert c4["image_ok"] is True and c4["scaling"] == "strong" and c4["config"]["band_rows"] * n >= 16384
    if n == 1:
        assert d["legs"]["c3"]["image_ok"] is True and "issue" in d["legs"]["c3"]["roofline"]
        cb = d["cpu_baseline"]
        assert cb["kind"] == "reference" and cb["scalar_mt"]["whole_frame"] is True and "avx_mt" in cb
    else:
        assert d["with_gather"]["images_match_the_ranks_frames"] is True
        assert c4["with_gather"]["ms_per_step"] < c4["with_gather_nccl"]["ms_per_step"]
