#!/bin/bash
mkdir -p gpurun_out/j4
python -m pytest tests -m gpu -x -q > gpurun_out/j4/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j4/pytest.log
tail -5 gpurun_out/j4/pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j4/bench_default.json 2> gpurun_out/j4/bench_default.err ) 2> gpurun_out/j4/bench_default.time
echo "bench rc=$?"; tail -3 gpurun_out/j4/bench_default.time; tail -5 gpurun_out/j4/bench_default.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-legs --no-cpu-baseline --defer-verdict > gpurun_out/j4/bench_defer.json 2> gpurun_out/j4/bench_defer.err
python - <<'P'
import json
for f in ("bench_default","bench_defer"):
    try:
        d=json.load(open(f"gpurun_out/j4/{f}.json"))
        print(f, d["ms_per_step"], d["ms_per_step_best"], d["value"], d.get("image_ok"), d["e2e"]["ms_per_step"], d["stage_ms"])
        for k,l in d.get("legs",{}).items():
            print("  leg",k,l["ms_per_step"],l["value"],l["image_ok"],l["roofline"]["frac"],l["stage_ms"], l.get("e2e",{}).get("ms_per_step"))
    except Exception as e: print(f,"failed",e)
P
