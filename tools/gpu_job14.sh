#!/bin/bash
mkdir -p gpurun_out/j14
python -m pytest tests -m gpu -x -q > gpurun_out/j14/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j14/pytest.log
tail -3 gpurun_out/j14/pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j14/bench_default.json 2> gpurun_out/j14/bench_default.err ) 2> gpurun_out/j14/bench_default.time
echo "bench rc=$?"; tail -3 gpurun_out/j14/bench_default.time; tail -5 gpurun_out/j14/bench_default.err
python - <<'Q'
import json
d=json.load(open("gpurun_out/j14/bench_default.json"))
print(d["ms_per_step"], d["value"], d.get("image_ok"), d["e2e"]["ms_per_step"], d["stage_ms"], d["config"]["tile"])
for k,l in d.get("legs",{}).items():
    print("  leg",k,l["ms_per_step"],l["value"],l["image_ok"],l["roofline"]["frac"],l["stage_ms"], l.get("e2e",{}).get("ms_per_step"), l["config"]["tile"])
Q
python bench.py --config c3 --scale 0.01 --no-legs --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/j14/bench_c3_floor.json 2>gpurun_out/j14/bench_c3_floor.err
python -c "
import json; d=json.load(open('gpurun_out/j14/bench_c3_floor.json')); print('floor', d['ms_per_step'], d['stage_ms'])"
