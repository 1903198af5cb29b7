#!/bin/bash
# Round-2 evidence at HEAD: GPU suite, default bench line, launch list, full captures of the C2 and C3 kernels.
O=gpurun_out/j15; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time
echo "bench rc=$?"; tail -3 $O/bench_default.time
python bench.py --config c3 --steps 20 --warmup 5 --no-legs > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?"
python bench.py --config c3 --scale 0.01 --no-legs --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_c3_floor.json 2>$O/bench_c3_floor.err
python - <<'Q'
import json
for f in ("bench_default","bench_c3","bench_c3_floor"):
    d=json.load(open(f"gpurun_out/j15/{f}.json"))
    print(f, d["ms_per_step"], d["value"], d.get("image_ok"), d.get("e2e",{}).get("ms_per_step"), d["stage_ms"], d["config"]["tile"], d["roofline"]["frac"])
    for k,l in d.get("legs",{}).items():
        print("  leg",k,l["ms_per_step"],l["value"],l["image_ok"],l["roofline"]["frac"],l["stage_ms"], l.get("e2e",{}).get("ms_per_step"), l["config"]["tile"])
Q
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-legs --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'setup_kernel|scatter_kernel|raster_kernel|tile_scan|lookback_scan|zrange_kernel' -s 15 -c 5 \
    -o $O/c2_kernels -f python tools/raster_sweep.py --config c2 --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'setup_kernel|raster_kernel' -s 6 -c 2 \
    -o $O/c3_kernels -f python tools/raster_sweep.py --config c3 --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la $O
