#!/bin/bash
# Round-2 evidence at HEAD (1 GPU): GPU suite, bench lines of every config / shading mode, reference arm,
# launch list, full captures.
O=gpurun_out/j26; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo "default rc=$?"; tail -3 $O/bench_default.time
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c2_reference.json 2> $O/bench_c2_reference.err; echo "reference rc=$?"
for cfg in c1 c3 c5; do python bench.py --config $cfg --steps 20 --warmup 5 --no-legs > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"; done
for cfg in c2 c3; do
  python bench.py --config $cfg --phong --steps 20 --warmup 5 --no-legs --no-cpu-baseline > $O/bench_${cfg}_phong.json 2> $O/bench_${cfg}_phong.err
  python bench.py --config $cfg --textured --steps 20 --warmup 5 --no-legs --no-cpu-baseline > $O/bench_${cfg}_textured.json 2> $O/bench_${cfg}_textured.err
  python bench.py --config $cfg --textured --phong --steps 20 --warmup 5 --no-legs --no-cpu-baseline > $O/bench_${cfg}_textured_phong.json 2> $O/bench_${cfg}_textured_phong.err
done
for sc in 0.01 0.05 0.2; do python bench.py --config c3 --scale $sc --no-legs --no-cpu-baseline --steps 20 --warmup 5 > $O/bench_c3_scale$sc.json 2>$O/bench_c3_scale$sc.err; done
python - <<'Q'
import json,glob
for f in sorted(glob.glob("gpurun_out/j26/bench_*.json")):
    try: d=json.load(open(f))
    except Exception as e: print(f, "unreadable", e); continue
    if d.get("impl")=="reference": print(f, d["value"], d["unit"]); continue
    print(f.split("/")[-1], round(d["ms_per_step"],4), round(d["value"],1), d.get("image_ok"), "e2e", round(d.get("e2e",{}).get("ms_per_step",0),3), {k:round(v,4) for k,v in d["stage_ms"].items()}, d["config"]["tile"], round(d["roofline"]["frac"],4))
    for k,l in d.get("legs",{}).items():
        print("  leg",k,round(l["ms_per_step"],4),round(l["value"],1),l["image_ok"],round(l["roofline"]["frac"],4),{a:round(b,4) for a,b in l["stage_ms"].items()}, l.get("e2e",{}).get("ms_per_step"))
Q
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-legs --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'setup_kernel|scatter_kernel|raster_kernel|tile_scan|lookback_scan|zrange_kernel' -s 15 -c 5 \
    -o $O/c2_kernels -f python tools/raster_sweep.py --config c2 --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'setup_kernel|raster_kernel' -s 6 -c 2 \
    -o $O/c3_kernels -f python tools/raster_sweep.py --config c3 --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel' -s 3 -c 1 \
    -o $O/c3_raster_textured -f python tools/raster_sweep.py --config c3 --textured --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_c3t.log 2>&1; echo "ncu c3 textured rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel' -s 3 -c 1 \
    -o $O/c3_raster_phong -f python tools/raster_sweep.py --config c3 --phong --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_c3p.log 2>&1; echo "ncu c3 phong rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'setup_kernel' -s 3 -c 1 \
    -o $O/floor_setup_split -f python tools/raster_sweep.py --config c3 --scale 0.01 --tiles 0x0 --pend 4 --refill 12 --frames 2 > $O/ncu_floor.log 2>&1; echo "ncu floor rc=$?"
ls -la $O | head -50
