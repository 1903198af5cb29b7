#!/bin/bash
mkdir -p gpurun_out/j12
python -m pytest tests -m gpu -x -q > gpurun_out/j12/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j12/pytest.log
tail -3 gpurun_out/j12/pytest.log
: > gpurun_out/j12/sweep.log
for v in "" _floor128 _floor64 _hot2 _hot1 _minb4; do
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so python tools/raster_sweep.py --config c3 --tiles 128x16 --pend 4 --refill 6,12 >> gpurun_out/j12/sweep.log 2>&1
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so python tools/raster_sweep.py --config c3 --scale 0.01 --tiles 128x16 --pend 4 --refill 12 >> gpurun_out/j12/sweep.log 2>&1
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so python tools/raster_sweep.py --config c2 --tiles 64x32,32x32 --pend 4 --refill 12 >> gpurun_out/j12/sweep.log 2>&1
done
python tools/raster_sweep.py --config c3 --tiles 128x16,128x8,256x8,256x4,64x16,128x32 --pend 4 --refill 4,8,16 >> gpurun_out/j12/sweep.log 2>&1
python tools/raster_sweep.py --config c2 --tiles 128x8,64x16,256x4 --pend 4 --refill 6,12 >> gpurun_out/j12/sweep.log 2>&1
python tools/raster_sweep.py --config c3 --tiles 128x16 --pend 2,6 --refill 8 >> gpurun_out/j12/sweep.log 2>&1
B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_stats.so python tools/raster_stats.py --config c3 > gpurun_out/j12/stats_c3.log 2>&1
python - <<'Q'
import json
for l in open("gpurun_out/j12/sweep.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["lib"],d["config"],d["tile"],"pend",d["pend"],"refill",d["refill"],"raster",d["raster_kernel"],"frame",d["frame"],d["same_image"])
    else: print(l.rstrip()[:200])
Q
tail -4 gpurun_out/j12/stats_c3.log
