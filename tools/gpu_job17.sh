#!/bin/bash
# request-ring raster kernel: parity first, then sweeps
O=gpurun_out/j17; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
: > $O/sweep.log
S="python tools/raster_sweep.py"
$S --config c3 --tiles 128x8 --pend 32,16,24 --refill 12,8,16 >> $O/sweep.log 2>&1
B200R_W8=1 $S --config c3 --tiles 128x8 --pend 32 --refill 12 >> $O/sweep.log 2>&1
$S --config c3 --tiles 128x16,256x8,256x4,64x16 --pend 32 --refill 12 >> $O/sweep.log 2>&1
$S --config c3 --scale 0.01 --tiles 128x8,256x4 --pend 32,8 --refill 12 >> $O/sweep.log 2>&1
$S --config c2 --tiles 64x16,128x8,64x32 --pend 32,16 --refill 12 >> $O/sweep.log 2>&1
for v in _minb4 _round8 _round2; do
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so $S --config c3 --tiles 128x8,128x16 --pend 32 --refill 12 >> $O/sweep.log 2>&1
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so $S --config c2 --tiles 64x16 --pend 32 --refill 12 >> $O/sweep.log 2>&1
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so $S --config c3 --scale 0.01 --tiles 128x8 --pend 32 --refill 12 >> $O/sweep.log 2>&1
done
python - <<'Q'
import json
for l in open("gpurun_out/j17/sweep.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["lib"],d["config"],d["tile"],"pend",d["pend"],"refill",d["refill"],"raster",d["raster_kernel"],"frame",d["frame"],d["same_image"])
    else: print(l.rstrip()[:200])
Q
