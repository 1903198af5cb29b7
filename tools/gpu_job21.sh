#!/bin/bash
# row-parallel set-up (SPLIT): parity, then sweeps
O=gpurun_out/j21; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -8 $O/pytest.log
S="python tools/raster_sweep.py --pend 4 --refill 12"
: > $O/sweep.log
for cfg in "c3 1.0" "c3 0.01" "c3 0.2"; do set -- $cfg
  B200R_SPLIT=0 $S --config $1 --scale $2 --tiles 128x8 >> $O/sweep.log 2>&1
  $S --config $1 --scale $2 --tiles 128x8 >> $O/sweep.log 2>&1
  for tpc in 16 32 64; do for rows in 16 32 64; do
    echo "tpc $tpc rows $rows" >> $O/sweep.log
    B200R_SPLIT_TPC=$tpc B200R_SPLIT_ROWS=$rows $S --config $1 --scale $2 --tiles 128x8 >> $O/sweep.log 2>&1
  done; done
done
$S --config c2 --tiles 64x16 >> $O/sweep.log 2>&1
$S --config c2 --scale 0.1 --tiles 64x16 >> $O/sweep.log 2>&1
B200R_SPLIT=0 $S --config c2 --scale 0.1 --tiles 64x16 >> $O/sweep.log 2>&1
$S --config c1 --tiles 64x16 >> $O/sweep.log 2>&1
B200R_SPLIT=0 $S --config c1 --tiles 64x16 >> $O/sweep.log 2>&1
python - <<'Q'
import json
for l in open("gpurun_out/j21/sweep.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["config"],d["tile"],"setup",d["setup_kernel"],"scatter",d["scatter_kernel"],"raster",d["raster_kernel"],"frame",d["frame"],d["same_image"])
    else: print(l.rstrip()[:200])
Q
