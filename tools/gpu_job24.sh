#!/bin/bash
O=gpurun_out/j24; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12"
: > $O/sweep.log
run() { echo "cfg $1 scale $2 tpc $3 rows $4" >> $O/sweep.log; B200R_SPLIT=2 B200R_SPLIT_TPC=$3 B200R_SPLIT_ROWS=$4 $S --config $1 --scale $2 --tiles 128x8 >> $O/sweep.log 2>&1; }
for pr in "32 64" "64 32" "64 64" "64 128" "32 128"; do set -- $pr; run c3 1.0 $1 $2; done
for pr in "16 64" "32 64" "32 32" "16 32"; do set -- $pr; run c3 0.2 $1 $2; done
for pr in "16 64" "32 64" "32 32" "16 32" "64 64"; do set -- $pr; run c3 0.5 $1 $2; done
B200R_SPLIT=0 $S --config c3 --scale 0.5 --tiles 128x8 >> $O/sweep.log 2>&1
for pr in "8 16" "8 32" "16 16"; do set -- $pr; run c3 0.05 $1 $2; done
B200R_SPLIT=0 $S --config c3 --scale 0.05 --tiles 128x8 >> $O/sweep.log 2>&1
python - <<'Q'
import json
for l in open("gpurun_out/j24/sweep.log"):
    if l.startswith("{"):
        d=json.loads(l); print("   setup",d["setup_kernel"],"frame",d["frame"],d["same_image"])
    else: print(l.rstrip()[:200])
Q
