#!/bin/bash
O=gpurun_out/j25; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12"
: > $O/sweep.log
for sc in 1.0 0.5; do
echo "scale $sc TALL=0" >> $O/sweep.log; B200R_TALL=0 $S --config c3 --scale $sc --tiles 128x8 >> $O/sweep.log 2>&1
for sh in 1 2; do for ch in 1 2 4 8 1000000; do
  echo "scale $sc shift $sh chunk $ch" >> $O/sweep.log; B200R_TALL_SHIFT=$sh B200R_TALL_CHUNK=$ch $S --config c3 --scale $sc --tiles 128x8 >> $O/sweep.log 2>&1
done; done; done
for sc in 0.2 0.01; do for ch in 4 1000000; do
  echo "scale $sc chunk $ch (split)" >> $O/sweep.log; B200R_TALL_CHUNK=$ch $S --config c3 --scale $sc --tiles 128x8 >> $O/sweep.log 2>&1
done; done
echo "c1" >> $O/sweep.log; $S --config c1 --tiles 64x16 >> $O/sweep.log 2>&1
echo "c1 TALL=0" >> $O/sweep.log; B200R_TALL=0 $S --config c1 --tiles 64x16 >> $O/sweep.log 2>&1
echo "c2" >> $O/sweep.log; $S --config c2 --tiles 64x16 >> $O/sweep.log 2>&1
python - <<'Q'
import json
for l in open("gpurun_out/j25/sweep.log"):
    if l.startswith("{"):
        d=json.loads(l); print("   setup",d["setup_kernel"],"frame",d["frame"],d["same_image"])
    else: print(l.rstrip()[:200])
Q
