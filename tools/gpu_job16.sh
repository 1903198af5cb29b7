#!/bin/bash
# raster lane statistics (clock build) for C3, the 4K floor and C2
O=gpurun_out/j16; mkdir -p $O
for a in "c3 1.0 128x8" "c3 0.01 128x8" "c2 1.0 64x16"; do set -- $a
  for v in stats stats1; do B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_$v.so python tools/raster_stats.py --config $1 --scale $2 --tile $3 > $O/${v}_$1_$2.log 2>&1; tail -4 $O/${v}_$1_$2.log; done
done
