#!/bin/bash
mkdir -p gpurun_out/j3
for v in r4 r4c r4c16 r4m16 r2c; do
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_$v.so python tools/raster_sweep.py --config c3 --tiles 128x16 --pend 4 --refill 8,12 >> gpurun_out/j3/sweep.log 2>&1
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_$v.so python tools/raster_sweep.py --config c2 --tiles 64x32,32x32 --pend 4 --refill 8 >> gpurun_out/j3/sweep.log 2>&1
done
grep -v "^$" gpurun_out/j3/sweep.log | cut -c1-400
