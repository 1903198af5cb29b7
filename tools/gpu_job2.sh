#!/bin/bash
mkdir -p gpurun_out/j2
for v in "" _r4; do
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 \
    -o gpurun_out/j2/c3_raster$v -f python tools/raster_sweep.py --config c3 --tiles 128x16 --pend 4 --refill 8 --frames 2 > gpurun_out/j2/ncu$v.log 2>&1
done
ls -la gpurun_out/j2
