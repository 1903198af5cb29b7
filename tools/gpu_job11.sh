#!/bin/bash
# ncu --set full of the raster kernel of a build variant:  tools/gpu_job11.sh <variant-suffix> <config> <tile> [scale]
mkdir -p gpurun_out/j11
v=$1; cfg=$2; tile=$3; scale=${4:-1.0}
B200R_LIB=$PWD/cpu_renderer_b200/libb200raster$v.so ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 \
    -o gpurun_out/j11/${cfg}_raster$v -f python tools/raster_sweep.py --config $cfg --scale $scale --tiles $tile --pend 4 --refill 12 --frames 2 > gpurun_out/j11/ncu_${cfg}$v.log 2>&1
tail -3 gpurun_out/j11/ncu_${cfg}$v.log
