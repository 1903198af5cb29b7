#!/bin/bash
mkdir -p gpurun_out/j8
export B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_stats.so
python tools/raster_stats.py --config c3 > gpurun_out/j8/c3.log 2>&1
python tools/raster_stats.py --config c3 --scale 0.01 > gpurun_out/j8/c3_001.log 2>&1
python tools/raster_stats.py --config c3 --scale 0.2 > gpurun_out/j8/c3_02.log 2>&1
python tools/raster_stats.py --config c2 --tile 64x32 > gpurun_out/j8/c2.log 2>&1
python tools/raster_stats.py --config c2 --tile 32x32 > gpurun_out/j8/c2_32.log 2>&1
tail -n 5 gpurun_out/j8/*.log
