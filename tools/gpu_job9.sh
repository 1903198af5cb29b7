#!/bin/bash
mkdir -p gpurun_out/j9
tools/micro/atoms_bench > gpurun_out/j9/atoms.log 2>&1
export B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_clocks.so
python tools/raster_stats.py --config c3 > gpurun_out/j9/c3.log 2>&1
python tools/raster_stats.py --config c3 --scale 0.01 > gpurun_out/j9/c3_001.log 2>&1
python tools/raster_stats.py --config c2 --tile 64x32 > gpurun_out/j9/c2.log 2>&1
cat gpurun_out/j9/atoms.log; tail -n 4 gpurun_out/j9/c*.log
