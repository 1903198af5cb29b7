#!/bin/bash
# developer job: parity tests, then raster sweeps (C3, C2)
mkdir -p gpurun_out/j1
python -m pytest tests -m gpu -x -q > gpurun_out/j1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1/pytest.log
tail -3 gpurun_out/j1/pytest.log
python tools/raster_sweep.py --config c3 --tiles 128x16,256x8,64x32 --pend 4,8,12,16 --refill 8 > gpurun_out/j1/sweep_c3_r8.log 2>&1
python tools/raster_sweep.py --config c3 --tiles 128x16 --pend 8 --refill 4,12,16 >> gpurun_out/j1/sweep_c3_r8.log 2>&1
for r in 4 6; do B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_r$r.so python tools/raster_sweep.py --config c3 --tiles 128x16 --pend 4,8,12 --refill 8 > gpurun_out/j1/sweep_c3_r$r.log 2>&1; done
python tools/raster_sweep.py --config c2 --tiles 64x32,128x16,32x32 --pend 4,8,16 --refill 8 > gpurun_out/j1/sweep_c2_r8.log 2>&1
B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_r4.so python tools/raster_sweep.py --config c2 --tiles 64x32 --pend 4,8 --refill 8 > gpurun_out/j1/sweep_c2_r4.log 2>&1
cat gpurun_out/j1/sweep_*.log | grep -v "^$" | tail -60
