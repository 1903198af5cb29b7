#!/bin/bash
O=gpurun_out/j18; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel' -s 3 -c 1 \
    -o $O/c3_raster -f python tools/raster_sweep.py --config c3 --tiles 128x8 --frames 2 > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_stats1.so python tools/raster_stats.py --config c3 --tile 128x8 > $O/stats1_c3.log 2>&1; tail -4 $O/stats1_c3.log
