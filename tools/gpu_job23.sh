#!/bin/bash
O=gpurun_out/j23; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:'setup_kernel' -s 3 -c 1 \
    -o $O/c3_setup_split -f python tools/raster_sweep.py --config c3 --tiles 128x8 --pend 4 --refill 12 --frames 2 > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
