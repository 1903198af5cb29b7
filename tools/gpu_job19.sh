#!/bin/bash
# fused gather tests + whole GPU suite + floor with per-tile chunk sizing
O=gpurun_out/j19; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gather.py -x -q > $O/pytest_gather.log 2>&1; echo "gather rc=$?"; tail -15 $O/pytest_gather.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python tools/raster_sweep.py --config c3 --scale 0.01 --tiles 128x8 --pend 4 --refill 12 > $O/sweep.log 2>&1
python tools/raster_sweep.py --config c3 --tiles 128x8 --pend 4 --refill 12 >> $O/sweep.log 2>&1
python tools/raster_sweep.py --config c2 --tiles 64x16 --pend 4 --refill 12 >> $O/sweep.log 2>&1
python tools/raster_sweep.py --config c4 --scale 0.1 --tiles 128x8 --pend 4 --refill 12 >> $O/sweep.log 2>&1
cat $O/sweep.log | cut -c1-300
