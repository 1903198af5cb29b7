#!/bin/bash
# setup_kernel: direct colour loads + 6 CTAs/SM (variant d6) against HEAD; C3 SPLIT sweep; GPU suite
O=gpurun_out/j31; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --frames 12 --tiles 0x0"
: > $O/sweep.log
for lib in libb200raster.so libb200raster_d6.so; do
  export B200R_LIB=$PWD/cpu_renderer_b200/$lib
  for cfg in c2 c3; do echo "== $lib $cfg" >> $O/sweep.log; $S --config $cfg >> $O/sweep.log 2>&1; done
  echo "== $lib c4 band 3/8" >> $O/sweep.log; $S --config c4 --band 3/8 --frames 4 >> $O/sweep.log 2>&1
  echo "== $lib c2 textured" >> $O/sweep.log; $S --config c2 --textured >> $O/sweep.log 2>&1
done
unset B200R_LIB
for combo in "32 32" "32 64" "64 64" "64 128" "24 64"; do
  set -- $combo
  echo "== split tpc=$1 rows=$2 c3" >> $O/sweep.log
  B200R_SPLIT=2 B200R_SPLIT_TPC=$1 B200R_SPLIT_ROWS=$2 $S --config c3 >> $O/sweep.log 2>&1
done
grep -E "^==|^\{" $O/sweep.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   setup',d['setup_kernel'],'scan',d['tile_scan_kernel'],'scatter',d['scatter_kernel'],'raster',d['raster_kernel'],'frame',d['frame'],'same',d['same_image'])
    else: print(l.rstrip())
"
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > $O/pytest.log 2>&1; tail -15 $O/pytest.log
B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_d6.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_d6.log 2>&1; tail -3 $O/pytest_d6.log
