#!/bin/bash
# raster_kernel: span records of the chunk WARPS tickets ahead are prefetched (default) against no prefetch (nopf)
O=gpurun_out/j37; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --frames 12 --tiles 0x0"
: > $O/sweep.log
for lib in libb200raster_nopf.so libb200raster.so; do
  export B200R_LIB=$PWD/cpu_renderer_b200/$lib
  for cfg in c2 c3 c1 c5; do echo "== $lib $cfg" >> $O/sweep.log; $S --config $cfg >> $O/sweep.log 2>&1; done
  echo "== $lib c3 scale 0.01" >> $O/sweep.log; $S --config c3 --scale 0.01 >> $O/sweep.log 2>&1
  echo "== $lib c2 phong" >> $O/sweep.log; $S --config c2 --phong >> $O/sweep.log 2>&1
  echo "== $lib c2 textured" >> $O/sweep.log; $S --config c2 --textured >> $O/sweep.log 2>&1
  echo "== $lib c4 band 3/8" >> $O/sweep.log; $S --config c4 --band 3/8 --frames 4 >> $O/sweep.log 2>&1
done
unset B200R_LIB
grep -E "^==|^\{" $O/sweep.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   setup',d['setup_kernel'],'scan',d['tile_scan_kernel'],'scatter',d['scatter_kernel'],'raster',d['raster_kernel'],'frame',d['frame'],'same',d['same_image'])
    else: print(l.rstrip())
"
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
