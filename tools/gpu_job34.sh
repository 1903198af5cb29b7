#!/bin/bash
# raster_kernel: idle lanes share the spans on hand (default) against one lane per span (nosplit)
O=gpurun_out/j34; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --frames 12 --tiles 0x0"
: > $O/sweep.log
for lib in libb200raster_nosplit.so libb200raster.so; do
  export B200R_LIB=$PWD/cpu_renderer_b200/$lib
  for sc in 0.01 0.05 0.2 1.0; do echo "== $lib c3 scale $sc" >> $O/sweep.log; $S --config c3 --scale $sc >> $O/sweep.log 2>&1; done
  for cfg in c2 c1 c5; do echo "== $lib $cfg" >> $O/sweep.log; $S --config $cfg >> $O/sweep.log 2>&1; done
  echo "== $lib c3 phong" >> $O/sweep.log; $S --config c3 --phong >> $O/sweep.log 2>&1
  echo "== $lib c3 textured" >> $O/sweep.log; $S --config c3 --textured >> $O/sweep.log 2>&1
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
