#!/bin/bash
# compile-time knobs of the raster kernel, one variant each, on C3 / C2 / 500-triangle frame
O=gpurun_out/j43; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --frames 10 --tiles 0x0"
: > $O/sweep.log
for lib in libb200raster.so libb200raster_hot8.so libb200raster_hot2.so libb200raster_fe128.so libb200raster_fe64.so libb200raster_bk16.so libb200raster_ct5.so libb200raster_rd8.so; do
  export B200R_LIB=$PWD/cpu_renderer_b200/$lib
  echo "== $lib c3" >> $O/sweep.log; $S --config c3 >> $O/sweep.log 2>&1
  echo "== $lib c2" >> $O/sweep.log; $S --config c2 >> $O/sweep.log 2>&1
done
grep -E "^==|^\{" $O/sweep.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   setup',d['setup_kernel'],'scatter',d['scatter_kernel'],'raster',d['raster_kernel'],'frame',d['frame'],'same',d['same_image'])
    else: print(l.rstrip())
"
