#!/bin/bash
# select_kernel as persistent CTAs with the next block's positions in flight
O=gpurun_out/j40; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --tiles 0x0"
: > $O/sweep.log
for b in 3/8 0/8 1/2; do echo "== c4 band $b" >> $O/sweep.log; $S --config c4 --band $b --frames 4 >> $O/sweep.log 2>&1; done
grep -E "^==|^\{" $O/sweep.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   setup',d['setup_kernel'],'scan',d['tile_scan_kernel'],'scatter',d['scatter_kernel'],'raster',d['raster_kernel'],'frame',d['frame'],'same',d['same_image'])
    else: print(l.rstrip())
"
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:select_kernel --log-file $O/select.csv $S --config c4 --band 3/8 --frames 1 > $O/ncu.log 2>&1
grep select_kernel $O/select.csv | tail -2 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gather.py -m gpu -x -q -k "preselection or band or c4 or gather" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
