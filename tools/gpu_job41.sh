#!/bin/bash
# zrange_kernel with 128-bit loads; then the whole GPU suite and the default bench line at HEAD
O=gpurun_out/j41; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --tiles 0x0 --frames 12"
: > $O/sweep.log
for cfg in c2 c3 c5; do echo "== $cfg" >> $O/sweep.log; $S --config $cfg >> $O/sweep.log 2>&1; done
grep -E "^==|^\{" $O/sweep.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   setup',d['setup_kernel'],'scan',d['tile_scan_kernel'],'scatter',d['scatter_kernel'],'raster',d['raster_kernel'],'frame',d['frame'],'same',d['same_image'])
    else: print(l.rstrip())
"
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:zrange_kernel --log-file $O/zrange.csv $S --config c5 --frames 1 > $O/ncu.log 2>&1
grep "zrange_kernel" $O/zrange.csv | tail -2 | awk -F'","' '{print $5, $NF}'
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python bench.py > $O/r02_bench_default.json 2> $O/bench.err; python -c "
import json; d=json.loads(open('$O/r02_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['stage_ms'], d['image_ok'], d['e2e']['value'], {k:(v['ms_per_step'], v['image_ok']) for k,v in d['legs'].items()})"
