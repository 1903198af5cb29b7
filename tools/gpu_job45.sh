#!/bin/bash
# ncu --set full of the positions-only passes at HEAD: select_kernel (C4 band) and zrange_kernel (C5)
O=gpurun_out/j45; mkdir -p $O; R=/tmp/reps; mkdir -p $R
S="python tools/raster_sweep.py --pend 4 --refill 12 --tiles 0x0 --frames 1"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'select_kernel' -s 2 -c 1 -o $R/select -f $S --config c4 --band 3/8 > $O/ncu_select.log 2>&1
$NCU -k regex:'zrange_kernel' -s 2 -c 1 -o $R/zrange -f $S --config c5 > $O/ncu_zrange.log 2>&1
python profiles/summarize.py $R/select.ncu-rep $O/select.txt > /dev/null 2>&1
python profiles/summarize.py $R/zrange.ncu-rep $O/zrange.txt > /dev/null 2>&1
cat $O/select.txt $O/zrange.txt > $O/r02_ncu_select_zrange.txt; head -26 $O/select.txt | cut -c1-160; grep -A12 "== zrange" $O/zrange.txt | cut -c1-160
