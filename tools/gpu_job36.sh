#!/bin/bash
# ncu --set full captures at HEAD: C2, C3, a 1/8 row band of C4, the 500-triangle 4K frame; launch list of the bench
O=gpurun_out/j36; mkdir -p $O; R=/tmp/reps; mkdir -p $R
S="python tools/raster_sweep.py --pend 4 --refill 12 --tiles 0x0"
# each command runs once without ncu first
$S --config c2 --frames 1 > $O/plain_c2.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'zrange_kernel|setup_kernel|lookback_scan|scatter_kernel|raster_kernel' -s 10 -c 5 -o $R/c2_kernels -f $S --config c2 --frames 1 > $O/ncu_c2.log 2>&1
$NCU -k regex:'setup_kernel|scatter_kernel|raster_kernel' -s 6 -c 3 -o $R/c3_kernels -f $S --config c3 --frames 1 > $O/ncu_c3.log 2>&1
$NCU -k regex:'setup_kernel|scatter_kernel|raster_kernel' -s 6 -c 3 -o $R/floor_kernels -f $S --config c3 --scale 0.01 --frames 1 > $O/ncu_floor.log 2>&1
$NCU -k regex:'select_kernel|setup_kernel|scatter_kernel|raster_kernel' -s 8 -c 4 -o $R/c4band_kernels -f $S --config c4 --band 3/8 --frames 1 > $O/ncu_c4band.log 2>&1
for n in c2_kernels c3_kernels floor_kernels c4band_kernels; do
  python profiles/summarize.py $R/$n.ncu-rep $O/r02b_ncu_$n.txt > /dev/null 2>&1
  ls -la $R/$n.ncu-rep
done
python bench.py --steps 2 --warmup 1 --no-legs --no-cpu-baseline > $O/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-legs --no-cpu-baseline > $O/ncu_launch.log 2>&1
head -60 $O/r02b_ncu_c4band_kernels.txt | cut -c1-160
grep -A22 "raster_kernel" $O/r02b_ncu_floor_kernels.txt | head -30 | cut -c1-200
