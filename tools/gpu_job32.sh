#!/bin/bash
# device-side edge table tests; where the raster kernel's time goes on a nearly empty 4K frame
O=gpurun_out/j32; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_cpp.py -m gpu -x -q -k "edge_table or whole or level0 or host or object" > $O/pytest_edge.log 2>&1; tail -4 $O/pytest_edge.log
for sc in 0.01 0.05 1.0; do
  echo "== clocks scale $sc" >> $O/stats.log
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_clk.so python tools/raster_stats.py --config c3 --scale $sc --tile 128x8 >> $O/stats.log 2>&1
  echo "== counters scale $sc" >> $O/stats.log
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_stats.so python tools/raster_stats.py --config c3 --scale $sc --tile 128x8 >> $O/stats.log 2>&1
done
cat $O/stats.log | cut -c1-900
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_floor.csv python tools/raster_sweep.py --config c3 --scale 0.01 --tiles 0x0 --frames 2 > $O/ncu.log 2>&1
python - <<'Q'
import csv
rows=[r for r in csv.reader(open("gpurun_out/j32/launches_floor.csv")) if len(r)>5]
hdr=None; n=0
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and r[hdr.index("Metric Name")]=="gpu__time_duration.sum" and "b200r" in r[hdr.index("Kernel Name")]:
        n+=1
        if n>24 and n<=40: print(r[hdr.index("Kernel Name")][:60], r[hdr.index("Metric Value")], r[hdr.index("Metric Unit")])
Q
