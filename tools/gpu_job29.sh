#!/bin/bash
O=gpurun_out/j29; mkdir -p $O
S="python tools/raster_sweep.py --pend 4 --refill 12 --frames 6"
: > $O/sweep.log
for b in 0/1 0/8 3/8 7/8 1/2; do echo "band $b" >> $O/sweep.log; $S --config c4 --tiles 0x0 --band $b >> $O/sweep.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c4_band3.csv python tools/raster_sweep.py --config c4 --tiles 0x0 --band 3/8 --frames 1 > $O/ncu.log 2>&1
python - <<'Q'
import json,csv,collections
for l in open("gpurun_out/j29/sweep.log"):
    if l.startswith("{"):
        d=json.loads(l); print("   setup",d["setup_kernel"],"scan",d["tile_scan_kernel"],"scatter",d["scatter_kernel"],"raster",d["raster_kernel"],"frame",d["frame"])
    else: print(l.rstrip()[:200])
rows=[r for r in csv.reader(open("gpurun_out/j29/launches_c4_band3.csv")) if len(r)>5]
hdr=None
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and r[hdr.index("Metric Name")]=="gpu__time_duration.sum":
        print(r[hdr.index("Kernel Name")][:60], r[hdr.index("Metric Value")], r[hdr.index("Metric Unit")])
Q
