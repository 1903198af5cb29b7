#!/bin/bash
mkdir -p gpurun_out/j5
N=${1:-2}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/j5/bench_n$N.json 2> gpurun_out/j5/bench_n$N.err ) 2> gpurun_out/j5/bench_n$N.time
echo "rc=$?"; tail -3 gpurun_out/j5/bench_n$N.time; tail -5 gpurun_out/j5/bench_n$N.err
python - $N <<'P'
import json,sys
N=sys.argv[1]
d=json.load(open(f"gpurun_out/j5/bench_n{N}.json"))
print(d["n_gpus"], d["ms_per_step"], d["ms_per_step_passes"], d["value"], d.get("image_ok"), d["e2e"]["ms_per_step"], d.get("with_gather"))
for k,l in d.get("legs",{}).items():
    print("  leg",k,l["ms_per_step"],l["value"],l["image_fnv"],l["image_ok"],l["roofline"]["frac"],l["stage_ms"], l.get("e2e"), l.get("with_gather"))
P
