#!/bin/bash
# N-GPU bench: default line (C2 frames + c3 / c4_bands legs) with the fused gather.  usage: gpu_job20.sh N
N=${1:-2}; O=gpurun_out/j20; mkdir -p $O
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2> $O/bench_n$N.time
echo "bench rc=$?"; tail -3 $O/bench_n$N.time; tail -5 $O/bench_n$N.err
python - $N <<'Q'
import json,sys
n=sys.argv[1]
d=json.load(open(f"gpurun_out/j20/bench_n{n}.json"))
def show(k,l):
    print(k, "ms", round(l["ms_per_step"],4), "value", round(l["value"],1), "image_ok", l.get("image_ok"), "gather", {a:(round(b,4) if isinstance(b,float) else b) for a,b in (l.get("with_gather") or {}).items() if a!="what"},
          "nccl", {a:(round(b,4) if isinstance(b,float) else b) for a,b in (l.get("with_gather_nccl") or {}).items() if a!="what"}, "e2e", l.get("e2e",{}).get("ms_per_step"))
show("main",d)
for k,l in d.get("legs",{}).items(): show(k,l)
Q
