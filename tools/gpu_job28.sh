#!/bin/bash
N=${1:-2}; O=gpurun_out/j28; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$R bench.py --gpus $N --config c4 --steps 10 --warmup 3 --no-legs > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err; echo "c4 rc=$?"
python - $N <<'Q'
import json,sys
n=sys.argv[1]
l=json.load(open(f"gpurun_out/j28/bench_c4_n{n}.json"))
print("ms", round(l["ms_per_step"],4), "value", round(l["value"],1), "image_ok", l.get("image_ok"), "stage", {a:round(b,3) for a,b in l["stage_ms"].items()},
      "\n gather", {a:(round(b,4) if isinstance(b,float) else b) for a,b in (l.get("with_gather") or {}).items() if a!="what"},
      "\n nccl", {a:(round(b,4) if isinstance(b,float) else b) for a,b in (l.get("with_gather_nccl") or {}).items() if a!="what"}, "e2e", l.get("e2e",{}).get("ms_per_step"))
Q
