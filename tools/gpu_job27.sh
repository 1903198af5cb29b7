#!/bin/bash
# 8-GPU evidence: default line (C2 frames + c4_bands leg, fused gather beside NCCL gather) and C5
N=${1:-8}; O=gpurun_out/j27; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time $R bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_default_n$N.json 2> $O/bench_default_n$N.err ) 2> $O/bench_default_n$N.time
echo "default rc=$?"; tail -3 $O/bench_default_n$N.time
( time $R bench.py --gpus $N --config c5 --steps 3 --warmup 3 --no-legs > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err ) 2> $O/bench_c5_n$N.time
echo "c5 rc=$?"; tail -3 $O/bench_c5_n$N.time
python - $N <<'Q'
import json,sys
n=sys.argv[1]
def show(k,l):
    print(k, "ms", round(l["ms_per_step"],4), "value", round(l["value"],1), "image_ok", l.get("image_ok"), "stage", {a:round(b,3) for a,b in l["stage_ms"].items()},
          "\n    gather", {a:(round(b,4) if isinstance(b,float) else b) for a,b in (l.get("with_gather") or {}).items() if a!="what"},
          "\n    nccl", {a:(round(b,4) if isinstance(b,float) else b) for a,b in (l.get("with_gather_nccl") or {}).items() if a!="what"}, "e2e", l.get("e2e",{}).get("ms_per_step"), l.get("e2e",{}).get("value"))
for f in ("default","c5"):
    try: d=json.load(open(f"gpurun_out/j27/bench_{f}_n{n}.json"))
    except Exception as e: print(f, "unreadable", e); continue
    show(f,d)
    for k,l in d.get("legs",{}).items(): show("  leg "+k,l)
Q
