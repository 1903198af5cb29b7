#!/bin/bash
# default bench line on N GPUs (arg 1) the way the driver launches it
N=${1:-2}
O=gpurun_out/j39; mkdir -p $O
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_default_n$N.json 2> $O/bench_n$N.err ) 2> $O/time_n$N.txt
tail -3 $O/time_n$N.txt; tail -5 $O/bench_n$N.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 2 --warmup 3 --config c5 > $O/r02_bench_c5_n$N.json 2> $O/c5_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --impl reference > $O/r02_bench_reference_n$N.json 2> $O/ref_n$N.err
ls -la $O | awk '{print $5, $9}'
