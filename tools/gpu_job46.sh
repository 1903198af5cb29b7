#!/bin/bash
O=gpurun_out/j46; mkdir -p $O
python bench.py > $O/r02_bench_default.json 2> $O/bench.err; python -c "
import json; d=json.loads(open('$O/r02_bench_default.json').read().strip().splitlines()[-1])
cb=d['cpu_baseline']; print(d['ms_per_step'], d['value'], d['image_ok'], d['e2e']['value'], cb['value'], cb.get('scalar_1t'), {k:(v['ms_per_step'], v['image_ok']) for k,v in d['legs'].items()})"
