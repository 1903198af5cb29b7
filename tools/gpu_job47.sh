#!/bin/bash
O=gpurun_out/j47; mkdir -p $O
timeout 70 python bench.py --phong --no-cpu-baseline --steps 5 --warmup 3 > $O/r02_bench_c2_phong.json 2> $O/a.err
python -c "
import json; d=json.loads(open('$O/r02_bench_c2_phong.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('shaded_vs_oracle'))"
timeout 45 python bench.py --config c3 --textured --phong --no-cpu-baseline --steps 5 --warmup 3 > $O/r02_bench_c3_textured_phong.json 2> $O/b.err
python -c "
import json; d=json.loads(open('$O/r02_bench_c3_textured_phong.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('shaded_vs_oracle'))"
