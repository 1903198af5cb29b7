#!/bin/bash
# final check of the committed tree: GPU suite, smoke, a short default bench line
O=gpurun_out/j44; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log | cut -c1-200
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['image_ok'], d['e2e']['value'], {k:(v['ms_per_step'], v['image_ok']) for k,v in d['legs'].items()})"
