#!/bin/bash
# end-of-round evidence at HEAD: GPU suite, smoke, default bench line, reference arm, the other configs / shading modes
O=gpurun_out/j38; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py > $O/r02_bench_default.json 2> $O/bench_default.err ) 2> $O/time_default.txt; tail -3 $O/time_default.txt
( time python bench.py --impl reference > $O/r02_bench_c2_reference.json 2> $O/bench_ref.err ) 2> $O/time_ref.txt; tail -3 $O/time_ref.txt
python bench.py --config c3 --steps 20 > $O/r02_bench_c3.json 2> $O/c3.err
python bench.py --config c1 --steps 20 > $O/r02_bench_c1.json 2> $O/c1.err
python bench.py --config c5 --steps 3 > $O/r02_bench_c5.json 2> $O/c5.err
for sc in 0.01 0.05 0.2; do python bench.py --config c3 --steps 20 --scale $sc --no-cpu-baseline > $O/r02_bench_c3_scale$sc.json 2> $O/c3s.err; done
python bench.py --steps 20 --phong --no-cpu-baseline > $O/r02_bench_c2_phong.json 2> $O/v.err
python bench.py --steps 20 --textured --no-cpu-baseline > $O/r02_bench_c2_textured.json 2>> $O/v.err
python bench.py --steps 20 --textured --phong --no-cpu-baseline > $O/r02_bench_c2_textured_phong.json 2>> $O/v.err
python bench.py --config c3 --steps 20 --phong --no-cpu-baseline > $O/r02_bench_c3_phong.json 2>> $O/v.err
python bench.py --config c3 --steps 20 --textured --no-cpu-baseline > $O/r02_bench_c3_textured.json 2>> $O/v.err
python bench.py --config c3 --steps 20 --textured --phong --no-cpu-baseline > $O/r02_bench_c3_textured_phong.json 2>> $O/v.err
ls -la $O/*.json | awk '{print $5, $9}'
