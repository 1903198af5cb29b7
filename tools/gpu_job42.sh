#!/bin/bash
mkdir -p gpurun_out/j42
{ echo "# GPU test suite (-m gpu) against the CHECKED build: every list index and shared-memory address the kernels compute is"
  echo "# asserted in range on the device (cpu_renderer_b200/csrc: B200R_ASSERT, -DB200R_CHECKED; tools/build_variant.sh checked)."
  echo "# compute-sanitizer answers on this pool: 'compute-sanitizer is closed on this pool and stays closed'."
  echo "# A failed device assert traps the kernel; the next API call returns B200R_E_CUDA and the test fails."
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_checked.so python -m pytest tests -m gpu -q -x 2>&1 | tail -6
  echo; echo "# scheduling independence: the same C3-like frame under 16 combinations of tile shape and park / refill thresholds"
  B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_checked.so python tools/raster_sweep.py --config c3 --scale 0.2 --tiles 128x8,64x16,128x16,256x8 --pend 1,4,16,32 --refill 8 --frames 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['tile'], 'pend', d['pend'], 'refill', d['refill'], 'same_image', d['same_image'])
    else:
        print(l.rstrip())
"
} > gpurun_out/j42/checked_build.txt 2>&1
cat gpurun_out/j42/checked_build.txt | tail -22
