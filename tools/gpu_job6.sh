#!/bin/bash
# sanitizer evidence + the two-device test; run with --gpus 2
mkdir -p gpurun_out/j6
python -m pytest tests/test_gpu_multi_device.py -q -m gpu > gpurun_out/j6/two_devices.log 2>&1; tail -3 gpurun_out/j6/two_devices.log
export CUDA_VISIBLE_DEVICES=0
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j6/memcheck_smoke.log 2>&1; tail -4 gpurun_out/j6/memcheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_frames.py 4000 > gpurun_out/j6/memcheck_frames.log 2>&1; tail -4 gpurun_out/j6/memcheck_frames.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python tools/sanitize_frames.py 1500 > gpurun_out/j6/racecheck_frames.log 2>&1; tail -6 gpurun_out/j6/racecheck_frames.log
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_frames.py 1500 > gpurun_out/j6/synccheck_frames.log 2>&1; tail -3 gpurun_out/j6/synccheck_frames.log
timeout 300 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_frames.py 1500 > gpurun_out/j6/initcheck_frames.log 2>&1; tail -3 gpurun_out/j6/initcheck_frames.log
