#!/bin/bash
mkdir -p gpurun_out/j13
python -m pytest tests -m gpu -x -q > gpurun_out/j13/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j13/pytest.log
tail -3 gpurun_out/j13/pytest.log
B200R_LIB=$PWD/cpu_renderer_b200/libb200raster_checked.so python -m pytest tests -m gpu -x -q > gpurun_out/j13/pytest_checked.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j13/pytest_checked.log
tail -3 gpurun_out/j13/pytest_checked.log
