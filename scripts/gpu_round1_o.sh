#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "raster or sixteen or kernel" > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_o.log
echo "=== config 4"; timeout 600 python scripts/bench_configs.py 4 2>gpurun_out/o4.err | tee gpurun_out/o_cfg4.json | cut -c1-500; tail -2 gpurun_out/o4.err
