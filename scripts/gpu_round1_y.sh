#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "raster or golden or sixteen or kernel" > gpurun_out/pytest_gpu_y.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_y.log
for k in 4 1; do
  GSI_PARAMS="kernel=$k" timeout 600 python scripts/bench_configs.py 4 2> gpurun_out/y_$k.err > gpurun_out/y_cfg4_k$k.json; cat gpurun_out/y_cfg4_k$k.json | cut -c1-700; tail -2 gpurun_out/y_$k.err
done
