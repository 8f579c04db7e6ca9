#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_m.log
echo "=== config 4"; timeout 600 python scripts/bench_configs.py 4 2>gpurun_out/m4.err | tee gpurun_out/m_cfg4.json | cut -c1-700; tail -2 gpurun_out/m4.err
echo "=== config 4 cpl1"; GSI_PARAMS="cells_per_leaf=1" timeout 600 python scripts/bench_configs.py 4 2>>gpurun_out/m4.err | tee gpurun_out/m_cfg4_cpl1.json | cut -c1-700
echo "=== config 4 cpl16"; GSI_PARAMS="cells_per_leaf=16" timeout 600 python scripts/bench_configs.py 4 2>>gpurun_out/m4.err | tee gpurun_out/m_cfg4_cpl16.json | cut -c1-700
