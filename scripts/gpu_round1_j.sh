#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_j.log
for c in 2 3 4; do echo "=== config $c"; timeout 600 python scripts/bench_configs.py $c 2>gpurun_out/j$c.err | tee gpurun_out/j_cfg$c.json | cut -c1-900; tail -2 gpurun_out/j$c.err; done
echo "=== config 4 walk kernel"; GSI_PARAMS="kernel=0" timeout 600 python scripts/bench_configs.py 4 2>>gpurun_out/j4.err | tee gpurun_out/j_cfg4_walk.json | cut -c1-600
echo "=== config 4 simple cpl1"; GSI_PARAMS="cells_per_leaf=1" timeout 600 python scripts/bench_configs.py 4 2>>gpurun_out/j4.err | tee gpurun_out/j_cfg4_cpl1.json | cut -c1-600
echo "=== config 3 simple kernel"; GSI_PARAMS="kernel=1" timeout 600 python scripts/bench_configs.py 3 2>>gpurun_out/j3.err | tee gpurun_out/j_cfg3_simple.json | cut -c1-900
echo "=== bench 2e8"; timeout 600 python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000 > gpurun_out/j_bench.json 2>gpurun_out/j_bench.err; cut -c1-1500 gpurun_out/j_bench.json
