#!/bin/bash
mkdir -p gpurun_out
echo "=== config 3 split (walk kernel)"; timeout 600 python scripts/bench_configs.py 3 2>gpurun_out/h3.err | tee gpurun_out/h_cfg3_walk.json; tail -2 gpurun_out/h3.err
echo "=== config 3 simple kernel"; GSI_PARAMS="kernel=1" timeout 600 python scripts/bench_configs.py 3 2>>gpurun_out/h3.err | tee gpurun_out/h_cfg3_simple.json
echo "=== config 4 walk"; timeout 600 python scripts/bench_configs.py 4 2>gpurun_out/h4.err | tee gpurun_out/h_cfg4_walk.json; tail -2 gpurun_out/h4.err
echo "=== config 4 simple"; GSI_PARAMS="kernel=1" timeout 600 python scripts/bench_configs.py 4 2>>gpurun_out/h4.err | tee gpurun_out/h_cfg4_simple.json
echo "=== config 4 simple cpl1"; GSI_PARAMS="kernel=1,cells_per_leaf=1" timeout 600 python scripts/bench_configs.py 4 2>>gpurun_out/h4.err | tee gpurun_out/h_cfg4_simple_cpl1.json
echo "=== config 2"; timeout 600 python scripts/bench_configs.py 2 2>gpurun_out/h2.err | tee gpurun_out/h_cfg2.json
