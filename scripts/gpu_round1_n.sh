#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_n.log
for c in 4 2 3; do echo "=== config $c stage"; GSI_PARAMS="kernel=3" timeout 600 python scripts/bench_configs.py $c 2>gpurun_out/n$c.err | tee gpurun_out/n_cfg${c}_stage.json | cut -c1-800; tail -2 gpurun_out/n$c.err; done
echo "=== bench 2e8 stage"; timeout 600 python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000 --param kernel=3 > gpurun_out/n_bench_stage.json 2>gpurun_out/n_bench.err; cut -c1-400 gpurun_out/n_bench_stage.json; tail -2 gpurun_out/n_bench.err
echo "=== bench 2e8 walk"; timeout 600 python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000 --param kernel=0 > gpurun_out/n_bench_walk.json 2>gpurun_out/n_bench.err; cut -c1-400 gpurun_out/n_bench_walk.json
echo "=== bench 1e5 nodes stage"; timeout 600 python bench.py --nodes 100000 --queries 100000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --param kernel=3 > gpurun_out/n_bench_stage_1e5.json 2>gpurun_out/n_bench.err; cut -c1-400 gpurun_out/n_bench_stage_1e5.json
echo "=== bench 1e5 nodes walk"; timeout 600 python bench.py --nodes 100000 --queries 100000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --param kernel=0 > gpurun_out/n_bench_walk_1e5.json 2>gpurun_out/n_bench.err; cut -c1-400 gpurun_out/n_bench_walk_1e5.json
