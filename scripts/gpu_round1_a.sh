#!/bin/bash
# First GPU session: parity tests, smoke, a first bench line, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "=== pytest gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
echo "=== bench small"
timeout 900 python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
for cpl in 1 16; do
  timeout 600 python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --cells-per-leaf $cpl --e2e-queries 1000000 > gpurun_out/bench_cpl$cpl.json 2> gpurun_out/bench_cpl$cpl.err; echo "cpl $cpl exit $?"; cat gpurun_out/bench_cpl$cpl.json
done
echo "=== config 2 (1e5 nodes, 1e7 queries)"
timeout 600 python bench.py --nodes 100000 --queries 10000000 --steps 5 --warmup 3 --stats --no-cpu-baseline --e2e-queries 10000000 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"; cat gpurun_out/bench_cfg2.json
echo "=== full default bench"
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full exit $?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
