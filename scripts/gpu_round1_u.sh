#!/bin/bash
# full GPU suite + default bench (both arms) + ncu launch list + ncu --set full of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/u_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_u.log
timeout 900 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/u_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/u_bench_ref.json 2> gpurun_out/u_bench_ref.err; echo "ref exit $?"; cut -c1-800 gpurun_out/u_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000"
$CMD > gpurun_out/u_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/u_launches.csv $CMD > gpurun_out/u_ncu_launches.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/u_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_walk -s 3 -c 1 -o gpurun_out/prof_k_walk_r1c $CMD > gpurun_out/u_ncu_full.log 2>&1
echo "ncu full exit $?"
