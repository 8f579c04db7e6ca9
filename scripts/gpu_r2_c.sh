#!/bin/bash
# round 2, call C: compact seed table (slab v5): direct-walk A/B (cells per leaf, leaf stride) with DRAM bytes, new tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "golden or reentrant or pageable or config3_1e5 or config2_1e5 or save_load" > gpurun_out/r2c_pytest_small.log 2>&1; echo "pytest small exit $?"; tail -4 gpurun_out/r2c_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 200000000 --verify-queries 200000000"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2c_$name.err").read()[-1500:])
PY
}
run direct --param reorder=0
run direct_cpl2 --param reorder=0 --cells-per-leaf 2
run direct_cpl1 --param reorder=0 --cells-per-leaf 1
run direct_cpl8 --param reorder=0 --cells-per-leaf 8
run direct_s128 --param reorder=0 --param leaf_stride=128
run direct_s128_cpl2 --param reorder=0 --param leaf_stride=128 --cells-per-leaf 2
run direct_w5 --param reorder=0 --param walk_ctas_per_sm=5
run direct_w6 --param reorder=0 --param walk_ctas_per_sm=6
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "reorder=0" "reorder=0 leaf_stride=128"; do
  tag=$(echo $v | tr ' =' '__')
  CMD="python scripts/profile_locate.py 1e6 100000000 1 2 $v"
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2c_walk_$tag.csv $CMD > gpurun_out/r2c_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"
done
