#!/bin/bash
# round 2, call B: small reorder/derivative tests first (fail fast), A/B direct vs pipeline, per-kernel ncu, then the full suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r2b_box.txt 2>&1; nproc >> gpurun_out/r2b_box.txt
timeout 300 python -m pytest tests -m gpu -x -q --timeout=120 -k "reorder_pipeline_on_the_golden or reorder_pipeline_edge or deriv or weather_golden" > gpurun_out/r2b_pytest_small.log 2>&1; echo "pytest small exit $?"; tail -4 gpurun_out/r2b_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 200000000 --verify-queries 200000000"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2b_$name.err").read()[-1500:])
PY
}
run pipe --param reorder=1
run direct --param reorder=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
CMD="python scripts/profile_locate.py 1e6 134217728 1 2 reorder=1"
timeout 200 $CMD > gpurun_out/r2b_plain.log 2>&1 && timeout 400 ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm' -s 3 -c 3 --csv --log-file gpurun_out/r2b_pipe_metrics.csv $CMD > gpurun_out/r2b_ncu.log 2>&1
echo "ncu metrics exit $?"
run pipe_t512 --param reorder=1 --param tile_threads=512
run pipe_r1 --param reorder=1 --param tile_ranges=1
run pipe_r4 --param reorder=1 --param tile_ranges=4
run pipe_tl500 --param reorder=1 --param tile_leaves=500
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2b_pytest.log
