#!/bin/bash
# round 2, call F: tile kernel with coordinate prefetch + hoisted-reciprocal division, k_bin L2 prefetch
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout=200 -k "division or reorder_pipeline_on_the_golden or reorder_pipeline_edge or weather_golden or config1_golden" > gpurun_out/r2f_pytest_small.log 2>&1; echo "pytest small exit $?"; tail -4 gpurun_out/r2f_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2f_$name.err").read()[-1500:])
PY
}
run pipe --param reorder=1
run direct --param reorder=0 --param walk_ctas_per_sm=5
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
CMD="python scripts/profile_locate.py 1e6 134217728 1 2 reorder=1"
timeout 200 $CMD > gpurun_out/r2f_plain.log 2>&1 && timeout 400 ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm' -s 3 -c 3 --csv --log-file gpurun_out/r2f_pipe_metrics.csv $CMD > gpurun_out/r2f_ncu.log 2>&1
echo "ncu metrics exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_tile_walk|k_bin' -s 2 -c 2 -f -o gpurun_out/r2f_pipe_full $CMD > gpurun_out/r2f_ncu_full.log 2>&1; echo "ncu full exit $?"; ls -la gpurun_out/r2f_pipe_full.ncu-rep
run pipe_t512 --param reorder=1 --param tile_threads=512
run pipe_r1 --param reorder=1 --param tile_ranges=1
