#!/bin/bash
# round 2, call A: GPU suite at HEAD (new contexts + pipeline v2), then direct vs pipeline A/B and per-kernel ncu metrics
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r2a_box.txt 2>&1
nproc >> gpurun_out/r2a_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2a_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000 --verify-queries 200000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/r2a_$name.json 2> gpurun_out/r2a_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2a_$name.err").read()[-1500:])
PY
}
run direct --param reorder=0
run pipe --param reorder=1
run pipe_t512 --param reorder=1 --param tile_threads=512
run pipe_r1 --param reorder=1 --param tile_ranges=1
run pipe_r4 --param reorder=1 --param tile_ranges=4
run pipe_tl500 --param reorder=1 --param tile_leaves=500
run pipe_tl2000 --param reorder=1 --param tile_leaves=1900
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
CMD="python scripts/profile_locate.py 1e6 134217728 1 2 reorder=1"
$CMD > gpurun_out/r2a_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm' -s 3 -c 3 --csv --log-file gpurun_out/r2a_pipe_metrics.csv $CMD > gpurun_out/r2a_ncu.log 2>&1
echo "ncu metrics exit $?"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_walk' -s 1 -c 1 -o gpurun_out/r2a_tile_walk $CMD > gpurun_out/r2a_ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
