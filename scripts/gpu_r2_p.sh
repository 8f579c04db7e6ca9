#!/bin/bash
# round 2, call P: k_walkc with the seed loads consumed a refill later; k_walkc2 (two walk sets per lane, double-buffered staging)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "fast_path_kernel or walk_trace or tiny_deferred or nonfinite or config3_mixed or million_nodes_uniform" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2p_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2p_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2p_$name.err").read()[-1500:])
PY
}
run walkc
run walkc2 --param walk_variant=3
run walkc2_w6 --param walk_variant=3 --param walk_ctas_per_sm=6
run walkc2_chunk512 --param walk_variant=3 --param walk_chunk=512
run walkc_cpl2 --cells-per-leaf 2
run walkc_cpl8 --cells-per-leaf 8
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
for v in "walk_variant=2" "walk_variant=3"; do
  tag=$(echo $v | tr ' =' '__')
  CMD="python scripts/profile_locate.py 1e6 134217728 1 2 $v"
  $CMD > gpurun_out/r2p_plain_$tag.log 2>&1 && \
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2p_walk_$tag.csv $CMD > gpurun_out/r2p_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"; cat gpurun_out/r2p_plain_$tag.log | tail -1
done
for v in "walk_variant=2" "walk_variant=3" "walk_variant=0"; do python scripts/profile_locate.py 1e5 50000000 1 5 $v | tail -1; done
