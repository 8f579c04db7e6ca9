#!/bin/bash
# round 2, call K: which way of requesting the next leaf line ahead of time works (k_walkp, knob walk_pf)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2k_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2k_$name.err").read()[-1500:])
PY
}
run pf0 --param walk_pf=0
run pf2 --param walk_pf=2
run pf3 --param walk_pf=3
run pf4 --param walk_pf=4
run pf4_s128 --param walk_pf=4 --param leaf_stride=128
run pf2_s128 --param walk_pf=2 --param leaf_stride=128
run old --param walk_variant=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
for v in "walk_pf=4" "walk_pf=2" "walk_pf=3" "walk_variant=0"; do
  tag=$(echo $v | tr ' =' '__')
  CMD="python scripts/profile_locate.py 1e6 134217728 1 2 $v"
  $CMD > gpurun_out/r2k_plain_$tag.log 2>&1 && \
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2k_walk_$tag.csv $CMD > gpurun_out/r2k_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"; cat gpurun_out/r2k_plain_$tag.log | tail -1
done
