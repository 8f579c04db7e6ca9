#!/bin/bash
# round 2, call L: k_walkc (warp-cooperative cp.async staging of 128-byte leaves, hoisted reciprocals, 8 CTAs/SM) vs k_walk;
# device-located build vs host build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "fast_path_kernel or builder_with_device or walk_trace or save_load or slab_replication or tiny_deferred or nonfinite" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2l_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2l_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], "build_s", d["setup"]["host_build_seconds"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2l_$name.err").read()[-1500:])
PY
}
run walkc --param walk_variant=2 --param leaf_stride=128
run walkc_w6 --param walk_variant=2 --param leaf_stride=128 --param walk_ctas_per_sm=6
run walkc_w7 --param walk_variant=2 --param leaf_stride=128 --param walk_ctas_per_sm=7
run walkc_static --param walk_variant=2 --param leaf_stride=128 --param walk_dynamic=0
run walkc_chunk512 --param walk_variant=2 --param leaf_stride=128 --param walk_chunk=512
run old_hostbuild --param build_device=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ldgsts.sum
for v in "walk_variant=2 leaf_stride=128"; do
  tag=$(echo $v | tr ' =' '__')
  CMD="python scripts/profile_locate.py 1e6 134217728 1 2 $v"
  $CMD > gpurun_out/r2l_plain_$tag.log 2>&1 && \
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2l_walk_$tag.csv $CMD > gpurun_out/r2l_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"; cat gpurun_out/r2l_plain_$tag.log | tail -1
done
# config 2 shape (1e5 nodes, L2 resident)
for v in "walk_variant=2 leaf_stride=128" "walk_variant=0"; do python scripts/profile_locate.py 1e5 50000000 1 5 $v | tail -1; done
