#!/bin/bash
# round 2, call S: slab v6 (coarse seed blocks, base table <= 4 KB) + k_walkf (fresh batches + ring of walkers, seed bases in shared memory)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "fast_path_kernel or weather or tiny_deferred or nonfinite or config3_mixed or million_nodes_uniform or save_load or reorder or seed" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2s_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], "grid", d["setup"]["grid"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2s_$name.err").read()[-1500:])
PY
}
run walkc $B
run walkf $B --param walk_variant=3
run walkf_stage1 env GSI_LIB_PATH=$PWD/variants/libgsi_stage1.so $B --param walk_variant=3
run walkf_stage2 env GSI_LIB_PATH=$PWD/variants/libgsi_stage2.so $B --param walk_variant=3
run walkf_chunk512 $B --param walk_variant=3 --param walk_chunk=512
run walkf_w7 $B --param walk_variant=3 --param walk_ctas_per_sm=7
run walkf_w6 $B --param walk_variant=3 --param walk_ctas_per_sm=6
run walkc_1e9 $B --queries 1000000000 --verify-queries 1000
run walkf_1e9 $B --queries 1000000000 --verify-queries 1000 --param walk_variant=3
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_requests.sum,lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread
prof() { tag=$1; shift
  CMD="python scripts/profile_locate.py 1e6 134217728 1 2 $*"
  $CMD > gpurun_out/r2s_plain_$tag.log 2>&1 && \
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2s_metrics_$tag.csv $CMD > gpurun_out/r2s_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"; tail -1 gpurun_out/r2s_plain_$tag.log
}
prof walkc walk_variant=2
prof walkf walk_variant=3
GSI_LIB_PATH=$PWD/variants/libgsi_stage1.so prof walkf_stage1 walk_variant=3
GSI_LIB_PATH=$PWD/variants/libgsi_stage2.so prof walkf_stage2 walk_variant=3
for v in "walk_variant=2" "walk_variant=3"; do python scripts/profile_locate.py 1e5 50000000 1 5 $v | tail -1; done
