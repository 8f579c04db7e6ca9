#!/bin/bash
# round 2, call J: k_walkp (two walks per lane in rotation behind L2 prefetches, continuous dynamically dealt queue) vs k_walk
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "not stated and not million and not two_devices" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2j_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2j_$name.err").read()[-1500:])
PY
}
run walkp
run walk_old --param walk_variant=0
run walkp_static --param walk_dynamic=0
run walkp_chunk512 --param walk_chunk=512
run walkp_chunk8192 --param walk_chunk=8192
run walkp_w4 --param walk_ctas_per_sm=4
run walkp_s128 --param leaf_stride=128
run walkp_cpl8 --cells-per-leaf 8
run walkp_cpl2 --cells-per-leaf 2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
for v in "walk_variant=1" "walk_variant=1 leaf_stride=128"; do
  tag=$(echo $v | tr ' =' '__')
  CMD="python scripts/profile_locate.py 1e6 134217728 1 2 $v"
  $CMD > gpurun_out/r2j_plain_$tag.log 2>&1 && \
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2j_walk_$tag.csv $CMD > gpurun_out/r2j_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"
done
# default bench, full size
timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["verify"]["cpu_reference_mismatches_on_strided_subsample"])
for k,v in d.get("configs",{}).items(): print(k, {a:v.get(a) for a in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches")})
PY
