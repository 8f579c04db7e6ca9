#!/bin/bash
# reorder pipeline v2 (k_bin2 + k_tile_locate2): parity tests, A/B, per-kernel ncu metrics
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "reorder" > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_v.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["gpu_launches"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/v_$name.err").read()[-1500:])
PY
}
run direct --param reorder=0
run v2 --param reorder=1
run v2_bin0 --param reorder=1 --param bin_kernel=0
run v2_tile0 --param reorder=1 --param tile_kernel=0
run v2_s7 --param reorder=1 --param tile_shift=7
run v2_r8 --param reorder=1 --param tile_ranges=8
run v2_r2 --param reorder=1 --param tile_ranges=2
run v2_1e5 --param reorder=1 --nodes 100000
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
CMD="python scripts/profile_locate.py 1e6 1e8 1 2 reorder=1"
$CMD > gpurun_out/v_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm|k_walk|k_exact_list' --csv --log-file gpurun_out/v_reorder_launches.csv $CMD > gpurun_out/v_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/v_plain.log
