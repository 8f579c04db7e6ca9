#!/bin/bash
# shared-memory-staged tile kernel: parity, then A/B against the L1 tile kernel and the direct walk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "reorder or slab or two_devices or stress" > gpurun_out/sm_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/sm_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000 --verify-queries 200000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/sm_$name.json 2> gpurun_out/sm_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sm_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/sm_$name.err").read()[-1500:])
PY
}
run direct --param reorder=0
run smem --param reorder=1 --param tile_smem=1
run l1 --param reorder=1 --param tile_smem=0
run smem_r8 --param reorder=1 --param tile_smem=1 --param tile_ranges=8
run smem_t500 --param reorder=1 --param tile_smem=1 --param tile_leaves=500
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
CMD="python scripts/profile_locate.py 1e6 1e8 1 2 reorder=1"
$CMD > gpurun_out/sm_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm' -s 3 -c 3 --csv --log-file gpurun_out/sm_launches.csv $CMD > gpurun_out/sm_ncu.log 2>&1
echo "ncu exit $?"
