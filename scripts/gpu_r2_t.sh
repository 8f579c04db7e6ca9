#!/bin/bash
# round 2, call T: escape-free compact seeds (base = centre seed - 32768, strays take a neighbour's seed), k_walkc with the base table in shared memory; k_walkf source-level profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "fast_path_kernel or weather or tiny_deferred or nonfinite or config3_mixed or million_nodes_uniform or save_load or reorder or config1" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2t_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2t_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], "grid", d["setup"]["grid"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2t_$name.err").read()[-1500:])
PY
}
run walkc $B
run walkf $B --param walk_variant=3
run walkc_cpl6 $B --cells-per-leaf 6
run walkc_1e9 $B --queries 1000000000 --verify-queries 1000
run walkf_1e9 $B --queries 1000000000 --verify-queries 1000 --param walk_variant=3
for v in "walk_variant=2" "walk_variant=3"; do python scripts/profile_locate.py 1e5 50000000 1 5 $v | tail -1; done
CMD="python scripts/profile_locate.py 1e6 134217728 1 2 walk_variant=3"
$CMD > gpurun_out/r2t_plain_walkf.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_walkf' -s 1 -c 1 -f -o gpurun_out/r2t_walkf_full $CMD > gpurun_out/r2t_ncu_walkf.log 2>&1; echo "ncu walkf exit $?"; tail -1 gpurun_out/r2t_plain_walkf.log
CMD="python scripts/profile_locate.py 1e6 134217728 1 2 walk_variant=2"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_walkc' -s 1 -c 1 -f -o gpurun_out/r2t_walkc_full $CMD > gpurun_out/r2t_ncu_walkc.log 2>&1; echo "ncu walkc exit $?"
