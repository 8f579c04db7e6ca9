#!/bin/bash
# L2 fetch granularity sweep + ncu capture of k_walk
mkdir -p gpurun_out
B="python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/d_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d.get("walk_stats"), d["verify"]["fast_vs_exact_mismatches"])
except Exception as e: print("$name failed", e)
PY
}
run g32_cpl4 --param l2_fetch=32 --param walk_ctas_per_sm=4
run g64_cpl4 --param l2_fetch=64 --param walk_ctas_per_sm=4
run g128_cpl4 --param l2_fetch=128 --param walk_ctas_per_sm=4
run g32_cpl1 --param l2_fetch=32 --param walk_ctas_per_sm=4 --cells-per-leaf 1
run g32_cpl2 --param l2_fetch=32 --param walk_ctas_per_sm=4 --cells-per-leaf 2
run g32_cpl1_w5 --param l2_fetch=32 --param walk_ctas_per_sm=5 --cells-per-leaf 1
run g32_cpl1_w3 --param l2_fetch=32 --param walk_ctas_per_sm=3 --cells-per-leaf 1
run g32_cpl05 --param l2_fetch=32 --param walk_ctas_per_sm=4 --cells-per-leaf 0.5
CMD2="python scripts/profile_locate.py 1e6 5e7 1 3"
$CMD2 > gpurun_out/plain_prof_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -o gpurun_out/prof_walk_r1b $CMD2 > gpurun_out/ncu_full_d.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_d.log
