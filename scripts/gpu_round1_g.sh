#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/g_$name.json 2> gpurun_out/g_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d.get("walk_stats"), d["verify"]["fast_vs_exact_mismatches"])
except Exception as e: print("$name failed", e)
PY
}
run s128_cpl4 --param leaf_stride=128 --param walk_ctas_per_sm=4
run s128_cpl2 --param leaf_stride=128 --param walk_ctas_per_sm=4 --cells-per-leaf 2
run s128_cpl1 --param leaf_stride=128 --param walk_ctas_per_sm=4 --cells-per-leaf 1
run s128_cpl1_w5 --param leaf_stride=128 --param walk_ctas_per_sm=5 --cells-per-leaf 1
run s96_cpl2_w5 --param leaf_stride=96 --param walk_ctas_per_sm=5 --cells-per-leaf 2
echo "=== configs"
timeout 900 python scripts/bench_configs.py 2 3 4 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs exit $?"; cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
