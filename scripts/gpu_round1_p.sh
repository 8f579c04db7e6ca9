#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/p_$name.json 2> gpurun_out/p_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/p_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"])
except Exception as e: print("$name failed", e)
PY
}
run w8_1e6 --queries 200000000 --param walk_ctas_per_sm=8
run w6_1e6 --queries 200000000 --param walk_ctas_per_sm=6
run w4_1e6 --queries 200000000 --param walk_ctas_per_sm=4
run w8_1e5 --nodes 100000 --queries 100000000 --param walk_ctas_per_sm=8
run w6_1e5 --nodes 100000 --queries 100000000 --param walk_ctas_per_sm=6
run w4_1e5 --nodes 100000 --queries 100000000 --param walk_ctas_per_sm=4
