#!/bin/bash
# final default tuning at the real bench size (1e9 queries per launch)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-queries 1000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/aa_$name.json 2> gpurun_out/aa_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/aa_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/aa_$name.err").read()[-1500:])
PY
}
run w4_cpl4
run w5_cpl4 --param walk_ctas_per_sm=5
run w6_cpl4 --param walk_ctas_per_sm=6
run w4_cpl2 --cells-per-leaf 2
run w5_cpl2 --param walk_ctas_per_sm=5 --cells-per-leaf 2
run w4_cpl1 --cells-per-leaf 1
run w5_cpl1 --param walk_ctas_per_sm=5 --cells-per-leaf 1
run w4_cpl4_again
