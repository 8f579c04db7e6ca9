#!/bin/bash
# round 2, call Q: A/B builds of the library (streaming cache hints on the query/result arrays; cp.async.ca) at 2^28 and 10^9 queries
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 1000"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2q_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2q_$name.err").read()[-1500:])
PY
}
run base $B
run hints env GSI_LIB_PATH=$PWD/variants/libgsi_streamhints.so $B
run cpca env GSI_LIB_PATH=$PWD/variants/libgsi_cpca.so $B
run base_1e9 $B --queries 1000000000
run hints_1e9 env GSI_LIB_PATH=$PWD/variants/libgsi_streamhints.so $B --queries 1000000000
run base_w7_1e9 $B --queries 1000000000 --param walk_ctas_per_sm=7
run base_w6_1e9 $B --queries 1000000000 --param walk_ctas_per_sm=6
