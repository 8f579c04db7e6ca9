#!/bin/bash
# round 2, call W: does a larger L1 (124 KB) with all 32 warps resident help k_walkc?  256-thread CTAs, trimmed seed-base copy
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 1000"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2w_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2w_$name.err").read()[-1500:])
PY
}
run base $B
run t256_co72 env GSI_LIB_PATH=$PWD/variants/libgsi_t256.so $B
run t256s_co58 env GSI_LIB_PATH=$PWD/variants/libgsi_t256s.so $B --param walk_carveout=58
run t256s_co72 env GSI_LIB_PATH=$PWD/variants/libgsi_t256s.so $B --param walk_carveout=72
run t256s_co44 env GSI_LIB_PATH=$PWD/variants/libgsi_t256s.so $B --param walk_carveout=44
run t128s_co58 env GSI_LIB_PATH=$PWD/variants/libgsi_t128s.so $B --param walk_carveout=58
run t128s_co72 env GSI_LIB_PATH=$PWD/variants/libgsi_t128s.so $B --param walk_carveout=72
run t256s_co58_1e9 env GSI_LIB_PATH=$PWD/variants/libgsi_t256s.so $B --param walk_carveout=58 --queries 1000000000
run base_1e9 $B --queries 1000000000
