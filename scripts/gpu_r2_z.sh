#!/bin/bash
# round 2, call Z: cp.async.ca (leaf lines allocate in L1) against hot leaves (config 3: cage leaves, sparse regions)
mkdir -p gpurun_out
echo "== cg (default)"; timeout 300 python scripts/profile_cfg3.py 3 2>&1 | tail -1
echo "== ca"; GSI_LIB_PATH=$PWD/variants/libgsi_ca.so timeout 300 python scripts/profile_cfg3_classes.py 2>&1 | head -3
GSI_LIB_PATH=$PWD/variants/libgsi_ca.so timeout 300 python scripts/profile_cfg3.py 3 2>&1 | tail -1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 1000"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2z_$name.json 2> gpurun_out/r2z_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2z_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2z_$name.err").read()[-1500:])
PY
}
run cg $B
run ca env GSI_LIB_PATH=$PWD/variants/libgsi_ca.so $B
for lib in "" variants/libgsi_ca.so; do GSI_LIB_PATH=${lib:+$PWD/$lib} python scripts/profile_locate.py 1e5 50000000 1 5 | tail -1; done
