#!/bin/bash
# round 2, call U: shared-memory carve-out (L1 size) of k_walkc / k_walkf, k_walkf with 144-byte pitch; config-3 exact replay profile
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 1000"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2u_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2u_$name.err").read()[-1500:])
PY
}
for co in -1 100 86 72 58; do run walkc_co$co $B --param walk_carveout=$co; done
for co in -1 100 86 72 58; do run walkf_co$co $B --param walk_variant=3 --param walk_carveout=$co; done
run walkf_pitch144 env GSI_LIB_PATH=$PWD/variants/libgsi_pitch144.so $B --param walk_variant=3
run walkf_pitch144_co86 env GSI_LIB_PATH=$PWD/variants/libgsi_pitch144.so $B --param walk_variant=3 --param walk_carveout=86
timeout 300 python scripts/profile_cfg3.py 3 > gpurun_out/r2u_cfg3_plain.log 2>&1; tail -1 gpurun_out/r2u_cfg3_plain.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_exact_list4' -s 1 -c 1 -f -o gpurun_out/r2u_exact4_full python scripts/profile_cfg3.py 2 > gpurun_out/r2u_ncu_exact4.log 2>&1; echo "ncu exact4 exit $?"
timeout 300 python scripts/profile_cfg3.py 3 exact_coop=0 > gpurun_out/r2u_cfg3_plain_coop0.log 2>&1; tail -1 gpurun_out/r2u_cfg3_plain_coop0.log
