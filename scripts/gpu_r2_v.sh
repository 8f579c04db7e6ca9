#!/bin/bash
# round 2, call V: k_walkc with 112-byte pitch + explicit 72 % carve-out (8 CTAs, 92 KB L1); L1::no_allocate variant;
# deferred list that grows to what the last launch needed + four-lane overflow replay (config 3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "fast_path_kernel or weather or tiny_deferred or nonfinite or config3 or million_nodes_uniform or reorder_pipeline_edge or config1 or reentrant" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2v_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 1000"
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2v_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2v_$name.err").read()[-1500:])
PY
}
run walkc $B
for co in -1 58 86; do run walkc_co$co $B --param walk_carveout=$co; done
run walkc_w6 $B --param walk_ctas_per_sm=6
run walkc_w7 $B --param walk_ctas_per_sm=7
for co in 100 86 72; do run noalloc_co$co env GSI_LIB_PATH=$PWD/variants/libgsi_noalloc.so $B --param walk_carveout=$co; done
run walkc_1e9 $B --queries 1000000000
run walkf_1e9 $B --queries 1000000000 --param walk_variant=3
for v in "walk_variant=2" "walk_variant=3"; do python scripts/profile_locate.py 1e5 50000000 1 5 $v | tail -1; done
timeout 300 python scripts/profile_cfg3.py 3 > gpurun_out/r2v_cfg3_plain.log 2>&1; tail -1 gpurun_out/r2v_cfg3_plain.log
( time timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err ) 2> gpurun_out/r2v_bench.time; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["clocks"]["sm_mhz"])
for k,v in d.get("configs",{}).items(): print(k, {a:v.get(a) for a in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches")} if isinstance(v, dict) else v)
PY
