#!/bin/bash
# exact replay: entry nodes on/off x eager/lazy child fetch, on config 3 (20 % deferred) and in exact mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_w.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_w.log
BASE=$PWD/gsl_scattered_interpolation_b200/lib/libgsl_scattered_cuda.so
LAZY=$PWD/variants/lazy/libgsl_scattered_cuda.so
for v in "eager_entry:$BASE:exact_entry=1" "eager_root:$BASE:exact_entry=0" "lazy_entry:$LAZY:exact_entry=1" "lazy_root:$LAZY:exact_entry=0"; do
  IFS=: read name lib prm <<< "$v"
  GSI_LIB_PATH=$lib GSI_PARAMS="$prm" timeout 600 python scripts/bench_configs.py 3 2> gpurun_out/w_$name.err > gpurun_out/w_cfg3_$name.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/w_cfg3_$name.json").read().strip().split("\n")[-1])
    print("$name cfg3", "q/s %.3e"%d["queries_per_s"], "ms %.3f"%d["ms"], "locate %.3f exact %.3f"%(d["locate_ms"], d["exact_ms"]), d["fast_vs_exact_mismatches"], d["oracle_port_mismatches_on_2e5"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/w_$name.err").read()[-1000:])
PY
  GSI_LIB_PATH=$lib timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-queries 1000000 --queries 50000000 --param mode=1 --param $prm > gpurun_out/w_exact_$name.json 2>> gpurun_out/w_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/w_exact_$name.json"))
    print("$name exact-all 1e6 nodes", "q/s %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], d["verify"]["checksum_leaf"], d["verify"]["checksum_val"])
except Exception as e: print("$name exact failed", e); print(open("gpurun_out/w_$name.err").read()[-1000:])
PY
done
