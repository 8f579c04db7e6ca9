#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or slab or two_devices or reorder_pipeline_edge" > gpurun_out/vmm_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/vmm_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000 --verify-queries 0"
for rep in 1 2 3; do for v in 1 0; do
  timeout 600 $B --param vmm=$v > gpurun_out/vmm_${v}_$rep.json 2> gpurun_out/vmm_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/vmm_${v}_$rep.json")); print("vmm=$v rep $rep", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], d["verify"]["fast_vs_exact_mismatches"])
except Exception as e: print("failed", e); print(open("gpurun_out/vmm_${v}_$rep.err").read()[-800:])
PY
done; done
