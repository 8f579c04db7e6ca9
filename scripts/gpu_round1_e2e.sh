#!/bin/bash
# host-buffer (e2e) leg vs pipeline stage size
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --queries 100000000 --e2e-queries 100000000 --verify-queries 0"
for hc in 1048576 4194304 16777216 50000000; do
  timeout 600 $B --param host_chunk=$hc > gpurun_out/e2e_$hc.json 2> gpurun_out/e2e_$hc.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e2e_$hc.json")); print("host_chunk $hc", "e2e %.4e q/s"%d["e2e"]["value"], "(%.1f GB/s in, %.1f GB/s out)"%(d["e2e"]["value"]*16/1e9, d["e2e"]["value"]*12/1e9))
except Exception as e: print("failed", e); print(open("gpurun_out/e2e_$hc.err").read()[-800:])
PY
done
