#!/bin/bash
# k_walk chunk size (queries per warp-chunk) at the bench size: the streams' TLB footprint scales with it
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-queries 1000000 --verify-queries 0"
for c in 8192 4096 2048 1024 512 256; do
  timeout 600 $B --param walk_chunk=$c > gpurun_out/chunk_$c.json 2> gpurun_out/chunk_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/chunk_$c.json")); print("chunk $c", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"]["leaf"])
except Exception as e: print("failed", e); print(open("gpurun_out/chunk_$c.err").read()[-800:])
PY
done
