#!/bin/bash
# reorder pipeline: parity tests, then A/B against the direct walk at 10^6 nodes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "reorder" > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_q.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["gpu_launches"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/q_$name.err").read()[-1500:])
PY
}
run direct --param reorder=0
run reorder --param reorder=1
run reorder_s5 --param reorder=1 --param tile_shift=5
run reorder_s7 --param reorder=1 --param tile_shift=7
run reorder_r8 --param reorder=1 --param tile_ranges=8
run reorder_r1 --param reorder=1 --param tile_ranges=1
run reorder_c4k --param reorder=1 --param reorder_chunk=4096
run direct_1e5 --param reorder=0 --nodes 100000
run reorder_1e5 --param reorder=1 --nodes 100000
