#!/bin/bash
# A/B: walk kernel (lanes decoupled, 96-B leaves, 256-bit loads) vs simple kernel; occupancy and grid sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_c.log
B="python bench.py --queries 200000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000"
run() { name=$1; shift; timeout 600 $B "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d.get("walk_stats"), d["verify"]["fast_vs_exact_mismatches"], d["clocks"])
except Exception as e: print("$name failed", e)
PY
}
run walk5 --param kernel=0
run simple --param kernel=1
run walk4 --param kernel=0 --param walk_ctas_per_sm=4
run walk3 --param kernel=0 --param walk_ctas_per_sm=3
run walk5_cpl1 --param kernel=0 --cells-per-leaf 1
run walk5_cpl16 --param kernel=0 --cells-per-leaf 16
run walk5_chunk1k --param kernel=0 --param walk_chunk=1024
run walk5_chunk32k --param kernel=0 --param walk_chunk=32768
timeout 600 python bench.py --nodes 100000 --queries 100000000 --steps 3 --warmup 3 --stats --no-cpu-baseline --e2e-queries 1000000 > gpurun_out/ab_cfg2big.json 2>gpurun_out/ab_cfg2big.err; cat gpurun_out/ab_cfg2big.json | python -c "import json,sys; d=json.load(sys.stdin); print('1e5 nodes', d['value'], d['roofline']['frac'], d.get('walk_stats'))"
