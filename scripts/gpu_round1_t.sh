#!/bin/bash
# A/B of gather hints / streaming hints / leaf stride on the direct walk kernel, 10^6 nodes
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-queries 1000000 --queries 200000000 --param reorder=0"
run() { name=$1; lib=$2; shift; shift; GSI_LIB_PATH=$lib timeout 600 $B "$@" > gpurun_out/t_$name.json 2> gpurun_out/t_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t_$name.json"))
    print("$name", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/t_$name.err").read()[-1500:])
PY
}
BASE=$PWD/gsl_scattered_interpolation_b200/lib/libgsl_scattered_cuda.so
run base $BASE
run base_s128 $BASE --param leaf_stride=128
run v64 $PWD/variants/v64/libgsl_scattered_cuda.so
run v64_s128 $PWD/variants/v64/libgsl_scattered_cuda.so --param leaf_stride=128
run v128_s128 $PWD/variants/v128/libgsl_scattered_cuda.so --param leaf_stride=128
run vstream $PWD/variants/vstream/libgsl_scattered_cuda.so
run v64s $PWD/variants/v64s/libgsl_scattered_cuda.so
run v64s_w5 $PWD/variants/v64s/libgsl_scattered_cuda.so --param walk_ctas_per_sm=5
run v64s_cpl2 $PWD/variants/v64s/libgsl_scattered_cuda.so --cells-per-leaf 2
run v64s_cpl8 $PWD/variants/v64s/libgsl_scattered_cuda.so --cells-per-leaf 8
