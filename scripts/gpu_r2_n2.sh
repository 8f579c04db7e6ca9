#!/bin/bash
# round 2: 2 GPUs -- what makes one rank's k_walk 35 % slower than the other's in the weak-scaling leg?
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --no-strong --e2e-queries 1000000 --pageable-queries 0 --verify-queries 1000"
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/r2n2_$name.json 2> gpurun_out/r2n2_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2n2_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms by rank", [round(x,2) for x in d["ms_per_step_by_rank"]], d["clocks"]["by_rank"]["sm_mhz"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2n2_$name.err").read()[-800:])
PY
}
run base1 $B
run base2 $B
run chunk1024 $B --param walk_chunk=1024
run chunk256 $B --param walk_chunk=256
run keepcache GSI_BENCH_KEEP_CACHE=1 $B
run novmm $B --param vmm=0
run q5e8 $B --queries 500000000
run exact_coop0 $B --param exact_coop=0
