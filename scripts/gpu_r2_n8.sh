#!/bin/bash
# round 2, final sources: 8 GPUs of one box -- default bench under torchrun (weak + strong legs), per-rank times
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n8_gpus.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n8_bench_n8.json 2> gpurun_out/r2n8_bench_n8.err; echo "bench N=8 exit $?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2n8_bench_n8.json"))
    print("weak value %.4e"%d["value"], "ms by rank", [round(x,2) for x in d["ms_per_step_by_rank"]], "e2e %.3e"%d["e2e"]["value"], "pageable %.3e"%d["e2e_pageable"]["value"])
    s=d.get("scaling_strong"); print("   strong", s and ("%.4e"%s["value"], [round(x,2) for x in s["ms_per_step_by_rank"]]))
    print("   clocks", d["clocks"]["by_rank"]["sm_mhz"], d["clocks"]["reasons"], d["verify"]["fast_vs_exact_mismatches"], d["setup"].get("slab_broadcast_seconds"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2n8_bench_n8.err").read()[-2000:])
PY
