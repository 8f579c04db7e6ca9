#!/bin/bash
# round 2, final sources: 4 GPUs of one box -- default bench under torchrun at N = 4 and N = 2 (weak + strong legs), per-rank times
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n4c_gpus.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q --timeout=200 -k "two_devices" > gpurun_out/r2n4c_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2n4c_pytest.log
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n4c_bench_n$N.json 2> gpurun_out/r2n4c_bench_n$N.err; echo "bench N=$N exit $?"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2n4c_bench_n1.json 2> gpurun_out/r2n4c_bench_n1.err; echo "bench N=1 exit $?"
python - <<'PY'
import json
for N in (4, 2, 1):
    try:
        d=json.load(open(f"gpurun_out/r2n4c_bench_n{N}.json"))
        print(N, "weak value %.4e"%d["value"], "ms by rank", [round(x,2) for x in d["ms_per_step_by_rank"]], "e2e %.3e"%d["e2e"]["value"], "pageable %.3e"%d["e2e_pageable"]["value"])
        s=d.get("scaling_strong"); print("   strong", s and ("%.4e"%s["value"], [round(x,2) for x in s["ms_per_step_by_rank"]]))
        print("   clocks", d["clocks"]["by_rank"]["sm_mhz"], d["clocks"]["reasons"], d["verify"]["fast_vs_exact_mismatches"])
    except Exception as e:
        print(N, "failed", e); print(open(f"gpurun_out/r2n4c_bench_n{N}.err").read()[-1500:])
PY
