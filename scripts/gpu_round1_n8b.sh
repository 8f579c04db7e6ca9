#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv > gpurun_out/n8b_smi_before.csv
for rep in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$rep bench.py --gpus 8 --steps 5 --warmup 3 --e2e-queries 10000000 > gpurun_out/n8b_$rep.json 2> gpurun_out/n8b_$rep.err; echo "exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/n8b_$rep.json")); print("value %.4e ms %.2f"%(d["value"], d["ms_per_step"]), ["%.1f"%x for x in d["ms_per_step_by_rank"]], d["clocks"]["by_rank"])
PY
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv > gpurun_out/n8b_smi_after.csv; cat gpurun_out/n8b_smi_after.csv
