#!/bin/bash
# host topology of an 8-GPU box and the effect of NUMA-near binding on the host-buffer (e2e) leg
mkdir -p gpurun_out
N=${1:-8}
( nvidia-smi topo -m; echo; lscpu | head -30; echo; cat /sys/devices/system/node/online 2>/dev/null; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done ) > gpurun_out/ac_topo.txt 2>&1
for bind in 1 0; do
  if [ $bind = 0 ]; then export GSI_NO_BIND=1; else unset GSI_NO_BIND; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$bind bench.py --gpus $N --steps 3 --warmup 3 --queries 200000000 > gpurun_out/ac_bench_n${N}_bind$bind.json 2> gpurun_out/ac_bench_n${N}_bind$bind.err; echo "bench n$N bind=$bind exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/ac_bench_n${N}_bind$bind.json")); print("value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["e2e"].get("cpu_affinity"))
PY
done
