#!/bin/bash
# weak scaling on ONE 8-GPU box: N = 1, 2, 4, 8 back to back (10^9 queries per GPU per step)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --e2e-queries 20000000 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 5 --warmup 3 --e2e-queries 20000000 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N exit $?"
done
python - <<'PY'
import json
base=None
for N in (1,2,4,8):
    d=json.load(open(f"gpurun_out/scale_n{N}.json"))
    base=base or d["value"]
    print(f"N={N} value {d['value']:.4e} ms/step {d['ms_per_step']:.2f} speedup {d['value']/base:.2f} e2e {d['e2e']['value']:.3e} mismatches {d['verify']['fast_vs_exact_mismatches']['leaf']}")
PY
