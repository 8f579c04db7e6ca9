#!/bin/bash
# round 2: 2 GPUs -- does the per-rank straggling (39 vs 53-62 ms with the statically dealt k_walk) survive the
# dynamically dealt k_walkc?  Per-CTA traces of every rank.
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 2 --no-strong --e2e-queries 1000000 --pageable-queries 0 --verify-queries 1000"
run() { name=$1; shift; timeout 300 "$@" --trace-out gpurun_out/r2n2b_trace_$name > gpurun_out/r2n2b_$name.json 2> gpurun_out/r2n2b_$name.err; python - <<PY
import json, numpy as np
try:
    d=json.load(open("gpurun_out/r2n2b_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms by rank", [round(x,2) for x in d["ms_per_step_by_rank"]], d["clocks"]["by_rank"]["sm_mhz"])
    for r in (0,1):
        try:
            tr=np.load("gpurun_out/r2n2b_trace_${name}_rank%d.npy"%r)
            dur=(tr[:,2].astype(np.int64)-tr[:,1].astype(np.int64))/1e6
            persm={}
            for sm,q in zip(tr[:,0],tr[:,3]): persm[int(sm)]=persm.get(int(sm),0)+int(q)
            v=np.array(sorted(persm.values()))
            print("   rank",r,"ctas",len(tr),"dur min/med/max %.2f %.2f %.2f"%(dur.min(),np.median(dur),dur.max()),"queries per SM min/med/max",v.min(),int(np.median(v)),v.max(),"sms",len(v))
        except Exception as e: print("   rank",r,"no trace",e)
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2n2b_$name.err").read()[-800:])
PY
}
for i in 1 2 3; do run walkc$i $B; done
for i in 1 2 3; do run walkc_static$i $B --param walk_dynamic=0; done
for i in 1 2; do run walkp_static$i $B --param walk_variant=1 --param walk_dynamic=0 --param leaf_stride=96; done
run old $B --param walk_variant=0 --param leaf_stride=96
