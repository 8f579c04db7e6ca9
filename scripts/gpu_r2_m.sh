#!/bin/bash
# round 2, call M: new defaults (128-byte leaves, k_walkc): whole GPU suite, smoke, default bench + reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2m_smoke.log
( time timeout 900 python bench.py --trace-out gpurun_out/r2m_trace > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err ) 2> gpurun_out/r2m_bench.time; echo "bench exit $?"
python - <<'PY'
import json, numpy as np
d=json.load(open("gpurun_out/r2m_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["verify"]["cpu_reference_mismatches_on_strided_subsample"], "build", d["setup"]["host_build_seconds"])
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"])
for k,v in d.get("configs",{}).items(): print(k, {a:v.get(a) for a in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches")} if isinstance(v, dict) else v)
tr=np.load("gpurun_out/r2m_trace_rank0.npy")
dur=(tr[:,2].astype(np.int64)-tr[:,1].astype(np.int64))/1e6
print("trace: ctas", len(tr), "start spread ms %.3f"%((tr[:,1].max()-tr[:,1].min())/1e6), "dur min/med/max %.2f %.2f %.2f"%(dur.min(), np.median(dur), dur.max()), "queries/cta min/max", tr[:,3].min(), tr[:,3].max())
PY
