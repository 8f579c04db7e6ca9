#!/bin/bash
# round 2, call H: the whole GPU suite at HEAD, smoke, the default bench (both arms) as the driver runs them
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 ) > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2h_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2h_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err ) 2> gpurun_out/r2h_bench.time; echo "bench exit $?"; tail -3 gpurun_out/r2h_bench.time
( time timeout 600 python bench.py --impl reference > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err ) 2> gpurun_out/r2h_bench_ref.time; echo "ref exit $?"; tail -3 gpurun_out/r2h_bench_ref.time
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["verify"]["cpu_reference_mismatches_on_strided_subsample"])
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"] if d.get("e2e_pageable") else None)
print("cpu", d["cpu_baseline"])
for k,v in (d.get("configs") or {}).items(): print(k, v if not isinstance(v, dict) else {kk: v[kk] for kk in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches") if kk in v})
r=json.load(open("gpurun_out/r2h_bench_ref.json")); print("ref", r["value"], r["config"]["workload"]==d["config"]["workload"])
PY
