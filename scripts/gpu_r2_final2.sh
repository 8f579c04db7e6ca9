#!/bin/bash
# round 2, call F2: final sources (deferred-count feedback for asynchronous callers): GPU suite, smoke, default bench,
# launch list of the bench command and the --set full capture of k_walkc that bench.py quotes (profiles/r2_traffic.json)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r2f2_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2f2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f2_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2f2_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2f2_bench.json 2> gpurun_out/r2f2_bench.err ) 2> gpurun_out/r2f2_bench.time; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f2_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["verify"]["cpu_reference_mismatches_on_strided_subsample"], "build", d["setup"]["host_build_seconds"], d["clocks"]["sm_mhz"])
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"], "cpu", d["cpu_baseline"]["value"])
for k,v in d.get("configs",{}).items(): print(k, {a:v.get(a) for a in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches")} if isinstance(v, dict) else v)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f2_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --verify-queries 1000 > gpurun_out/r2f2_ncu_launches.log 2>&1; echo "launch list exit $?"
Q=134217728
WALK="python scripts/profile_locate.py 1e6 $Q 1 2 reorder=0"
$WALK > gpurun_out/r2_plain_walk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_walk' -s 1 -c 1 \
    -f -o gpurun_out/r2_walk_full $WALK > gpurun_out/r2_ncu_walk.log 2>&1; echo "ncu k_walk exit $?"
ncu -i gpurun_out/r2_walk_full.ncu-rep --page raw --csv > gpurun_out/r2_walk_full.csv 2>/dev/null
# profiles/r2_traffic.json is then written here, on the build machine: python scripts/ncu_traffic.py profiles/r2_traffic.json ...
