#!/bin/bash
# round 2, call F (final): state of the final sources -- whole GPU suite, smoke, default bench + reference arm, then the ncu
# evidence of the same sources (launch list of the bench command, --set full of k_walkc / pipeline / raster)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2f_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2> gpurun_out/r2f_bench.time; echo "bench exit $?"
( time timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err ) 2> gpurun_out/r2f_bench_ref.time; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["verify"]["cpu_reference_mismatches_on_strided_subsample"], "build", d["setup"]["host_build_seconds"], d["clocks"]["sm_mhz"])
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"], "cpu", d["cpu_baseline"]["value"])
for k,v in d.get("configs",{}).items(): print(k, {a:v.get(a) for a in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches")} if isinstance(v, dict) else v)
r=json.load(open("gpurun_out/r2f_bench_ref.json")); print("ref arm", r["value"], r["cpu_baseline"]["cores"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --verify-queries 1000 > gpurun_out/r2f_ncu_launches.log 2>&1; echo "launch list exit $?"
bash scripts/ncu_capture.sh
timeout 300 python scripts/profile_raster.py > gpurun_out/r2f_plain_raster.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_locate' -s 1 -c 1 -f -o gpurun_out/r2_raster_full \
   python scripts/profile_raster.py > gpurun_out/r2f_ncu_raster.log 2>&1; echo "ncu raster exit $?"
ncu -i gpurun_out/r2_raster_full.ncu-rep --page raw --csv > gpurun_out/r2_raster_full.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/r2_pipe_full.ncu-rep
