#!/bin/bash
# round 2, call O: in-process NVML sampler check; ncu --set full + source of k_walkc
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --verify-queries 1000 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], d["clocks"])
PY
CMD="python scripts/profile_locate.py 1e6 67108864 1 2"
$CMD > gpurun_out/r2o_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_walkc' -s 1 -c 1 -f -o gpurun_out/r2o_walkc_full $CMD > gpurun_out/r2o_ncu.log 2>&1; echo "ncu exit $?"; tail -1 gpurun_out/r2o_plain.log
