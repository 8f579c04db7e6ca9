#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; echo "bench exit $?"; wc -l gpurun_out/ab_bench.json; tail -3 gpurun_out/ab_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ab_bench.json"))
print("value %.4e"%d["value"], "frac %.4f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
print(d["verify"]); print(d["cpu_baseline"]); print(d["clocks"]); print(d["roofline"]["traffic"], d["roofline"]["ncu"])
PY
