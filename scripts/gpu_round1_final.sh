#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, both bench arms (N = 1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref exit $?"; wc -l gpurun_out/final_bench_ref.json
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; wc -l gpurun_out/final_bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench.json")); r=json.load(open("gpurun_out/final_bench_ref.json"))
print("ours value %.4e e2e %.4e frac %.4f launches %d"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("ref  value %.4e (%d cores)"%(r["value"], r["cpu_baseline"]["cores"]))
print(d["verify"])
PY
