#!/bin/bash
# round 2, call X: vertex memo (slab v7): GPU suite, config 3, default bench with configs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2x_pytest.log
timeout 300 python scripts/profile_cfg3.py 3 > gpurun_out/r2x_cfg3_plain.log 2>&1; tail -1 gpurun_out/r2x_cfg3_plain.log
timeout 300 python scripts/profile_cfg3.py 3 vertex_memo=0 > gpurun_out/r2x_cfg3_nomemo.log 2>&1; tail -1 gpurun_out/r2x_cfg3_nomemo.log
( time timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err ) 2> gpurun_out/r2x_bench.time; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2x_bench.json"))
print("value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], d["clocks"]["sm_mhz"], "setup", d["setup"])
for k,v in d.get("configs",{}).items(): print(k, {a:v.get(a) for a in ("value","frac","kernel_ms","exact_replay_ms","fast_vs_exact_mismatches")} if isinstance(v, dict) else v)
PY
