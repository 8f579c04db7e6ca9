#!/bin/bash
# round 2, call G: k_walk2 (two walkers per lane, cp.async staging) vs k_walk
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout=200 -k "fast_path_kernel or raster_odd or weather_golden or config1_golden or config3_mixed or tiny_deferred or nonfinite or degenerate" > gpurun_out/r2g_pytest_small.log 2>&1; echo "pytest small exit $?"; tail -4 gpurun_out/r2g_pytest_small.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-queries 1000000 --pageable-queries 0 --queries 268435456 --verify-queries 268435456 --param reorder=0"
run() { name=$1; shift; timeout 240 $B "$@" > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_$name.json"))
    print("$name", "value %.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["verify"]["fast_vs_exact_mismatches"], "deferred", d["verify"]["deferred_queries_per_step"], "kernel_ms", d["roofline"]["kernel_ms_per_launch"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/r2g_$name.err").read()[-1500:])
PY
}
run walk2
run walk1_w5 --param walk2=0 --param walk_ctas_per_sm=5
run walk2_w3 --param walk_ctas_per_sm=3
run walk2_s128 --param leaf_stride=128
run walk2_cpl2 --cells-per-leaf 2
run walk2_cpl8 --cells-per-leaf 8
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
for v in "reorder=0" "reorder=0 leaf_stride=128"; do
  tag=$(echo $v | tr ' =' '__')
  CMD="python scripts/profile_locate.py 1e6 100000000 1 2 $v"
  timeout 400 ncu --metrics $M --clock-control none -k regex:'k_walk' -s 1 -c 1 --csv --log-file gpurun_out/r2g_walk_$tag.csv $CMD > gpurun_out/r2g_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_walk2' -s 1 -c 1 -f -o gpurun_out/r2g_walk2_full python scripts/profile_locate.py 1e6 100000000 1 2 reorder=0 > gpurun_out/r2g_ncu_full.log 2>&1; echo "ncu full exit $?"
# config 2 shape (1e5 nodes, L2 resident)
for v in "walk2=1" "walk2=0"; do python scripts/profile_locate.py 1e5 50000000 1 3 reorder=0 $v > /dev/null 2>&1; done
python - <<'PY'
import subprocess, sys, json, time
sys.path.insert(0, '.')
import torch
import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import synth
n, Q = 100000, 50_000_000
xa, ya = synth.uniform_nodes(n); za = synth.values(n, 1)
for w2 in (1, 0):
    w = g.ScatteredCuda(n, 1); w.set_param("walk2", w2); w.set_param("reorder", 0); w.set_param("profile", 1)
    w.init(xa, ya, za, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, n))
    dev = torch.device("cuda:0")
    dx = torch.empty(Q, dtype=torch.float64, device=dev); dy = torch.empty_like(dx)
    dz = torch.empty((1, Q), dtype=torch.float64, device=dev); dl = torch.empty(Q, dtype=torch.int32, device=dev)
    w.generate_uniform_device(0, synth.SEED_QUERIES, 0, Q, 0.0, 1.0, 0.0, 1.0, dx.data_ptr(), dy.data_ptr())
    for _ in range(3): w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr())
    w.reset_counters()
    for _ in range(5): w.eval_device_e(0, dx.data_ptr(), dy.data_ptr(), Q, dz.data_ptr(), Q, dl.data_ptr())
    inf = w.info()
    print("1e5 nodes walk2=%d: locate %.3f ms per 5e7 -> %.3e q/s" % (w2, 1e3 * inf["locate_seconds"] / 5, Q / (inf["locate_seconds"] / 5)))
    w.free()
PY
