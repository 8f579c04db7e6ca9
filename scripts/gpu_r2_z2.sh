#!/bin/bash
# round 2, call Z2: run-time choice of cp.async.ca / .cg for the leaf lines
mkdir -p gpurun_out
for l1 in 0 1; do echo "== 1e5 uniform nodes, walk_l1=$l1"; python scripts/profile_locate.py 1e5 50000000 1 5 walk_l1=$l1 | tail -1; done
for l1 in 0 1; do echo "== 3e5 uniform nodes, walk_l1=$l1"; python scripts/profile_locate.py 3e5 50000000 1 5 walk_l1=$l1 | tail -1; done
for l1 in 0 1; do echo "== 1e6 uniform nodes, walk_l1=$l1"; python scripts/profile_locate.py 1e6 134217728 1 3 walk_l1=$l1 | tail -1; done
for l1 in 0 1 -1; do echo "== config 3, walk_l1=$l1"; python scripts/profile_cfg3.py 3 walk_l1=$l1 | tail -1; done
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "fast_path_kernel or config3 or weather or tiny_deferred" 2>&1 | tail -2
