#!/bin/bash
# round 2, call Y: vertex-memo pass ahead of the replay kernels (list compaction)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "vertex_memo or fast_path_kernel or tiny_deferred or nonfinite or config3 or reorder_pipeline_edge or reentrant or weather" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2y_pytest.log
timeout 300 python scripts/profile_cfg3_classes.py 2>&1 | tail -4
timeout 300 python scripts/profile_cfg3.py 3 2>&1 | tail -1
timeout 300 python scripts/profile_cfg3.py 3 vertex_memo=0 2>&1 | tail -1
timeout 300 python scripts/profile_cfg3.py 3 exact_coop=0 2>&1 | tail -1
