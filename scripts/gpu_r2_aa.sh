#!/bin/bash
# round 2, call AA: four-lane replay with groups refilled as they finish (k_exact_list4p)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "vertex_memo or fast_path_kernel or tiny_deferred or nonfinite or config3 or reorder_pipeline_edge or reentrant or weather or trivial or degenerate or entry_nodes" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2aa_pytest.log
for coop in 1 2; do python scripts/profile_cfg3.py 3 exact_coop=$coop | tail -1; done
for c in 6 8; do echo "== $c CTAs/SM"; GSI_LIB_PATH=$PWD/variants/libgsi_e$c.so python scripts/profile_cfg3.py 3 | tail -1; done
python scripts/profile_cfg3.py 3 vertex_memo=0 | tail -1
