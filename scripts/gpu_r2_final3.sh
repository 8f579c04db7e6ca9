#!/bin/bash
# round 2, call F3: final sources after removing the dead experiment switches: GPU suite, default bench (short), ncu --set full of k_walkc
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 ) > gpurun_out/r2f2_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2f2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f2_smoke.log 2>&1; echo "smoke exit $?"
Q=134217728
WALK="python scripts/profile_locate.py 1e6 $Q 1 2 reorder=0"
$WALK > gpurun_out/r2_plain_walk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_walk' -s 1 -c 1 \
    -f -o gpurun_out/r2_walk_full $WALK > gpurun_out/r2_ncu_walk.log 2>&1; echo "ncu k_walk exit $?"
ncu -i gpurun_out/r2_walk_full.ncu-rep --page raw --csv > gpurun_out/r2_walk_full.csv 2>/dev/null
rm -f gpurun_out/r2_walk_full.ncu-rep
