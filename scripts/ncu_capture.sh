#!/bin/bash
# Regenerates the ncu evidence bench.py quotes (profiles/r2_traffic.json) for the sources in the
# tree: run on the GPU box (one GPU), AFTER the same command has exited 0 without ncu.
#   gpurun -- 'bash scripts/ncu_capture.sh'   -> gpurun_out/r2_*.ncu-rep, r2_*.csv, r2_traffic.json
# then copy gpurun_out/r2_traffic.json (and the csv summaries) into profiles/.
set -u
mkdir -p gpurun_out
Q=134217728
PIPE="python scripts/profile_locate.py 1e6 $Q 1 2 reorder=1"
WALK="python scripts/profile_locate.py 1e6 $Q 1 2 reorder=0"
$PIPE > gpurun_out/r2_plain_pipe.log 2>&1 || { echo "plain pipeline run failed"; tail -5 gpurun_out/r2_plain_pipe.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_bin|k_tile_walk|k_unpermute' -s 3 -c 3 \
    -f -o gpurun_out/r2_pipe_full $PIPE > gpurun_out/r2_ncu_pipe.log 2>&1; echo "ncu pipeline exit $?"
ncu -i gpurun_out/r2_pipe_full.ncu-rep --page raw --csv > gpurun_out/r2_pipe_full.csv 2>/dev/null
$WALK > gpurun_out/r2_plain_walk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_walk' -s 1 -c 1 \
    -f -o gpurun_out/r2_walk_full $WALK > gpurun_out/r2_ncu_walk.log 2>&1; echo "ncu k_walk exit $?"
ncu -i gpurun_out/r2_walk_full.ncu-rep --page raw --csv > gpurun_out/r2_walk_full.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/r2_traffic.json $Q 1000000 1 pipeline=gpurun_out/r2_pipe_full.csv \
    k_walk=gpurun_out/r2_walk_full.csv source="scripts/ncu_capture.sh, ncu --set full --clock-control none"
