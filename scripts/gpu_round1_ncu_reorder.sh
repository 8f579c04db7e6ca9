#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/profile_locate.py 1e6 1e8 1 2 reorder=1"
$CMD > gpurun_out/nr_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_bin|k_tile_locate|k_unpermute' -s 3 -c 3 -o gpurun_out/prof_reorder_r1e $CMD > gpurun_out/nr_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/nr_plain.log
