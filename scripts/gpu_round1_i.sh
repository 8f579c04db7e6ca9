#!/bin/bash
mkdir -p gpurun_out
echo "=== config 3 (list 1/4)"; timeout 600 python scripts/bench_configs.py 3 2>gpurun_out/i3.err | tee gpurun_out/i_cfg3.json | cut -c1-700
C="python scripts/profile_raster.py 16 2048 8192 0"
$C > gpurun_out/plain_raster.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -o gpurun_out/prof_raster_walk_r1 $C > gpurun_out/ncu_raster.log 2>&1; echo "ncu raster walk exit $?"
C="python scripts/profile_raster.py 16 2048 8192 1"
$C > gpurun_out/plain_raster1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -o gpurun_out/prof_raster_simple_r1 $C > gpurun_out/ncu_raster1.log 2>&1; echo "ncu raster simple exit $?"
