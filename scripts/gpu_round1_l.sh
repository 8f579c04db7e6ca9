#!/bin/bash
mkdir -p gpurun_out
echo "=== config 4"; timeout 600 python scripts/bench_configs.py 4 2>gpurun_out/l4.err | tee gpurun_out/l_cfg4.json | cut -c1-700
C="python scripts/profile_raster.py 16 2048 8192 2"
$C > gpurun_out/plain_raster_l.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -o gpurun_out/prof_raster_twophase_r1b $C > gpurun_out/ncu_raster_l.log 2>&1; echo "ncu raster exit $?"
