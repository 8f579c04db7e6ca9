#!/bin/bash
mkdir -p gpurun_out
C="python scripts/profile_raster.py 16 2048 8192 2"
$C > gpurun_out/plain_raster_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -o gpurun_out/prof_raster_twophase_r1 $C > gpurun_out/ncu_raster_k.log 2>&1; echo "ncu raster exit $?"
