#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/profile_raster.py"
$CMD > gpurun_out/x_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -o gpurun_out/prof_raster_r1d $CMD > gpurun_out/x_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/x_plain.log
