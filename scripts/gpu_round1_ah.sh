#!/bin/bash
mkdir -p gpurun_out
for mb in 128 192 224 256 288 320 384 512; do ./scripts/gather_probe3 0 $mb | grep -E "table|ld.global.nc.v4.u64|__ldg 4 B" | tr '\n' ' '; echo; done > gpurun_out/ah_gather_vs_table_size.txt 2>&1
cat gpurun_out/ah_gather_vs_table_size.txt
