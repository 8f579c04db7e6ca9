#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct
for cfg in "0 224" "32 224" "128 224" "0 1024" "32 64"; do
  set -- $cfg
  ./scripts/gather_probe3 $1 $2 > gpurun_out/s_plain_$1_$2.txt 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/s_ncu_$1_$2.csv ./scripts/gather_probe3 $1 $2 > /dev/null 2>&1
  echo "cfg $cfg exit $?"; cat gpurun_out/s_plain_$1_$2.txt
done
