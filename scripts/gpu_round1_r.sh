#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
CMD="python scripts/profile_locate.py 1e6 1e8 1 2 reorder=1"
$CMD > gpurun_out/r_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm|k_walk|k_exact_list' --csv --log-file gpurun_out/r_reorder_launches.csv $CMD > gpurun_out/r_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r_plain.log
CMD="python scripts/profile_locate.py 1e6 1e8 1 2 reorder=0"
$CMD > gpurun_out/r_plain0.log 2>&1 && ncu --metrics $M --clock-control none -k regex:'k_bin|k_tile|k_unperm|k_walk|k_exact_list' --csv --log-file gpurun_out/r_direct_launches.csv $CMD > gpurun_out/r_ncu0.log 2>&1
echo "ncu exit $?"
