"""Print the handful of ncu metrics this project tracks from an `ncu --page raw --csv` dump."""
import csv, sys
KEEP = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
 'dram__bytes_read.sum.per_second','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','lts__t_sectors.sum','lts__t_sectors.sum.pct_of_peak_sustained_elapsed',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','launch__registers_per_thread',
 'sm__warps_active.avg.pct_of_peak_sustained_active','sm__issue_active.avg.pct_of_peak_sustained_elapsed',
 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warp_latency_per_inst_issued.ratio',
 'smsp__sass_average_branch_targets_threads_uniform.pct','launch__grid_size','launch__block_size','launch__occupancy_limit_registers',
 'smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_srcunit_tex_lookup_hit.sum','lts__t_sectors_srcunit_tex_lookup_miss.sum',
 'lts__average_t_sector_hit_rate_srcunit_tex_realtime.pct','lts__d_sectors_fill_device.sum','lts__t_sectors_srcnode_gpc_op_read_lookup_miss.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:90] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in KEEP:
            print(f"  {h:85s} {units[i]:16s} {r[i]}")
