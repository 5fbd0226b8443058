#!/usr/bin/env python
"""Print selected raw metrics of an .ncu-rep, one column per profiled launch."""
import csv, io, subprocess, sys
rep = sys.argv[1]
keys = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','local_load','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','l1tex__data_pipe_lsu_wavefronts.sum']
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for k in keys + sys.argv[2:]:
    if k in hdr:
        i = hdr.index(k); print(k[:72].ljust(72), rows[1][i][:8].ljust(8), [r[i][:14] for r in rows[2:]])
