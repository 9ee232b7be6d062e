"""Summarise an .ncu-rep: key raw metrics per launch + SASS region breakdown.  usage: ncu_summary.py rep [kernel-regex]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','launch__block_size',
 'smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed.avg.per_cycle_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.max']
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== kernel:', r[hdr.index('Kernel Name')][:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w); print(f'  {w} = {r[i]} {units[i]}')
