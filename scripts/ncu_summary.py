#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full` report (the metrics profiles/README.md quotes), as CSV:
python scripts/ncu_summary.py report.ncu-rep > profiles/<name>_ncu_full.csv"""
import csv, subprocess, sys
want=['gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed.sum','smsp__inst_executed_pipe_fp64.sum','smsp__inst_executed_pipe_lsu.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; idx={h:i for i,h in enumerate(hdr)}
print('metric,' + ','.join(f'launch{i}' for i in range(len(rows)-2)) + ',unit')
print('kernel,' + ','.join(r[idx['Kernel Name']].split('(')[0][-24:] for r in rows[2:]))
for w in want:
    if w in idx: print(w + ',' + ','.join(r[idx[w]] for r in rows[2:]) + ',' + units[idx[w]])
