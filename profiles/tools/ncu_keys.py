"""Print the handful of ncu raw-page metrics used in DESIGN.md / profiles/*.md from `ncu -i X --page raw --csv`."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
for i, h in enumerate(hdr):
    if h in keys or ('smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')):
        vals = [r[i] for r in data]
        try:
            if 'stalled' in h and float(vals[0]) < 0.05:
                continue
        except ValueError:
            pass
        print(f"{h:85s} {units[i]:14s} {vals}")
