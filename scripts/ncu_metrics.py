"""Reads `ncu -i X.ncu-rep --page raw --csv` output (stdin or file) and prints selected metrics."""
import csv, sys, re
path = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
    "launch__occupancy", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum$", "sm__inst_executed_pipe_",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__cycles_elapsed.avg$", "smsp__cycles_active.avg$",
    "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "launch__waves", "pipe_fma", "pipe_alu", "pipe_fp64", "shared_ld", "shared_st", "wavefronts_mem_shared"]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
for i, h in enumerate(hdr):
    if any(re.search(p, h) for p in pats):
        print(f"{h} [{units[i]}]: {[r[i] for r in rows[2:]]}")
