"""Condense an ncu report (`ncu -i X.ncu-rep --page raw --csv`) to the metrics DESIGN.md quotes.
    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py ["comment"] > profiles/X_summary.csv"""
import csv, sys
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
     "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
     "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "launch__registers_per_thread", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "smsp__thread_inst_executed_per_inst_executed.ratio",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
h, u = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
cols = [m for m in M if m in idx]
for c in sys.argv[1:]:
    print("# " + c)
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + cols)
w.writerow([""] + [u[idx[m]] for m in cols])
for r in rows[2:]:
    w.writerow([r[idx["Kernel Name"]]] + [r[idx[m]] for m in cols])
