#!/usr/bin/env python
"""Key metrics per captured kernel from `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("----", r[idx["Kernel Name"]][:90])
    for w in want:
        if w in idx:
            print(f"  {w:78s} {r[idx[w]][:24]:>16s} {units[idx[w]]}")
