#!/usr/bin/env python3
"""Condense an ncu --set full report (ncu -i X.ncu-rep --page raw --csv) into one block of key metrics per launch.
Usage: python tools/ncu_summary.py gpurun_out/prof_TAG.ncu-rep > profiles/TAG_ncu_full_summary.txt"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_misc_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
idx = {}
for w in want:
    for i, h in enumerate(hdr):
        if h == w or h.endswith("." + w) or h.endswith(w): idx.setdefault(w, i)
for r in rows[2:]:
    print(f"== launch {r[0]}: {r[4]}  grid {r[8]} block {r[7]}")
    for w in want:
        if w in idx and r[idx[w]] not in ("", "no data"): print(f"   {w:86s} {r[idx[w]]:>18s} {units[idx[w]]}")
