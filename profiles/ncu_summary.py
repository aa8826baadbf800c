#!/usr/bin/env python
"""Readable summary of an `ncu --set full --page raw --csv` dump (one row per launch):
    python profiles/ncu_summary.py profiles/r2/r2f_default_raw.csv.gz > profiles/r2/r2f_ncu_summary.txt"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
rows = list(csv.reader(io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem"),
        ("launch__occupancy_limit_registers", "occ_reg"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
        ("smsp__inst_executed.sum", "warp_instr"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall_barrier%"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio")]
cols = [(hdr.index(k), n) for k, n in want if k in hdr]
print(f"# {path}: {len(data)} launches (one view-step: conv1 = launches 0-4, conv2 = launches 5-9)")
for li, r in enumerate(data):
    print(f"--- launch {li}")
    for i, n in cols:
        v = r[i]
        try:
            v = f"{float(v.replace(',', '')):.4g}"
        except ValueError:
            v = v[:100]
        print(f"    {n:16s} {v} {units[i]}")
