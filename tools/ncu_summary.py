#!/usr/bin/env python
"""tools/ncu_summary.py -- the handful of ncu metrics that matter for an HBM-bound kernel, from .ncu-rep files.

    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...]      (runs `ncu -i ... --page raw --csv` here)
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(path, "no data")
        continue
    header, units = rows[0], rows[1]
    for row in rows[2:]:
        rec = dict(zip(header, row))
        print(f"== {path}: {rec.get('Kernel Name', '?')[:100]}")
        for k in KEYS:
            if k in rec:
                print(f"   {k:90s} {rec[k]:>18s} {units[header.index(k)]}")
