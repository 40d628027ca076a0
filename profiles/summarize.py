#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the compact per-launch metric table kept under
profiles/: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<round>_<what>.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "kernel", "metric", "unit", "value"])
    for n, r in enumerate(rows[2:]):
        k = r[hdr.index("Kernel Name")]
        for m in WANT:
            if m in hdr:
                w.writerow([n, k[:90], m, units[hdr.index(m)], r[hdr.index(m)]])


if __name__ == "__main__":
    main(sys.argv[1])
