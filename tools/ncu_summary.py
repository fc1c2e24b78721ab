#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) as a small text file for profiles/.

usage: ncu_summary.py <report.ncu-rep> [out.txt]
Reads the raw page through `ncu -i ... --page raw --csv` and prints, per captured launch, the
duration, DRAM bytes, L2 hit rate, achieved occupancy and the warp-stall breakdown.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_fp64.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out.write("# ncu --set full summary of %s\n" % rep.split("/")[-1])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.write("\nkernel: %s\n" % d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                out.write("  %-86s %16s %s\n" % (k, d[k], units[hdr.index(k)]))
    if out is not sys.stdout:
        out.close()


if __name__ == "__main__":
    main()
