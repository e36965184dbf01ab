#!/usr/bin/env python3
"""Summarise an ncu report (exported with `ncu -i X.ncu-rep --page raw --csv`) for the kernels of this repo:
duration, DRAM bytes, pipe/issue utilisation, occupancy and the warp-stall breakdown.
Usage: python profiles/ncu_summary.py report.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub not in name:
            continue
        print("===", name[:100])
        for k in KEYS:
            if k in hdr:
                print("  %-80s %18s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("  stall %-40s %8.3f warps/issue" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))


if __name__ == "__main__":
    main()
