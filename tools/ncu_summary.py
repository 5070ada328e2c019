#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into the few numbers the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/<name>.txt]
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name", "?")[:110])
        for k in KEEP:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        stalls = [(float(v), h) for h, v in d.items()
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                  and v not in ("", "n/a")]
        if not stalls:
            stalls = [(float(v), h) for h, v in d.items()
                      if h.startswith("smsp__average_warp_latency_issue_stalled_") and v not in ("", "n/a")]
        stalls.sort(reverse=True)
        print("  -- top warp stall reasons (warps stalled per issue-active cycle):")
        for v, h in stalls[:8]:
            print("     %-60s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "")
                                        .replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    main()
