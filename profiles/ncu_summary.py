#!/usr/bin/env python
"""Print the handful of ncu metrics we steer by from a .ncu-rep (read here, no GPU needed).

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[-1]
        print(f"== {name}   ({len(rs)} captures, last shown)")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:72s} {r[i]:>18s} {units[i]}")
        if "--stalls" in sys.argv:
            st = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
            for v, h in sorted(st, reverse=True)[:8]:
                print(f"   stall {h:66s} {v:10.3f}")


main()
