#!/usr/bin/env python
"""Summarise an ncu --set full report (one row per profiled launch): the metrics DESIGN.md quotes.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rN_x_summary.csv"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_static"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + [f"{k} [{units[col[k]]}]" for k in KEYS if k in col])
    for r in rows[2:]:
        w.writerow([r[col["Kernel Name"]][:60]] + [r[col[k]] for k in KEYS if k in col])


if __name__ == "__main__":
    main()
