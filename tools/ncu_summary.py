#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the per-kernel table kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_kernels.csv"""
import csv
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time_us"), ("launch__grid_size", "grid"),
    ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps_per_cycle"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):  # the raw page already exported on the GPU box (report too large to copy back)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([c[1] + (f" [{units[hdr.index(c[0])]}]" if c[0] in hdr and units[hdr.index(c[0])] else "") for c in COLS])
        for r in rows[2:]:
            w.writerow([r[hdr.index(c[0])] if c[0] in hdr else "" for c in COLS])
    print("wrote", out, len(rows) - 2, "kernels")


if __name__ == "__main__":
    main()
