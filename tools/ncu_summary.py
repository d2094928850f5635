#!/usr/bin/env python3
"""Markdown table of the key ncu metrics of every kernel launch in an .ncu-rep (read here, without a GPU).
Usage: tools/ncu_summary.py report.ncu-rep [frames_or_pairs_per_launch]"""
import csv, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {n: hdr.index(n) for n in hdr}
def g(r, n, d="-"):
    return r[col[n]] if n in col else d
print("| kernel | time us | warp instr (M) | instr / unit | lanes | issue active % | ALU pipe % | FMA pipe % | LSU pipe % | warps active % | regs | DRAM read MB | DRAM write MB | DRAM % of peak |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = g(r, "Kernel Name").split("(")[0].replace("void ", "")
    t = float(g(r, "gpu__time_duration.sum").replace(",", ""))
    tu = rows[1][col["gpu__time_duration.sum"]]
    if tu == "ms": t *= 1000.0
    inst = float(g(r, "smsp__inst_executed.sum").replace(",", ""))
    def f(n):
        try: return "%.1f" % float(g(r, n).replace(",", ""))
        except Exception: return "-"
    print("| %s | %.1f | %.2f | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        name, t, inst / 1e6, ("%.0f" % (inst / units)) if units else "-", f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        f("sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread"), f("dram__bytes_read.sum"), f("dram__bytes_write.sum"),
        f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
