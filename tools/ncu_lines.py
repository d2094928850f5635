#!/usr/bin/env python3
"""Per-source-line share of executed warp instructions of one kernel from an .ncu-rep captured with --import-source on
(reads `ncu --page source --print-source cuda,sass --csv`).  Usage: tools/ncu_lines.py report.ncu-rep kernel_name [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []; fpath = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) <= ie:
        continue
    if r[0] != "":
        try:
            lines.append([fpath, int(r[0]), r[1].strip()[:110], int(r[ie] or 0), int(r[te] or 0)])
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
print("kernel %s: %d warp instructions, %.1f threads/instr" % (kern, tot, sum(l[4] for l in lines) / max(tot, 1)))
for f, n, s, v, t in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%5.1f%%  lanes %4.1f  %s:%d  %s" % (100.0 * v / tot, t / max(v, 1), f, n, s))
