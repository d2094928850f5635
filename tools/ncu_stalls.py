#!/usr/bin/env python3
"""Per-source-line share of the warp-stall samples of one kernel (where the TIME goes, as opposed to tools/ncu_lines.py's executed
instructions) from an .ncu-rep captured with --set full --import-source on.  Prints the two dominant stall reasons of each line.
Usage: tools/ncu_stalls.py report.ncu-rep kernel_name [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []; fpath = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; isamp = hdr.index("# Samples")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= isamp or r[0] == "":
        continue
    try:
        n = int(r[isamp] or 0)
    except ValueError:
        continue
    reasons = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:2]
    lines.append((fpath, int(r[0]), r[1].strip()[:100], n, reasons))
tot = sum(l[3] for l in lines)
print("kernel %s: %d stall samples" % (kern, tot))
for f, ln, src, n, rs in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%5.1f%%  %-28s %s:%d  %s" % (100.0 * n / max(tot, 1), " ".join("%s %d%%" % (h, 100 * v // max(n, 1)) for v, h in rs if v), f, ln, src))
