"""Summarise ncu output brought back in gpurun_out/ into markdown tables for profiles/.
  python tools/ncu_summary.py launches gpurun_out/r01h_launches_c3.csv
  python tools/ncu_summary.py full gpurun_out/r01h_sift.ncu-rep
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5 and r[0].strip('"').isdigit()]
    acc = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        ns = float(r[-1])
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns
    tot = sum(v[1] for v in acc.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, errors="ignore")
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    want = OrderedDict([("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"),
                        ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
                        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
                        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instr"),
                        ("smsp__issue_active.avg.pct", "issue active %"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")])
    idx = [(hdr.index(k), v) for k, v in want.items() if k in hdr]
    units = rd[1]
    print("| " + " | ".join(v for _, v in idx) + " |\n|" + "---|" * len(idx))
    for r in rd[2:]:
        if len(r) < len(hdr):
            continue
        cells = []
        for i, v in idx:
            x = r[i].split("(")[0][:60] if v == "kernel" else r[i]
            if v in ("time", "dram read", "dram write"):
                x = "%s %s" % (r[i], units[i])
            cells.append(x)
        print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
