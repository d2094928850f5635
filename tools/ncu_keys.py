"""Print the key metrics of every kernel in an .ncu-rep (run here, no GPU needed): tools/ncu_keys.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    for i in idx:
        print("%-90s %s %s" % (hdr[i], r[i], rows[1][i]))
    print()
