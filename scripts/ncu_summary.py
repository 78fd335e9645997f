#!/usr/bin/env python
"""Print the handful of ncu metrics we track from a .ncu-rep (raw page) + lane-utilisation buckets from the source page."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size"]
d = dict(zip(hdr, zip(units, vals)))
for w in want:
    if w in d:
        print(f"{w:75s} {d[w][1]} {d[w][0]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(d[h][1])
        if v > 0.3:
            print(f"stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
ci, ct = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
tot = 0; b = [0, 0, 0, 0]
for r in rows[2:]:
    try: ie, te = int(r[ci]), int(r[ct])
    except Exception: continue
    tot += ie
    lanes = te / max(ie, 1)
    b[0 if lanes < 4 else 1 if lanes < 12 else 2 if lanes < 24 else 3] += ie
print("warp-instructions", tot, "lane buckets <4,<12,<24,>=24:", [f"{100*x/max(tot,1):.1f}%" for x in b])
