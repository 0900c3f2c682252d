#!/usr/bin/env python3
"""Print stall reasons and key throughput metrics per kernel of an `ncu --set full` report.  usage: ncu_stalls.py report.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")[:14] for r in rows[2:]]
print(" " * 40, " ".join(n.rjust(14) for n in names))
want = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "not_issued" not in h and h.endswith("ratio")]
want += ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for h in want:
    if h not in hdr:
        continue
    i = hdr.index(h)
    short = h.replace("smsp__average_warps_issue_stalled_", "st_").replace("_per_issue_active.ratio", "")[:40]
    print(short.ljust(40), " ".join(r[i][:12].rjust(14) for r in rows[2:]))
