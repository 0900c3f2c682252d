#!/usr/bin/env python3
"""Summarise an ncu launch list (csv) and an `ncu --set full` report into profiles/<tag>_summary.md.
usage: python tools/make_profile_summary.py <tag> <launches.csv> <report.ncu-rep> "<command>" """
import collections
import csv
import subprocess
import sys

tag, launches, rep, cmd = sys.argv[1:5]
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out = [f"# {tag} ncu summary", "", f"Command: `{cmd}` (one B200, after the same command exited 0 without ncu).", "",
       "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
       "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
out.append(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | 1.000 |")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def g(r, n):
    return r[hdr.index(n)]
def scaled(r, n, want):
    v, u = float(g(r, n)), units[hdr.index(n)]
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
    return v * f / want
out += ["", "## `ncu --set full --clock-control none --import-source on` (selected kernels)", "",
        "| kernel | grid | time us | dram read MB | dram write MB | dram % | sm % | warps active % | regs | warp instr |",
        "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows[2:]:
    out.append("| `%s` | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %d |" % (
        g(r, "Kernel Name").split("(")[0].replace("void ", ""), g(r, "launch__grid_size"), scaled(r, "gpu__time_duration.sum", 1),
        scaled(r, "dram__bytes_read.sum", 1e6), scaled(r, "dram__bytes_write.sum", 1e6),
        float(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")), float(g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")),
        float(g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")), g(r, "launch__registers_per_thread"), int(float(g(r, "smsp__inst_executed.sum")))))
# machine-readable DRAM traffic per captured launch (bench.py reads it for roofline.traffic)
import json
traffic = collections.defaultdict(list)
for r in rows[2:]:
    name = g(r, "Kernel Name").split("(")[0].replace("void ", "")
    traffic[name].append({"grid": int(g(r, "launch__grid_size")), "time_us": round(scaled(r, "gpu__time_duration.sum", 1), 1),
                          "dram_read_bytes": int(scaled(r, "dram__bytes_read.sum", 1)), "dram_write_bytes": int(scaled(r, "dram__bytes_write.sum", 1))})
json.dump({"command": cmd, "source": f"ncu --set full, {rep.split('/')[-1]}", "kernels": traffic}, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
open(f"profiles/{tag}_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
