#!/usr/bin/env python3
"""Per-CUDA-source-line totals (warp instructions executed, stall samples) of one kernel in an
`ncu --set full --import-source on` report.   usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import os
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout if len(sys.argv) < 5 else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre, "-s", sys.argv[4], "-c", "1"], capture_output=True, text=True).stdout
fpath, hdr, seen_fn, out = "", None, 0, []
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        seen_fn += 1; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not r[0] or seen_fn > len(set([fpath])) * 99:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        out.append((int(d["Instructions Executed"]), int(d["# Samples"]), fpath, int(r[0]), r[1].strip()[:110],
                    {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}))
    except (KeyError, ValueError):
        pass
ti, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total warp instr {ti}  samples {ts}")
for o in sorted(out, key=lambda o: -o[1])[:top]:
    st = ",".join(f"{k}:{v}" for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3])
    print(f"{o[0] / ti * 100:5.1f}%i {o[1] / ts * 100:5.1f}%s {o[2]}:{o[3]:<4} {o[4]}  [{st}]")
