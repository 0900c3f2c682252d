#!/usr/bin/env python3
"""Print a compact summary of bench.py's JSON line (stdin).  usage: python bench.py ... | python tools/summarize.py [label]"""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)

label = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline", {})
print(label, "value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "launches", d.get("gpu_launches"))
if r:
    print("  dominant", r["kernel"], "achieved", r["achieved"], "frac", r["frac"], "share", r["share_of_step"])
    print("  per_category_ms", r["per_category_ms"])
    print("  stages", r["stages"])
