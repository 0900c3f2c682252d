#!/usr/bin/env python3
"""Print a compact summary of bench.py's JSON line (stdin).  usage: python bench.py ... | python tools/summarize.py [label]  or  python tools/summarize.py FILE"""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)

import os

label = sys.argv[1] if len(sys.argv) > 1 else ""
if label and os.path.isfile(label):                        # a file name instead of stdin (never block on an empty stdin)
    text = open(label).read()
elif sys.stdin.isatty():
    sys.exit("usage: bench.py ... | summarize.py [label]   or   summarize.py FILE")
else:
    text = sys.stdin.read()
line = [l for l in text.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline", {})
print(label, "value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "launches", d.get("gpu_launches"))
if r:
    print("  dominant", r["kernel"], "achieved", r["achieved"], "frac", r["frac"], "share", r["share_of_step"])
    print("  per_category_ms", r["per_category_ms"])
    print("  stages", r["stages"])
