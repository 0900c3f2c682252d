"""Throughput of the Re-Pair candidate on the S3 mix.  usage: python tools/repair_bench.py [MiB] [block KiB]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.stages import Context
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n = mib << 20
d = torch.from_numpy(synth.s3_mix(n)).cuda()
off = np.arange(0, n + 1, bk << 10, dtype=np.int64)
c = Context(n, len(off))
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    out, oo = c.repair_encode(d, off)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("encode %d MiB in %d KiB blocks: %.2f s = %.1f MB/s, ratio %.3f" % (mib, bk, dt, n / dt / 1e6, int(oo[-1]) / n), flush=True)
sizes = np.diff(oo)
print("per-block ratios (first 8):", [round(float(s) / (bk << 10), 3) for s in sizes[:8]])
torch.cuda.synchronize(); t = time.perf_counter()
back = c.repair_decode(out, oo, off)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("decode %.2f s = %.1f MB/s, roundtrip %s" % (dt, n / dt / 1e6, bool(torch.equal(back[:n], d[:n]))))
