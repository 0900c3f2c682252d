"""Per-category kernel times of the BBWT+MTF+Rice stage on the S3 mix (BASELINE cfg 5) for one block size.
usage: python tools/profile_cfg5.py [block_KiB] [MiB]"""
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline

bk = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n, block = mib << 20, bk << 10
data = synth.s3_mix(n)
off = np.arange(0, n + 1, block, dtype=np.int64)
pipe = BlockPipeline(n, len(off), profile_k2=False)
d = torch.from_numpy(data).cuda()
pipe.encode_device(d, off)
torch.cuda.synchronize()
pipe.profile_reset(); pipe.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pipe.encode_device(d, off); e1.record(); torch.cuda.synchronize()
prof = pipe.profile_read()
cnt = pipe.ctx.counters()
print(json.dumps({"block_kib": bk, "mib": mib, "ms": round(e0.elapsed_time(e1), 2), "MBps": round(n / e0.elapsed_time(e1) / 1e3, 1), "counters": cnt,
                  "kernel_ms": {k: round(v["ms"], 2) for k, v in prof.items() if v["launches"]},
                  "launches": {k: v["launches"] for k, v in prof.items() if v["launches"]}}))
