"""Per-round trace and per-category times of BBWT+MTF+KF-Rice on the S3 mix (cfg 4/5 data).  usage: python tools/trace_s3.py [MiB] [block KiB]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
bk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n = mib << 20
data = synth.s3_mix(n)
off = np.arange(0, n + 1, bk << 10, dtype=np.int64)
pipe = BlockPipeline(n, len(off) - 1, device=0)
d = torch.from_numpy(data).cuda()
pipe.encode_device(d, off)
torch.cuda.synchronize()
t = time.perf_counter(); pipe.encode_device(d, off); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("encode_device %.1f ms  %.1f MB/s" % (dt * 1e3, n / dt / 1e6))
pipe.profile_reset(); pipe.profile(True); pipe.encode_device(d, off); prof = pipe.profile_read(); pipe.profile(False)
print({k: round(v["ms"], 2) for k, v in prof.items() if v["ms"] > 0.05})
print(pipe.ctx.counters())
