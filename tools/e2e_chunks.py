"""Time BlockPipeline.encode_host with different chunk counts (H2D/D2H overlap study)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline
n = 256 << 20
off = np.arange(0, n + 1, 1 << 20, dtype=np.int64)
h = torch.from_numpy(synth.s1_text(n)).pin_memory()
p = BlockPipeline(n, 256)
d = h.cuda()
for _ in range(2):
    p.encode_device(d, off)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3):
    p.encode_device(d, off)
torch.cuda.synchronize(); print("device-resident ms", (time.perf_counter() - t) / 3 * 1e3)
for ch in (1, 2, 4, 8):
    p.encode_host(h, off, chunks=ch)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3):
        r = p.encode_host(h, off, chunks=ch)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print("chunks", ch, "ms", round(dt * 1e3, 2), "MB/s", round(n / dt / 1e6, 1), "d2h", r["d2h_bytes"])
