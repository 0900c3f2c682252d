import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.stages import Context
d = synth.s3_mix(8 << 20)
c = Context(1 << 20, 4)
for k in range(8):
    blk = torch.from_numpy(d[k << 20:(k + 1) << 20].copy()).cuda()
    off = np.array([0, 1 << 20], dtype=np.int64)
    pay, po = c.lz77_encode(blk, off, 4096, 0)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        back = c.lz77_decode(pay, po, off, 4096)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(k, "payload", int(po[-1]), "decode ms", round(dt * 1e3, 2), bool(torch.equal(back[:1 << 20], blk)))
