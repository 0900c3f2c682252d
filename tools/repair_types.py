"""Re-Pair time per block type of the S3 mix (one block per launch).  usage: python tools/repair_types.py [block KiB]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.stages import Context
bk = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = bk << 10
mix = synth.s3_mix(8 << 20)
names = ["S1", "S1", "S2g", "S2s", "S2p", "S2c", "S1", "random"]
c = Context(n, 4)
off = np.array([0, n], dtype=np.int64)
for t in (tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 2, 3, 4, 5, 7)):
    d = torch.from_numpy(mix[t << 20:(t << 20) + n].copy()).cuda()
    c.repair_encode(d, off)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, oo = c.repair_encode(d, off)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    p = out[:int(oo[-1])].cpu().numpy()
    # 'R','P', ULEB 256 (2 bytes), ULEB nrules
    k, sh, i = 0, 0, 4
    while True:
        b = int(p[i]); k |= (b & 0x7F) << sh; sh += 7; i += 1
        if b < 128: break
    print("%-7s %7.3f s  rules %7d  payload %8d  -> %.1f us/round" % (names[t], dt, k, int(oo[-1]), dt / max(k, 1) * 1e6), flush=True)
