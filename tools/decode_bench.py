"""Time the decode chain (KF model 2 and V22 bbwt) on the cfg-2 corpus: Rice/gamma parse -> inverse MTF -> inverse BBWT."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
off = np.arange(0, n + 1, 1 << 20, dtype=np.int64)
d = torch.from_numpy(synth.s1_text(n)).cuda()
p = BlockPipeline(n, mib)
r = p.encode_device(d, off)
c = p.ctx
kf, kfo = r["kf_payload"].clone(), r["kf_off"]
k2, k2o = r["k2_payload"].clone(), r["k2_off"]
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3, out
t1, m = timed(lambda: c.rice_kf_decode(kf, kfo, off))
t1b, m2 = timed(lambda: c.rice_k2_decode(k2, k2o, off, 0))
t2, L = timed(lambda: c.mtf_decode(m, off))
t3, x = timed(lambda: c.bbwt_inverse(L, off))
assert torch.equal(x[:n], d[:n]) and torch.equal(m2[:n], m[:n])
tot = t1 + t2 + t3
print(f"decode {mib} MiB: rice_kf {t1:.1f} ms, rice_k2 {t1b:.1f} ms, mtf {t2:.1f} ms, bbwt_inv {t3:.1f} ms -> KF chain {tot:.1f} ms = {n/tot/1e3:.0f} MB/s")
p.profile_reset(); p.profile(True)
c.rice_kf_decode(kf, kfo, off); c.mtf_decode(m, off); c.bbwt_inverse(L, off); c.rice_k2_decode(k2, k2o, off, 0)
pr = p.profile_read(); p.profile(False)
print({k: (round(v["ms"], 2), v["launches"]) for k, v in pr.items() if v["launches"]})
