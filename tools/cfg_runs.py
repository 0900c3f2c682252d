"""Measurements for the non-headline BASELINE configs (3, 4, 5) at reduced corpus sizes; prints one JSON line per run.
usage: python tools/cfg_runs.py [cfg3|cfg4|cfg5|all] [MiB]"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200.stages import Context

what = sys.argv[1] if len(sys.argv) > 1 else "all"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = mib << 20


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps, out


def cfg3():
    data = synth.s2_mixed(n)
    d = torch.from_numpy(data).cuda()
    off = np.arange(0, n + 1, 1 << 20, dtype=np.int64)
    c = Context(n, mib)
    for (w, ml, wc) in ((255, 127, 0), (4096, 0, 4096), (65536, 0, 0)):
        t, (pay, po) = timed(lambda: c.lz77_encode(d, off, w, ml), reps=1)
        td, back = timed(lambda: c.lz77_decode(pay, po, off, wc), reps=1)
        ok = bool(torch.equal(back[:n], d[:n]))
        print(json.dumps({"cfg": 3, "corpus": "S2 mixed", "mib": mib, "window": w, "max_len": ml, "encode_MBps": round(n / t / 1e6, 1),
                          "decode_MBps": round(n / td / 1e6, 1), "ratio": round(int(po[-1]) / n, 4), "roundtrip": ok}), flush=True)
    c.close()


def cfg4():
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    import warnings
    warnings.simplefilter("ignore")
    data = synth.s3_mix(n).tobytes()
    for name, fn, dec in (("KOLR fixed 1 MiB", lambda: V.compress_blocks_fixed(data, 1 << 20), V.decompress),
                          ("KOLM target 1 MiB (CDC)", lambda: KF.compress(data, 1 << 20), KF.decompress)):
        t0 = time.perf_counter(); blob = fn(); t = time.perf_counter() - t0
        t0 = time.perf_counter(); blob = fn(); t = min(t, time.perf_counter() - t0)
        t0 = time.perf_counter(); back = dec(blob); td = time.perf_counter() - t0
        print(json.dumps({"cfg": 4, "corpus": "S3 mix", "mib": mib, "call": name, "compress_MBps": round(n / t / 1e6, 1),
                          "decompress_MBps": round(n / td / 1e6, 1), "ratio": round(len(blob) / n, 4), "roundtrip": back == data}), flush=True)


def cfg5():
    data = synth.s3_mix(n)
    d = torch.from_numpy(data).cuda()
    for bs in (64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20):
        off = np.arange(0, n + 1, bs, dtype=np.int64)
        c = Context(n, len(off))
        def enc():
            L = c.bbwt_forward(d, off); m = c.mtf_encode(L, off); return c.rice_kf_encode(m, off) + (m,)
        t, (pay, po, m) = timed(enc, reps=1)
        rounds = c.counters()
        def dec():
            return c.bbwt_inverse(c.mtf_decode(c.rice_kf_decode(pay, po, off), off), off)
        td, back = timed(dec, reps=1)
        print(json.dumps({"cfg": 5, "corpus": "S3 mix", "mib": mib, "block_bytes": bs, "encode_MBps": round(n / t / 1e6, 1),
                          "decode_MBps": round(n / td / 1e6, 1), "ratio_kf_model2": round(int(po[-1]) / n, 4),
                          "rounds_plain": rounds["rounds_plain"], "rounds_cyclic": rounds["rounds_cyclic"],
                          "roundtrip": bool(torch.equal(back[:n], d[:n]))}), flush=True)
        c.close(); del c
        torch.cuda.empty_cache()


for f in ([cfg3, cfg4, cfg5] if what == "all" else [globals()[what]]):
    f()
