import cProfile, pstats, sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
from kolmogorovlike_datacompressor_b200 import kolm_final as KF
n = 128 << 20
data = synth.s3_mix(n).tobytes()
V.compress_blocks_fixed(data[:8 << 20], 1 << 20)
for name, fn in (("KOLR", lambda: V.compress_blocks_fixed(data, 1 << 20)), ("KOLM", lambda: KF.compress(data, 1 << 20))):
    fn()
    t = time.perf_counter(); pr = cProfile.Profile(); pr.enable(); blob = fn(); pr.disable(); dt = time.perf_counter() - t
    print(name, "compress", round(n / dt / 1e6, 1), "MB/s", len(blob))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
for name, fn, dec in (("KOLR", lambda: V.compress_blocks_fixed(data, 1 << 20), V.decompress), ("KOLM", lambda: KF.compress(data, 1 << 20), KF.decompress)):
    blob = fn(); dec(blob)
    t = time.perf_counter(); pr = cProfile.Profile(); pr.enable(); back = dec(blob); pr.disable(); dt = time.perf_counter() - t
    print(name, "decompress", round(n / dt / 1e6, 1), "MB/s", back == data)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(16)
