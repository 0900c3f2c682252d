import cProfile, pstats, sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 32) << 20
data = synth.s3_mix(n).tobytes()
V.compress_blocks_fixed(data[:1 << 20], 2048)
blob = V.compress_blocks_fixed(data, 2048)
t = time.perf_counter(); pr = cProfile.Profile(); pr.enable(); blob = V.compress_blocks_fixed(data, 2048); pr.disable(); dt = time.perf_counter() - t
print("KOLR 2KiB compress", round(n / dt / 1e6, 1), "MB/s", len(blob), "Re-Pair stopped early:", V._engine().ctx.encode_blocks_stats())
for _ in range(2):
    t = time.perf_counter(); V.compress_blocks_fixed(data, 2048); print("  unprofiled", round(n / (time.perf_counter() - t) / 1e6, 1), "MB/s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
V.decompress(blob)
t = time.perf_counter(); pr = cProfile.Profile(); pr.enable(); back = V.decompress(blob); pr.disable(); dt = time.perf_counter() - t
print("KOLR 2KiB decompress", round(n / dt / 1e6, 1), "MB/s", back == data)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
