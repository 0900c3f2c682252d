import sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
n = 32 << 20
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
data = synth.s3_mix(n).tobytes()
V.compress_blocks_fixed(data[:1 << 20], bs)
t = time.perf_counter(); blob = V.compress_blocks_fixed(data, bs); dt = time.perf_counter() - t
print("KOLR", bs, "compress", round(n / dt / 1e6, 1), "MB/s", len(blob))
