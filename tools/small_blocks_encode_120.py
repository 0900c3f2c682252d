import sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
n = 120 << 20
data = synth.s3_mix(n).tobytes()
for rep in range(2):
    t = time.perf_counter(); blob = V.compress_blocks_fixed(data, 2048); dt = time.perf_counter() - t
    print("KOLR 2KiB compress", rep, round(n / dt / 1e6, 1), "MB/s", len(blob), flush=True)
