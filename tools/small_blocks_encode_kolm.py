import sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final as KF
n = 64 << 20
data = synth.s3_mix(n).tobytes()
KF.compress(data[:1 << 20])
t = time.perf_counter(); blob = KF.compress(data); dt = time.perf_counter() - t
print("KOLM default (CDC ~8 KiB) compress", round(n / dt / 1e6, 1), "MB/s", len(blob))
t = time.perf_counter(); back = KF.decompress(blob); dt = time.perf_counter() - t
print("KOLM decompress", round(n / dt / 1e6, 1), "MB/s", back == data)
