"""Kernel time distribution of KOLR decompress on the S3 mix at 1 MiB blocks (run under
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`): the profiler range covers the second decompress only."""
import sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = mib << 20
data = synth.s3_mix(n).tobytes()
blob = V.compress_blocks_fixed(data, 1 << 20)
V.decompress(blob)
torch.cuda.synchronize()
torch.cuda.profiler.start()
t = time.perf_counter(); back = V.decompress(blob); torch.cuda.synchronize(); dt = time.perf_counter() - t
torch.cuda.profiler.stop()
print("KOLR decompress %d MiB: %.3f s = %.1f MB/s, ok %s" % (mib, dt, n / dt / 1e6, back == data))
