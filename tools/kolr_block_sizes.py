"""compress_blocks_fixed / decompress of the KOLR drop-in on the S3 mix at several block sizes (second call at each size).
usage: python tools/kolr_block_sizes.py [MiB] [block sizes ...]"""
import sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
import torch
from kolmogorovlike_datacompressor_b200 import synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sizes = [int(v) for v in sys.argv[2:]] or [2048, 8192, 16384, 65536]
mix = synth.s3_mix(mib << 20)
for bs in sizes:
    n = min(mib << 20, 65535 * bs)
    data = mix[:n].tobytes()
    V.compress_blocks_fixed(data[:min(n, 8 << 20)], bs)
    V.compress_blocks_fixed(data, bs)
    torch.cuda.synchronize(); t = time.perf_counter(); blob = V.compress_blocks_fixed(data, bs); torch.cuda.synchronize(); dt = time.perf_counter() - t
    st = V._engine().ctx.encode_blocks_stats()
    V.decompress(blob)
    t = time.perf_counter(); back = V.decompress(blob); dd = time.perf_counter() - t
    names = V._parse(blob)[0]
    hist = {}
    for nm in names:
        hist[nm] = hist.get(nm, 0) + 1
    print("block %7d  %4d MiB  compress %7.1f MB/s  decompress %7.1f MB/s  ratio %.3f  stopped %d/%d  %s  roundtrip %s" % (
        bs, n >> 20, n / dt / 1e6, n / dd / 1e6, len(blob) / n, st["repair_stopped_early"], len(names), hist, back == data), flush=True)
