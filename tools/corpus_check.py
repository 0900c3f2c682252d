"""Large-input check of the container-sequence path: S3 mix of N MiB through corpus.compress_corpus / decompress_corpus (both
profiles), round trip + throughput.  usage: python tools/corpus_check.py [MiB] [block KiB]"""
import json, sys, time, warnings
sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import corpus, synth
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
data = synth.s3_mix(mib << 20).tobytes()
for prof in ("kolr", "kolm"):
    corpus.compress_corpus(data[:8 << 20], bk << 10, prof)
    t0 = time.perf_counter(); parts = corpus.compress_corpus(data, bk << 10, prof, max_bytes=(mib << 20) // 3 + 1); t1 = time.perf_counter()
    back = corpus.decompress_corpus(parts, prof); t2 = time.perf_counter()
    print(json.dumps({"profile": prof, "mib": mib, "block_kib": bk, "containers": len(parts), "bytes": sum(len(p) for p in parts),
                      "compress_MBps": round(len(data) / (t1 - t0) / 1e6, 1), "decompress_MBps": round(len(data) / (t2 - t1) / 1e6, 1),
                      "roundtrip": back == data}), flush=True)
