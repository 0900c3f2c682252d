"""Times the reference's own Python code paths (north_star: "kolm_final.cpp (-O3) and Python paths") on a bounded prefix of the
bench corpora.  The Python reference lives only in the build container (/root/reference is not on the GPU box), so this
script runs HERE, single-threaded as the reference is, and writes profiles/python_reference_timing.json, which bench.py quotes
(labelled with where it was measured).
usage: python tools/time_python_reference.py [KiB]      (default 64 KiB of S1 text and of the S3 mix)"""
import json
import os
import platform
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kolmogorovlike_datacompressor_b200 import synth      # noqa: E402
from oracle import ref_loader as R                           # noqa: E402

kib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
KF, V = R.load_kf(), R.load_v22()
res = {"measured_in": "build container (the Python reference cannot travel to the GPU box)", "host": platform.processor() or platform.machine(),
       "cores_used": 1, "python": platform.python_version(), "prefix_kib": kib, "runs": {}}
for cname, corpus in (("s1_text", synth.s1_text(kib << 10).tobytes()), ("s3_mix_first_segments", bytes(synth.s3_mix(8 << 20)[::128][:kib << 10]))):
    for name, enc, dec in (("kolm_final.py compress(target_block=8192)", lambda d: KF.compress(d), KF.decompress),
                           ("kolm_final_researched_v2-2.py compress_blocks_fixed(2048)", lambda d: V.compress_blocks_fixed(d, 2048), V.decompress)):
        t0 = time.perf_counter(); blob = enc(corpus); t1 = time.perf_counter(); back = dec(blob); t2 = time.perf_counter()
        assert back == corpus
        res["runs"][f"{name} on {cname}"] = {"bytes": len(corpus), "container_bytes": len(blob), "compress_s": round(t1 - t0, 2), "decompress_s": round(t2 - t1, 2),
                                             "compress_MBps": round(len(corpus) / (t1 - t0) / 1e6, 5), "decompress_MBps": round(len(corpus) / (t2 - t1) / 1e6, 5)}
        print(name, cname, res["runs"][f"{name} on {cname}"], flush=True)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "python_reference_timing.json")
json.dump(res, open(out, "w"), indent=1)
print("wrote", out)
