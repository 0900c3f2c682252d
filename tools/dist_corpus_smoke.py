"""One-process check of dist.compress_kolr_fixed_corpus / decompress_kolr_corpus on the GPU (world size 1, NCCL): containers equal
compress_blocks_fixed of the drop-in, the sharded decompress returns the input.  usage: python tools/dist_corpus_smoke.py [MiB] [block]"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from kolmogorovlike_datacompressor_b200 import dist as kd, synth
from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
mix = synth.s3_mix(mib << 20)
cuts = [0, (mib << 20) // 3 + 777, (mib << 20) // 3 + 777, mib << 20]          # three containers, the middle one empty
parts = [mix[cuts[i]:cuts[i + 1]] for i in range(3)]
st = {}
t = time.perf_counter()
conts = kd.compress_kolr_fixed_corpus([len(p) for p in parts], lambda k, a, b: parts[k][a:b], bs, stats=st)
print("compress", round(time.perf_counter() - t, 3), {k: round(v, 3) for k, v in st.items() if k.endswith("_s")})
same = all(c == V.compress_blocks_fixed(p.tobytes(), bs) for c, p in zip(conts, parts))
pieces = kd.decompress_kolr_corpus(conts, gather=False)
ok = all(torch.equal(torch.from_numpy(parts[k][a:b]).cuda(), y[:b - a]) for k, a, b, y in pieces)
back = kd.decompress_kolr_corpus(conts)
print("containers equal the drop-in's:", same, " sharded decompress:", ok, " gathered decompress:", [bytes(b) == p.tobytes() for b, p in zip(back, parts)])
dist.destroy_process_group()
