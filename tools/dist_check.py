"""Multi-GPU check of the sharded drop-ins (NCCL): the containers assembled from block ranges encoded on N GPUs equal the
single-GPU containers byte for byte; prints the sharded throughput.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py [MiB]"""
import json, os, sys, time, warnings
sys.path.insert(0, ".")
import torch, torch.distributed as dist
warnings.simplefilter("ignore")
from kolmogorovlike_datacompressor_b200 import dist as kd, synth
from kolmogorovlike_datacompressor_b200 import kolm_final as KF, kolm_final_researched_v2_2 as V

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = synth.s3_mix(mib << 20).tobytes()
res = {}
for name, sharded, single, dec, sdec in (("KOLR fixed 1 MiB", lambda: kd.compress_kolr_fixed(data, 1 << 20), lambda: V.compress_blocks_fixed(data, 1 << 20), V.decompress, kd.decompress_kolr),
                                         ("KOLM target 1 MiB", lambda: kd.compress_kolm(data, 1 << 20), lambda: KF.compress(data, 1 << 20), KF.decompress, kd.decompress_kolm)):
    sharded(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); blob = sharded(); dist.barrier(); dt = time.perf_counter() - t0
    # every rank needs the container for the sharded decode
    box = [blob]
    dist.broadcast_object_list(box, src=0)
    blob = box[0]
    sdec(blob); dist.barrier()
    t0 = time.perf_counter(); back = sdec(blob); dist.barrier(); ddt = time.perf_counter() - t0
    if rank == 0:
        single(); t0 = time.perf_counter(); ref = single(); d1 = time.perf_counter() - t0
        dec(blob); t0 = time.perf_counter(); ok = dec(blob) == data; d2 = time.perf_counter() - t0
        res[name] = {"identical": blob == ref, "roundtrip": ok, "sharded_roundtrip": back == data, "sharded_MBps": round(len(data) / dt / 1e6, 1),
                     "single_gpu_MBps": round(len(data) / d1 / 1e6, 1), "sharded_decode_MBps": round(len(data) / ddt / 1e6, 1),
                     "single_gpu_decode_MBps": round(len(data) / d2 / 1e6, 1), "container_bytes": len(blob)}
    dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "mib": mib, "results": res}))
dist.destroy_process_group()
