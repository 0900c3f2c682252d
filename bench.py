#!/usr/bin/env python3
"""bench.py — headline benchmark of the block-transform hot path (BASELINE.json configs[1]):

    BBWT + MTF + Rice stage, 1 MiB blocks, 256 MiB synthetic low-entropy text-like corpus per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # the reference's own CPU code on host cores

One step = one pass of the hot path over the whole per-GPU corpus (256 blocks of 1 MiB):
Lyndon factorisation + BBWT rotation sort -> MTF -> KF model-2 token coder (kolm_final.py model 2) and the five
V22 Rice(k=2) variants' exact costs + one packed variant (kolm_final_researched_v2-2.py models 2-6).
`value`  : uncompressed MB/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through BlockPipeline.encode_host_many (pinned host input -> H2D -> kernels -> D2H payloads, every step;
           copies of neighbouring steps overlap the kernels); `single_batch_call` = BlockPipeline.encode_host, nothing overlapped.
`roofline`: the dominant kernel category, algorithmic bytes / event-timed duration vs MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MIB = 1 << 20
METRIC = "compress_throughput_bbwt_mtf_rice"
UNIT = "MB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(busy)}


# -------------------------------------------------------------------------------------------------
def cpu_port_baseline(corpus, block, sample_blocks):
    """The oracle (CPU restatement) on a bounded sample of the same workload, 1 thread."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    nbytes = 0
    for b in range(sample_blocks):
        blk = corpus[b * block:(b + 1) * block].tobytes()
        L = O.bbwt_forward(blk)
        m = O.mtf_encode(L)
        O.kf_rice_pack(m)
        for f in (0, 1, 4, 8, 16):
            O.v22_rice_pack(m, f)
        nbytes += len(blk)
    dt = time.perf_counter() - t0
    return {"value": round(nbytes / dt / 1e6, 3), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample_blocks} x {block >> 10} KiB blocks of the same corpus through oracle/kolm_oracle.cpp "
                      f"(bbwt+mtf+kf pack+5 k2 packs), {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.

    oracle/_ref/kolm_final_cpp (= /root/reference/final/kolm_final.cpp, g++ -O3, hard-coded 8 KiB CDC blocks;
    BWT+MTF+Rice/LZ77/XOR selection) when it was built, one process per core on disjoint slices;
    else the oracle port (one thread per core via processes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from kolmogorovlike_datacompressor_b200 import synth
    cores = len(os.sched_getaffinity(0))
    exe = os.path.join(ROOT, "oracle", "_ref", "kolm_final_cpp")
    per_core = (2 if os.path.exists(exe) else 1) * MIB
    corpus = synth.s1_text(per_core * cores)
    times = []
    tmp = tempfile.mkdtemp(prefix="kolm_ref_")
    kind = "reference" if os.path.exists(exe) else "port"
    if kind == "reference":
        for c in range(cores):
            corpus[c * per_core:(c + 1) * per_core].tofile(os.path.join(tmp, f"in{c}.bin"))

    def one_step():
        t0 = time.perf_counter()
        if kind == "reference":
            ps = [subprocess.Popen([exe, "-c", os.path.join(tmp, f"in{c}.bin"), os.path.join(tmp, f"out{c}.bin")]) for c in range(cores)]
            for p in ps:
                if p.wait() != 0:
                    raise RuntimeError("kolm_final_cpp failed")
        else:
            code = ("import sys; sys.path.insert(0, %r); from oracle import oracle as O; d=open(sys.argv[1],'rb').read();"
                    "L=O.bbwt_forward(d); m=O.mtf_encode(L); O.kf_rice_pack(m); [O.v22_rice_pack(m,f) for f in (0,1,4,8,16)]") % ROOT
            for c in range(cores):
                corpus[c * per_core:(c + 1) * per_core].tofile(os.path.join(tmp, f"in{c}.bin"))
            ps = [subprocess.Popen([sys.executable, "-c", code, os.path.join(tmp, f"in{c}.bin")]) for c in range(cores)]
            for p in ps:
                p.wait()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    for _ in range(args.steps):
        times.append(one_step())
    total = per_core * cores * args.steps
    dt = sum(times)
    val = round(total / dt / 1e6, 3)
    sample = (f"{cores} processes x {per_core // MIB} MiB slices of the S1 corpus per step; "
              + ("kolm_final.cpp -O3 (own 8 KiB CDC blocks, its full model selection)" if kind == "reference"
                 else "oracle port, 1 MiB blocks"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1000 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "cfg2: BBWT+MTF+Rice stage, 1 MiB blocks, 256 MiB S1 text-like corpus per GPU",
                       "reference_sample": "bounded CPU sample of the same corpus; kolm_final.cpp cuts its own 8 KiB CDC blocks (hard-coded) and runs its "
                                           "full model selection" if kind == "reference" else "bounded CPU sample of the same corpus, 1 MiB blocks",
                       "block_bytes": 8192 if kind == "reference" else MIB},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    block = args.block_kib << 10
    nbytes = args.mib * MIB
    nblocks = (nbytes + block - 1) // block
    off = np.minimum(np.arange(nblocks + 1, dtype=np.int64) * block, nbytes)
    # weak scaling: every rank compresses its own shard of the job (independent blocks, no data-path collective)
    corpus = synth.s1_text(nbytes, seed=0xC0FFEE + rank)
    pipe = BlockPipeline(nbytes, nblocks, device=local)
    h_in = torch.from_numpy(corpus).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        r = pipe.encode_device(d_in, off)
    barrier()
    pipe.profile_reset()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        r = pipe.encode_device(d_in, off)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = gmax(ev0.elapsed_time(ev1))
    counters = pipe.ctx.counters()
    launches = counters["launches"]
    kf_bytes, k2_bytes = int(r["kf_off"][-1]), int(r["k2_off"][-1])
    value = world * nbytes * args.steps / (ms / 1e3) / 1e6

    # ---- end to end through the public call (host buffers) -----------------------------------
    # (a) one batch per call; (b) the streaming call a multi-batch corpus goes through: every step still copies its 256 MiB in and
    # its payloads out, but the copies of neighbouring steps run under the kernels (two input buffers, two pinned payload buffers)
    for _ in range(1):
        pipe.encode_host(h_in, off)
    barrier()
    t0 = time.perf_counter()
    single_steps = max(1, min(args.steps, 3))
    for _ in range(single_steps):
        res = pipe.encode_host(h_in, off)
    barrier()
    single_s = gmax(time.perf_counter() - t0)
    for res in pipe.encode_host_many((h_in, off) for _ in range(2)):
        pass
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)
    d2h = 0
    for res in pipe.encode_host_many((h_in, off) for _ in range(e2e_steps)):
        d2h += res["d2h_bytes"]
    barrier()
    e2e_s = gmax(time.perf_counter() - t0)
    e2e = {"value": round(world * nbytes * e2e_steps / e2e_s / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": res["h2d_bytes"] * world,
           "d2h_bytes_per_step": d2h // e2e_steps * world, "steps": e2e_steps, "call": "BlockPipeline.encode_host_many (streaming, copies overlap the neighbouring step)",
           "single_batch_call": {"value": round(world * nbytes * single_steps / single_s / 1e6, 2), "unit": UNIT, "steps": single_steps,
                                 "call": "BlockPipeline.encode_host (copy in, encode, copy out; nothing overlapped across calls)"}}

    # ---- decode chain (device resident): KF payload -> Rice/gamma parse -> inverse MTF -> inverse BBWT --------------
    c = pipe.ctx
    kf_dev, kf_off = r["kf_payload"], r["kf_off"]
    def decode_once():
        m_ = c.rice_kf_decode(kf_dev, kf_off, off)
        return c.bbwt_inverse(c.mtf_decode(m_, off), off)
    x = decode_once()
    roundtrip_ok = bool(torch.equal(x[:nbytes], d_in[:nbytes]))
    barrier()
    dsteps = max(1, min(args.steps, 3))
    dv0, dv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dv0.record()
    for _ in range(dsteps):
        decode_once()
    dv1.record()
    barrier()
    dms = gmax(dv0.elapsed_time(dv1))
    decode = {"metric": "decompress_throughput_rice_mtf_bbwt", "value": round(world * nbytes * dsteps / (dms / 1e3) / 1e6, 2), "unit": UNIT,
              "ms_per_step": round(dms / dsteps, 3), "steps": dsteps, "roundtrip_bit_exact": roundtrip_ok}
    del x

    # ---- per-kernel event timing (one extra, untimed-for-value step) -----------------------------
    pipe.profile_reset()
    pipe.profile(True)
    pipe.encode_device(d_in, off)
    prof = pipe.profile_read()
    pipe.profile(False)
    rounds = pipe.ctx.counters()
    peak, peak_src = peaks()
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    dom_name, dom_v = dom
    ach = dom_v["alg_bytes"] / (dom_v["ms"] / 1e3) / 1e9 if dom_v["ms"] > 0 else 0.0
    step_ms_prof = sum(v["ms"] for v in prof.values())
    stages = {}
    P = 5
    Rp, Rc = rounds["rounds_plain"], rounds["rounds_cyclic"]
    sort_ms = sum(prof[k]["ms"] for k in ("boot_keys", "radix_hist", "radix_scan", "radix_scatter", "rerank", "apply_ranks", "gather", "plan",
                                          "build_tiles", "lyndon_scan", "bbwt_emit"))
    model_bytes = nbytes * ((Rc * (36 + 24 * P) + 6) + (Rp * (36 + 24 * P) + 8))
    mtf_ms = sum(prof[k]["ms"] for k in ("mtf_pre", "mtf_scan", "mtf_main"))
    rice_ms = sum(prof[k]["ms"] for k in ("rice_cost", "rice_plan", "rice_pack", "zero_fill"))
    def frac(b, t):
        return round(b / (t / 1e3) / 1e9 / peak, 4) if t > 0 else None
    stages["bbwt_sort"] = {"ms": round(sort_ms, 3), "rounds_plain": Rp, "rounds_cyclic": Rc,
                           "model_bytes_per_input_byte": (Rc + Rp) * (36 + 24 * P) + 14, "frac_of_peak_model": frac(model_bytes, sort_ms),
                           "records_sorted_per_input_byte": round(rounds["records_sorted"] / nbytes, 3)}
    stages["mtf"] = {"ms": round(mtf_ms, 3), "alg_bytes_per_input_byte": 2, "frac_of_peak": frac(2 * nbytes, mtf_ms)}
    c_ratio = (kf_bytes + k2_bytes) / nbytes
    stages["rice"] = {"ms": round(rice_ms, 3), "alg_bytes_per_input_byte": round(4 + c_ratio, 3), "frac_of_peak": frac((4 + c_ratio) * nbytes, rice_ms)}
    traffic, traffic_note = ncu_traffic(dom_name, args)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "traffic_note": traffic_note, "alg_bytes_per_launch": int(dom_v["alg_bytes"] / max(1, dom_v["launches"])),
                "peak_source": peak_src, "launches": dom_v["launches"], "share_of_step": round(dom_v["ms"] / step_ms_prof, 3),
                "per_category_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}, "stages": stages}

    if rank == 0:
        cpu = cpu_port_baseline(corpus, block, args.cpu_blocks) if world == 1 and args.cpu_blocks > 0 else None
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": "cfg2: BBWT+MTF+Rice stage, 1 MiB blocks, 256 MiB S1 text-like corpus per GPU",
                           "corpus_mib_per_gpu": args.mib, "block_bytes": block, "blocks_per_gpu": int(nblocks),
                           "l2": "inputs (256 MiB) and scratch (GBs) are larger than L2, no flush needed",
                           "outputs": {"kf_payload_bytes": kf_bytes, "k2_payload_bytes": k2_bytes}},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "decode": decode}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic(category, args):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/r1d_traffic.json, else r1c; written by tools/make_profile_summary.py).
    ncu captured only the largest launches of the category, so the figure is their mean; it is not measured by this run."""
    kname = {"rerank": "k_rerank", "radix_scatter": "k_radix_scatter", "gather": "k_gather"}.get(category)
    pdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    path = next((os.path.join(pdir, t + "_traffic.json") for t in ("r1d", "r1c") if os.path.isfile(os.path.join(pdir, t + "_traffic.json"))), "")
    if kname is None or not path or args.mib != 256 or args.block_kib != 1024:
        return None, "no ncu capture for this kernel/workload"
    doc = json.load(open(path))
    rows = [(k, r) for k, v in doc["kernels"].items() if k.startswith(kname) for r in v]
    if not rows:
        return None, "kernel not in the ncu capture"
    gmax_ = max(r["grid"] for _, r in rows)
    rows = [(k, r) for k, r in rows if r["grid"] == gmax_]           # the full-size launches (all records of the batch)
    mean = sum(r["dram_read_bytes"] + r["dram_write_bytes"] for _, r in rows) / len(rows)
    # algorithmic bytes of the SAME launches (grid = tiles of 4096 records; bytes per record as in the KL() launch macros)
    # k_rerank<1>: key32 bootstrap; <2> (deep bootstrap) and <0> (rounds) also gather a second key
    per_rec = {"k_rerank": lambda k: 16 if k.startswith("k_rerank<1") else 20, "k_radix_scatter": lambda k: 16, "k_gather": lambda k: 4}[kname]
    alg = sum(r["grid"] * 4096 * per_rec(k) for k, r in rows) / len(rows)
    return int(mean), "mean of the %d full-size launches in %s (%s): DRAM %.2f GB vs %.2f GB algorithmic for those launches (ratio %.2f)" % (
        len(rows), "profiles/" + os.path.basename(path), doc["source"], mean / 1e9, alg / 1e9, mean / alg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mib", type=int, default=256, help="corpus MiB per GPU")
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--cpu-blocks", type=int, default=16, help="blocks in the bounded cpu_baseline sample (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
