#!/usr/bin/env python3
"""bench.py — benchmark of the block-transform hot path over the whole BASELINE.json matrix, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]               # our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # the reference's own CPU code on host cores

Headline (`value`, `e2e`, `roofline`) = BASELINE configs[1] (cfg 2): BBWT + MTF + Rice stage, 1 MiB blocks, 256 MiB synthetic
low-entropy text-like corpus per GPU.  One step = one pass of the hot path over the whole per-GPU corpus (256 blocks of 1 MiB):
Lyndon factorisation + BBWT rotation sort -> MTF -> KF model-2 token coder (kolm_final.py model 2) and the five V22 Rice(k=2)
variants' exact costs + one packed variant (kolm_final_researched_v2-2.py models 2-6).
`value`   : uncompressed MB/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`     : the same through BlockPipeline.encode_host_many (pinned host input -> H2D -> kernels -> D2H payloads, every step).
`roofline`: the dominant kernel category, algorithmic bytes / event-timed duration vs MEASURED_PEAKS.json; `stages` = MTF / Rice / sort.
`parity`  : the GPU payloads of the first blocks compared, in this run, with the CPU oracle's (the run fails if they differ).
`legs`    : the other BASELINE configs, bounded in time, each with its round-trip flag and sampled oracle payload checks:
    cfg3_lz77            LZ77 match search + greedy parse, three (window, cap) parameter sets, 256 MiB of the S2 mix (N = 1)
    cfg4_full_pipeline   kolm_final_researched_v2-2 compress_blocks_fixed + decompress with per-block model selection, all ten
                         candidates exact, S3 corpus of 1 GiB containers at 1 MiB blocks, BLOCK-SHARDED over the N GPUs with the NCCL
                         gather of the compressed stream inside the timed region (strong scaling: the corpus does not grow with N);
                         `default_block`: the same corpus, sharded the same way, at the default block size of compress_blocks_fixed
                         (8 KiB; 16 containers of 256 MiB); at N = 1 also kolm_final compress/decompress (KOLM)
    cfg5_block_sweep     block sizes 64 KiB .. 16 MiB, encode and decode, S3 mix, per GPU (weak)
    default_block_sizes  the reference's default block sizes (2 KiB KOLR, 8 KiB KOLM) through the drop-ins' fused per-batch call (N = 1)
"""
from __future__ import annotations

import argparse
import hashlib
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
K2_FLAGS = (0, 1, 4, 8, 16)


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
# CPU side: the oracle (test infrastructure) as checker and as the timed CPU baseline
# -------------------------------------------------------------------------------------------------
def oracle_stage(blk: bytes):
    """The cfg-2 step of one block through the CPU oracle -> (bbwt, mtf, kf payload, [five k2 payloads])."""
    from oracle import oracle as O
    L = O.bbwt_forward(blk)
    m = O.mtf_encode(L)
    return L, m, O.kf_rice_pack(m), [O.v22_rice_pack(m, f) for f in K2_FLAGS]


def cpu_port_baseline(corpus, block, sample_blocks):
    """The oracle (CPU restatement) on a bounded sample of the same workload, 1 thread.  Returns (record, per-block outputs)."""
    t0 = time.perf_counter()
    nbytes, outs = 0, []
    for b in range(sample_blocks):
        blk = corpus[b * block:(b + 1) * block].tobytes()
        outs.append(oracle_stage(blk))
        nbytes += len(blk)
    dt = time.perf_counter() - t0
    return {"value": round(nbytes / dt / 1e6, 3), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample_blocks} x {block >> 10} KiB blocks of the same corpus through oracle/kolm_oracle.cpp "
                      f"(bbwt+mtf+kf pack+5 k2 packs), {dt:.1f} s"}, outs


_PORT_CODE = ("import sys; sys.path.insert(0, %r); from oracle import oracle as O; D=open(sys.argv[1],'rb').read(); B=int(sys.argv[2])\n"
              "for a in range(0, len(D), B):\n"
              "    d=D[a:a+B]; L=O.bbwt_forward(d); m=O.mtf_encode(L); O.kf_rice_pack(m); [O.v22_rice_pack(m,f) for f in (0,1,4,8,16)]\n") % ROOT


def cpu_port_all_cores(block=MIB, blocks_per_core=3, seed_off=0):
    """The same-work CPU figure: the oracle port on every host core at once (one process per core, one 1 MiB block of the S1
    corpus each per step) — what the reference's algorithm costs on this box at the bench's block size."""
    from kolmogorovlike_datacompressor_b200 import synth
    cores = len(os.sched_getaffinity(0))
    corpus = synth.s1_text(block * cores * blocks_per_core, seed=0xC0FFEE + seed_off)
    tmp = tempfile.mkdtemp(prefix="kolm_port_")
    per = block * blocks_per_core
    for c in range(cores):
        corpus[c * per:(c + 1) * per].tofile(os.path.join(tmp, f"in{c}.bin"))
    from oracle import oracle as O
    O.lib()                                                     # built before the clock starts
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, "-c", _PORT_CODE, os.path.join(tmp, f"in{c}.bin"), str(block)]) for c in range(cores)]
    for p in ps:
        if p.wait() != 0:
            raise RuntimeError("oracle port process failed")
    dt = time.perf_counter() - t0
    return {"value": round(per * cores / dt / 1e6, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} processes x {blocks_per_core} block(s) of {block >> 10} KiB of the S1 corpus, the cfg-2 step through "
                      f"oracle/kolm_oracle.cpp, {dt:.1f} s (process start included)"}


def python_reference_record():
    """Timings of the reference's own Python files (kolm_final.py, kolm_final_researched_v2-2.py).  The Python reference cannot
    travel to the GPU box, so these were measured in the build container by tools/time_python_reference.py and are quoted."""
    p = os.path.join(ROOT, "profiles", "python_reference_timing.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.

    oracle/_ref/kolm_final_cpp (= /root/reference/final/kolm_final.cpp, g++ -O3, hard-coded 8 KiB CDC blocks;
    BWT+MTF+Rice/LZ77/XOR selection) when it was built, one process per core on disjoint slices;
    else the oracle port (one thread per core via processes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from kolmogorovlike_datacompressor_b200 import synth
    cores = len(os.sched_getaffinity(0))
    exe = os.path.join(ROOT, "oracle", "_ref", "kolm_final_cpp")
    per_core = (2 if os.path.exists(exe) else 1) * MIB
    corpus = synth.s1_text(per_core * cores)
    times = []
    tmp = tempfile.mkdtemp(prefix="kolm_ref_")
    kind = "reference" if os.path.exists(exe) else "port"
    for c in range(cores):
        corpus[c * per_core:(c + 1) * per_core].tofile(os.path.join(tmp, f"in{c}.bin"))

    def one_step():
        t0 = time.perf_counter()
        if kind == "reference":
            ps = [subprocess.Popen([exe, "-c", os.path.join(tmp, f"in{c}.bin"), os.path.join(tmp, f"out{c}.bin")]) for c in range(cores)]
        else:
            ps = [subprocess.Popen([sys.executable, "-c", _PORT_CODE, os.path.join(tmp, f"in{c}.bin"), str(MIB)]) for c in range(cores)]
        for p in ps:
            if p.wait() != 0:
                raise RuntimeError("reference process failed")
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
    # the same-work figure beside it: the reference's algorithm at the bench's own block size (1 MiB), all cores
    try:
        line["same_work_port_all_cores"] = cpu_port_all_cores()
    except Exception as e:                                       # the headline of this arm stands without it
        line["same_work_port_all_cores"] = {"error": str(e)}
    py = python_reference_record()
    if py:
        line["python_reference"] = py
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"                        # the version banner goes to stdout: the line must stay alone there
        if self.world == 1 and "MASTER_PORT" not in os.environ:
            import socket
            s = socket.socket(); s.bind(("127.0.0.1", 0)); os.environ["MASTER_PORT"] = str(s.getsockname()[1]); s.close()
            os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        # one NCCL group at every N (a single rank at N = 1): the sharded legs run the same code path at 1, 2, 4 and 8 GPUs
        import datetime
        dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=torch.device("cuda", self.local),
                                timeout=datetime.timedelta(seconds=240))   # a rank that fell out of step fails the run in minutes, not in ten
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def gmax(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gsum(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gall(self, ok: bool) -> bool:
        return self.gsum(0.0 if ok else 1.0) == 0.0


def ev_pair(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def leg_cfg2(env, args):
    """Headline: cfg 2.  Returns the fields of the JSON line that belong to it."""
    import numpy as np
    torch = env.torch
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline
    world, rank, local = env.world, env.rank, env.local
    block = args.block_kib << 10
    nbytes = args.mib * MIB
    nblocks = (nbytes + block - 1) // block
    off = np.minimum(np.arange(nblocks + 1, dtype=np.int64) * block, nbytes)
    # weak scaling: every rank compresses its own shard of the job (independent blocks, no data-path collective)
    corpus = synth.s1_text(nbytes, seed=0xC0FFEE + rank)
    # the CPU baseline runs beside the GPU work (the oracle releases the GIL); its outputs are the parity check's expectation
    cpu_box = {}
    cpu_thread = None
    if rank == 0 and args.cpu_blocks > 0:
        def cpu_job():
            cpu_box["res"] = cpu_port_baseline(corpus, block, min(args.cpu_blocks, int(nblocks)))
        cpu_thread = threading.Thread(target=cpu_job, daemon=True)
    pipe = BlockPipeline(nbytes, nblocks, device=local)
    h_in = torch.from_numpy(corpus).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        r = pipe.encode_device(d_in, off)
    env.barrier()
    pipe.profile_reset()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = ev_pair(torch)
    env.barrier()
    ev0.record()
    for _ in range(args.steps):
        r = pipe.encode_device(d_in, off)
    ev1.record()
    env.barrier()
    clocks = sampler.stop()
    ms = env.gmax(ev0.elapsed_time(ev1))
    counters = pipe.ctx.counters()
    launches = counters["launches"]
    kf_bytes, k2_bytes = int(r["kf_off"][-1]), int(r["k2_off"][-1])
    value = world * nbytes * args.steps / (ms / 1e3) / 1e6
    # ---- end to end through the public call (host buffers) -----------------------------------
    for _ in range(1):
        pipe.encode_host(h_in, off)
    env.barrier()
    t0 = time.perf_counter()
    single_steps = max(1, min(args.steps, 3))
    for _ in range(single_steps):
        res = pipe.encode_host(h_in, off)
    env.barrier()
    single_s = env.gmax(time.perf_counter() - t0)
    for res in pipe.encode_host_many((h_in, off) for _ in range(2)):
        pass
    env.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)
    d2h = 0
    for res in pipe.encode_host_many((h_in, off) for _ in range(e2e_steps)):
        d2h += res["d2h_bytes"]
    env.barrier()
    e2e_s = env.gmax(time.perf_counter() - t0)
    e2e = {"value": round(world * nbytes * e2e_steps / e2e_s / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": res["h2d_bytes"] * world,
           "d2h_bytes_per_step": d2h // e2e_steps * world, "steps": e2e_steps, "call": "BlockPipeline.encode_host_many (streaming, copies overlap the neighbouring step)",
           "single_batch_call": {"value": round(world * nbytes * single_steps / single_s / 1e6, 2), "unit": UNIT, "steps": single_steps,
                                 "call": "BlockPipeline.encode_host (copy in, encode, copy out; nothing overlapped across calls)"}}

    # ---- decode chain (device resident): KF payload -> Rice/gamma parse -> inverse MTF -> inverse BBWT --------------
    c = pipe.ctx
    r = pipe.encode_device(d_in, off)
    kf_dev, kf_off = r["kf_payload"], r["kf_off"]

    def decode_once():
        m_ = c.rice_kf_decode(kf_dev, kf_off, off)
        return c.bbwt_inverse(c.mtf_decode(m_, off), off)
    x = decode_once()
    roundtrip_ok = bool(torch.equal(x[:nbytes], d_in[:nbytes]))
    env.barrier()
    dsteps = max(1, min(args.steps, 3))
    pipe.profile_reset()
    pipe.profile(True)
    decode_once()
    dprof = pipe.profile_read()
    pipe.profile(False)
    dv0, dv1 = ev_pair(torch)
    env.barrier()
    dv0.record()
    for _ in range(dsteps):
        decode_once()
    dv1.record()
    env.barrier()
    dms = env.gmax(dv0.elapsed_time(dv1))
    peak, peak_src = peaks()
    c_kf = kf_bytes / nbytes
    decode = {"metric": "decompress_throughput_rice_mtf_bbwt", "value": round(world * nbytes * dsteps / (dms / 1e3) / 1e6, 2), "unit": UNIT,
              "ms_per_step": round(dms / dsteps, 3), "steps": dsteps, "roundtrip_bit_exact": roundtrip_ok,
              # whole chain against the minimum traffic of a fused decoder: read the payload, write the block ((c + 1) B/B)
              "roofline": {"bound": "hbm", "alg_bytes_per_input_byte": round(c_kf + 1, 3), "achieved": round((c_kf + 1) * nbytes / (dms / dsteps / 1e3) / 1e9, 1),
                           "peak": peak, "unit": "GB/s", "frac": round((c_kf + 1) * nbytes / (dms / dsteps / 1e3) / 1e9 / peak, 4),
                           "per_category_ms": {k: round(v["ms"], 3) for k, v in dprof.items() if v["launches"]}}}
    del x

    # ---- per-kernel event timing (one extra, untimed-for-value step) -----------------------------
    pipe.profile_reset()
    pipe.profile(True)
    r = pipe.encode_device(d_in, off)
    prof = pipe.profile_read()
    pipe.profile(False)
    rounds = pipe.ctx.counters()
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    dom_name, dom_v = dom
    ach = dom_v["alg_bytes"] / (dom_v["ms"] / 1e3) / 1e9 if dom_v["ms"] > 0 else 0.0
    step_ms_prof = sum(v["ms"] for v in prof.values())
    stages = {}
    P = 5
    Rp, Rc = rounds["rounds_plain"], rounds["rounds_cyclic"]
    sort_keys = ("boot_keys", "radix_hist", "radix_scan", "radix_scatter", "rerank", "apply_ranks", "gather", "plan", "build_tiles", "lyndon_scan", "bbwt_emit")
    sort_ms = sum(prof[k]["ms"] for k in sort_keys if k in prof)
    model_bytes = nbytes * ((Rc * (36 + 24 * P) + 6) + (Rp * (36 + 24 * P) + 8))
    mtf_ms = sum(prof[k]["ms"] for k in ("mtf_pre", "mtf_scan", "mtf_main") if k in prof)
    rice_ms = sum(prof[k]["ms"] for k in ("rice_cost", "rice_plan", "rice_pack", "zero_fill") if k in prof)

    def frac(b, t):
        return round(b / (t / 1e3) / 1e9 / peak, 4) if t > 0 else None
    stages["bbwt_sort"] = {"ms": round(sort_ms, 3), "rounds_plain": Rp, "rounds_cyclic": Rc,
                           "model_bytes_per_input_byte": (Rc + Rp) * (36 + 24 * P) + 14, "frac_of_peak_model": frac(model_bytes, sort_ms),
                           "records_sorted_per_input_byte": round(rounds["records_sorted"] / nbytes, 3)}
    stages["mtf"] = {"ms": round(mtf_ms, 3), "alg_bytes_per_input_byte": 2, "frac_of_peak": frac(2 * nbytes, mtf_ms)}
    c_ratio = (kf_bytes + k2_bytes) / nbytes
    # 2 + c per coder (cost pass + pack pass + payload); the KF and V22 cost passes share one read here, so the stage's
    # algorithmic bytes are (1 cost read + 2 pack reads) + payloads = 3 + c when fused, 4 + c as two separate coders — the larger
    # figure would flatter the fraction, the smaller is quoted
    stages["rice"] = {"ms": round(rice_ms, 3), "alg_bytes_per_input_byte": round(3 + c_ratio, 3), "frac_of_peak": frac((3 + c_ratio) * nbytes, rice_ms),
                      "alg_bytes_two_coders": round(4 + c_ratio, 3)}
    traffic, traffic_note = ncu_traffic(dom_name, args)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "traffic_note": traffic_note, "alg_bytes_per_launch": int(dom_v["alg_bytes"] / max(1, dom_v["launches"])),
                "peak_source": peak_src, "launches": dom_v["launches"], "share_of_step": round(dom_v["ms"] / step_ms_prof, 3),
                "per_category_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}, "stages": stages}

    # ---- parity in the run: GPU payloads of the sampled blocks vs the oracle's -----------------
    fields = {"value": round(value, 2), "ms_per_step": round(ms / args.steps, 3), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
              "roofline": roofline, "decode": decode,
              "config": {"workload": "cfg2: BBWT+MTF+Rice stage, 1 MiB blocks, 256 MiB S1 text-like corpus per GPU",
                         "corpus_mib_per_gpu": args.mib, "block_bytes": block, "blocks_per_gpu": int(nblocks),
                         "l2": "inputs (256 MiB) and scratch (GBs) are larger than L2, no flush needed",
                         "outputs": {"kf_payload_bytes": kf_bytes, "k2_payload_bytes": k2_bytes}}}
    finish = None
    if cpu_thread is not None:
        # GPU side of the comparison: BBWT, MTF, the KF payload + parameters and all five V22 payloads of the sampled blocks
        nchk = min(args.cpu_blocks, int(nblocks))
        hi = int(off[nchk])
        g_bbwt = r["bbwt"][:hi].cpu().numpy().tobytes()
        g_mtf = r["mtf"][:hi].cpu().numpy().tobytes()
        g_kf = r["kf_payload"][:int(r["kf_off"][nchk])].cpu().numpy().tobytes()
        g_kfoff = np.array(r["kf_off"][:nchk + 1])
        g_k2 = {}
        for fl in K2_FLAGS:
            p_, o_, sz_ = pipe.ctx.rice_k2_encode(r["mtf"], off[:nchk + 1], fl)
            g_k2[fl] = (p_[:int(o_[-1])].cpu().numpy().tobytes(), o_, sz_)
        cpu_thread.start()                                          # the CPU sample runs under the remaining GPU legs; compared at the end

        def finish():
            cpu_thread.join()
            cpu, outs = cpu_box["res"]
            bad = []
            for b in range(nchk):
                L, m, kf, k2s = outs[b]
                a, e = int(off[b]), int(off[b + 1])
                if g_bbwt[a:e] != L: bad.append((b, "bbwt"))
                if g_mtf[a:e] != m: bad.append((b, "mtf"))
                if g_kf[int(g_kfoff[b]):int(g_kfoff[b + 1])] != kf: bad.append((b, "kf_payload"))
                for i, fl in enumerate(K2_FLAGS):
                    pay, o_, sz_ = g_k2[fl]
                    if pay[int(o_[b]):int(o_[b + 1])] != k2s[i]: bad.append((b, "k2_payload_%d" % fl))
                    if int(sz_[b, i]) != len(k2s[i]): bad.append((b, "k2_size_%d" % fl))
            return cpu, {"blocks": nchk, "identical": not bad, "compared": "bbwt, mtf, KF payload, five V22 payloads and sizes per block vs oracle/kolm_oracle.cpp",
                         "mismatches": bad[:8]}
    del pipe, d_in, h_in, r
    torch.cuda.empty_cache()
    return fields, finish


def leg_cfg3(env, args):
    """BASELINE cfg 3: LZ77 hash match search + greedy parse on the S2 mix (gradient / sine / pattern / checker megabytes), 1 MiB
    blocks, the reference's two parameter sets and the 64 KiB window of BASELINE's figure."""
    import numpy as np
    torch = env.torch
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200.stages import Context
    from oracle import oracle as O
    n = args.cfg3_mib * MIB
    nb = n // MIB
    off = np.arange(nb + 1, dtype=np.int64) * MIB
    corpus = synth.s2_mixed(n)
    x = torch.from_numpy(corpus).cuda()
    c = Context(n, nb, env.local)
    out = {"workload": f"{args.cfg3_mib} MiB of the S2 mix, 1 MiB blocks, 1 GPU", "sets": {}}
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=4)
    for window, cap in ((255, 127), (4096, 0), (65536, 0)):
        futs = [pool.submit(O.lz77_encode_fast, corpus[b * MIB:(b + 1) * MIB].tobytes(), window, cap) for b in range(min(4, nb))]
        pay, po = c.lz77_encode(x, off, window, cap)
        e0, e1 = ev_pair(torch)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.leg_steps):
            pay, po = c.lz77_encode(x, off, window, cap)
        e1.record()
        torch.cuda.synchronize()
        enc_ms = e0.elapsed_time(e1) / args.leg_steps
        wc = 0 if cap else window
        y = c.lz77_decode(pay, po, off, wc)
        d0, d1 = ev_pair(torch)
        torch.cuda.synchronize()
        d0.record()
        for _ in range(args.leg_steps):
            y = c.lz77_decode(pay, po, off, wc)
        d1.record()
        torch.cuda.synchronize()
        dec_ms = d0.elapsed_time(d1) / args.leg_steps
        rt = bool(torch.equal(y[:n], x[:n]))
        hp = pay[:int(po[min(4, nb)])].cpu().numpy().tobytes()
        same = all(hp[int(po[b]):int(po[b + 1])] == f.result() for b, f in enumerate(futs))
        out["sets"][f"window{window}_cap{cap or 'none'}"] = {
            "encode_MBps": round(n / enc_ms / 1e3, 1), "decode_MBps": round(n / dec_ms / 1e3, 1), "payload_bytes": int(po[-1]),
            "roundtrip_bit_exact": rt, "oracle_blocks_identical": bool(same), "oracle_blocks": min(4, nb)}
    out["ok"] = all(v["roundtrip_bit_exact"] and v["oracle_blocks_identical"] for v in out["sets"].values())
    c.close()
    del x
    torch.cuda.empty_cache()
    return out


class S3Corpus:
    """BASELINE cfg 4 / 5 corpus: containers of `gib` GiB of the S3 mix; container k uses seeds (7 + k, 0xC0FFEE + k).  A rank
    generates a container the first time it needs it (before the timed region: see `prefetch`)."""

    def __init__(self, ncont, cont_bytes):
        self.sizes = [cont_bytes] * ncont
        self.cache = {}

    def get(self, k):
        if k not in self.cache:
            from kolmogorovlike_datacompressor_b200 import synth
            self.cache[k] = synth.s3_mix(self.sizes[k], seed=7 + k, text_seed=0xC0FFEE + k)
        return self.cache[k]

    def load(self, k, a, b):
        return self.get(k)[a:b]                                      # a view: the engine uploads straight from it


def _cfg4_pass(env, sizes, load, get, bs, what):
    """One compress + decompress of the corpus (sizes[k] bytes of container k from load(k, a, b) / get(k)) at block size bs, sharded by
    blocks over the ranks of the run; see leg_cfg4."""
    torch = env.torch
    from kolmogorovlike_datacompressor_b200 import dist as kd
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    from oracle import oracle as O
    ncont, cbytes = len(sizes), sizes[0]
    blocks = kd.corpus_blocks(sizes, bs)
    parts = kd.partition_blocks(kd._virtual_bounds(blocks), env.world)
    b0, b1 = parts[env.rank]
    for k in sorted({blocks[i][0] for i in range(b0, b1)}):          # generation is not part of the path: before the clock
        get(k)
    total = sum(sizes)
    # sampled oracle expectation (rank 0): 1 MiB blocks: the first eight blocks of container 0 = one of every S3 segment kind;
    # smaller blocks: one block out of each of the first eight MiB
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=8)
    step = max(1, MIB // bs)
    picks = [i * step + (5 if step > 1 else 0) for i in range(8) if i * step + 5 < len(blocks) and blocks[i * step + 5][0] == 0] if env.rank == 0 else []
    futs = [pool.submit(O.encode_block, O.PROFILE_KOLR, get(0)[blocks[i][1]:blocks[i][2]].tobytes(), None, True) for i in picks]
    # warm-up on a small corpus: contexts, the Re-Pair slab pool, NCCL connections
    kfirst = blocks[b0][0] if b1 > b0 else 0
    wconts = kd.compress_kolr_fixed_corpus([min(cbytes, 16 * env.world * MIB)], lambda k, a, b: get(kfirst)[a:b], bs)
    kd.decompress_kolr_corpus(wconts, gather=False)                  # ... and the rank 0 -> rank r connections of the payload scatter
    env.barrier()
    st = {}
    e0, e1 = ev_pair(torch)
    t0 = time.perf_counter()
    e0.record()
    conts = kd.compress_kolr_fixed_corpus(sizes, load, bs, stats=st)
    e1.record()
    env.barrier()
    wall = env.gmax(time.perf_counter() - t0)
    ev_s = env.gmax(e0.elapsed_time(e1) / 1e3)
    # the same keys on every rank (each gmax is a collective): ranks other than 0 have no D2H / assembly phase
    phases = {k: round(env.gmax(st.get(k, 0.0)), 4) for k in ("load_s", "encode_s", "table_allgather_s", "payload_exchange_s", "d2h_s", "assemble_s")}
    out = {"workload": f"{ncont} containers x {cbytes // MIB} MiB of the S3 mix, {what}, all ten candidates exact; one corpus "
                       f"sharded by blocks over {env.world} GPU(s) (strong scaling)", "n_gpus": env.world, "scaling": "strong", "block_bytes": bs,
           "corpus_bytes": total, "compress_MBps": round(total / wall / 1e6, 1), "compress_s": round(wall, 3), "compress_s_cuda_events_max_over_ranks": round(ev_s, 3),
           "compress_phases_s_max_over_ranks": phases, "nccl_exchanged_bytes": int(st.get("exchanged_bytes", 0)) if env.rank == 0 else None}
    # decompress: containers live on rank 0; payload spans are scattered over NCCL, every rank decodes its block range (sharded output)
    env.barrier()
    dst_ = {}
    t0 = time.perf_counter()
    pieces = kd.decompress_kolr_corpus(conts, gather=False, stats=dst_)
    env.barrier()
    dwall = env.gmax(time.perf_counter() - t0)
    ok = True
    for k, a, b, y in pieces:
        ref = torch.from_numpy(get(k)[a:b]).cuda()
        ok = ok and bool(torch.equal(ref, y[:b - a]))
        del ref
    del pieces
    out.update({"decompress_MBps": round(total / dwall / 1e6, 1), "decompress_s": round(dwall, 3), "decompress_output": "sharded (each rank keeps its block range on its GPU)",
                "decompress_phases_s_max_over_ranks": {k: round(env.gmax(dst_.get(k, 0.0)), 4) for k in ("toc_s", "h2d_s", "payload_scatter_s", "decode_s")},
                "roundtrip_bit_exact": env.gall(ok)})
    if env.rank == 0:
        sha = hashlib.sha256()
        for c_ in conts:
            sha.update(c_)
        out["container_bytes"] = sum(len(c_) for c_ in conts)
        out["containers_sha256"] = sha.hexdigest()
        want = KNOWN_SHA.get((ncont, cbytes, bs))
        out["sha_equals_single_gpu"] = (sha.hexdigest() == want) if want else None
        if want is None:
            out["sha_note"] = "no recorded single-GPU value for this corpus size"
        names, starts, plens, olens, _, _ = V._parse(conts[0])
        hist = {}
        for nme in names:
            hist[nme] = hist.get(nme, 0) + 1
        out["container0_methods"] = hist
        bad = []
        for i, f in zip(picks, futs):
            mid, payload, _sizes = f.result()
            if V.KOLR_NAMES[mid] != names[i] or conts[0][starts[i]:starts[i] + plens[i]] != payload:
                bad.append(i)
        out["oracle_blocks"] = len(picks)
        out["oracle_blocks_identical"] = not bad
        if bad:
            out["oracle_mismatch_blocks"] = bad
        out["limiting_phase"] = max(phases.items(), key=lambda kv: kv[1])[0] if phases else None
    out["ok"] = bool(out["roundtrip_bit_exact"]) and (env.rank != 0 or (out.get("oracle_blocks_identical", True) and out.get("sha_equals_single_gpu") is not False))
    return out


def leg_cfg4(env, args):
    """BASELINE cfg 4: the full kolm_final_researched_v2-2 pipeline (compress_blocks_fixed + decompress, all ten candidates exact, TOC
    on the host) on a corpus of 1 GiB S3 containers, block-sharded over the N GPUs of the run — at 1 MiB blocks and, on the same bytes
    cut into 256 MiB containers, at the block size compress_blocks_fixed defaults to (8 KiB, V22.py:2332).  Strong scaling: the
    corpus is the same at every N; the timed region holds each rank's load of its own bytes, the encode, the NCCL table all_gather and
    payload gather, rank 0's D2H and container assembly."""
    torch = env.torch
    from oracle import oracle as O
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=8)
    ncont, cbytes = args.cfg4_containers, args.cfg4_container_mib * MIB
    corp = S3Corpus(ncont, cbytes)
    out = _cfg4_pass(env, corp.sizes, corp.load, corp.get, MIB, "fixed 1 MiB blocks")
    if args.cfg4_default_block > 0:
        # the same corpus as containers of at most 256 MiB (a KOLR container holds at most 65 535 blocks)
        sub = min(cbytes, 256 * MIB)
        per = (cbytes + sub - 1) // sub
        sizes2 = [min(sub, cbytes - (j % per) * sub) for j in range(ncont * per)]
        get2 = lambda j: corp.get(j // per)[(j % per) * sub:(j % per) * sub + sizes2[j]]
        env.torch.cuda.empty_cache()
        small = _cfg4_pass(env, sizes2, lambda j, a, b: get2(j)[a:b], get2, args.cfg4_default_block,
                           "fixed %d-byte blocks (the default of compress_blocks_fixed)" % args.cfg4_default_block)
        out["default_block"] = small
        out["ok"] = out["ok"] and small["ok"]
    # KOLM (kolm_final.compress / decompress) on one GPU: N = 1 only
    if env.world == 1 and args.kolm_mib > 0:
        from kolmogorovlike_datacompressor_b200 import kolm_final as KF
        data = corp.get(0)[:args.kolm_mib * MIB].tobytes()
        KF.compress(data[:8 * MIB], MIB)
        t0 = time.perf_counter(); blob = KF.compress(data, MIB); torch.cuda.synchronize(); t1 = time.perf_counter()
        back = KF.decompress(blob); t2 = time.perf_counter()
        names, starts, plens, olens, _ = KF._parse(blob)
        # sampled oracle check on the first blocks (CDC boundaries come from the container itself; they are oracle-checked in tests/)
        kbad, p = [], 0
        ids = {"raw": 0, "kf_xor": 1, "kf_bbwt": 2, "kf_lz77": 3}
        kf_futs = []
        for i in range(min(6, len(names))):
            kf_futs.append(pool.submit(O.encode_block, O.PROFILE_KOLM, data[p:p + olens[i]], None, True)); p += olens[i]
        for i, f in enumerate(kf_futs):
            mid, payload, _ = f.result()
            if ids[names[i]] != mid or blob[starts[i]:starts[i] + plens[i]] != payload:
                kbad.append(i)
        out["kolm"] = {"workload": f"kolm_final.compress(target_block = 1 MiB) on {args.kolm_mib} MiB of the S3 mix, 1 GPU", "compress_MBps": round(len(data) / (t1 - t0) / 1e6, 1),
                       "decompress_MBps": round(len(data) / (t2 - t1) / 1e6, 1), "roundtrip_bit_exact": back == data, "container_bytes": len(blob),
                       "oracle_blocks": len(kf_futs), "oracle_blocks_identical": not kbad}
        out["ok"] = out["ok"] and back == data and not kbad
    return out


# sha256 of the concatenated containers of leg_cfg4 produced on ONE GPU, keyed by (containers, bytes per container, block bytes):
# recorded from single-GPU runs of this same code; the N-GPU runs must reproduce them byte for byte.
KNOWN_SHA = {(4, 1024 * MIB, MIB): "191da9d501138af8f8d8dfe557c959b7e8456eb42f3770621f73c18d69386ce0",
             (16, 256 * MIB, 8192): "8d8c5c53063db0f5a2c523aa59e2d0653ac8b784ce7d5b28edb424d5aed6dcdc"}


def leg_cfg5(env, args):
    """BASELINE cfg 5: block sizes 64 KiB .. 16 MiB on the S3 mix, per GPU (weak): the cfg-2 step (BBWT + MTF + KF pack + V22 costs and
    one pack) and the KF decode chain, with the round trip and a sampled oracle comparison of block 0 (sizes the oracle finishes in
    seconds)."""
    import numpy as np
    torch = env.torch
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200.pipeline import BlockPipeline
    n = args.cfg5_mib * MIB
    corpus = synth.s3_mix(n, seed=7 + env.rank, text_seed=0xC0FFEE + env.rank)
    d_in = torch.from_numpy(corpus).cuda()
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=4)
    out = {"workload": f"{args.cfg5_mib} MiB of the S3 mix per GPU, {env.world} GPU(s), weak", "sizes": {}}
    pipe = None
    for kib in (64, 256, 1024, 4096, 16384):
        bs = kib << 10
        if bs > n:
            continue
        nb = (n + bs - 1) // bs
        off = np.minimum(np.arange(nb + 1, dtype=np.int64) * bs, n)
        fut = pool.submit(oracle_stage, corpus[:bs].tobytes()) if (env.rank == 0 and bs <= (4 << 20)) else None
        if pipe is None or pipe.max_blocks < nb:
            pipe = None
            torch.cuda.empty_cache()
            pipe = BlockPipeline(n, nb, device=env.local)
        r = pipe.encode_device(d_in, off)
        env.barrier()
        e0, e1 = ev_pair(torch)
        e0.record()
        for _ in range(args.leg_steps):
            r = pipe.encode_device(d_in, off)
        e1.record()
        env.barrier()
        enc_ms = env.gmax(e0.elapsed_time(e1)) / args.leg_steps
        c = pipe.ctx
        rounds = c.counters()
        kf_dev, kf_off = r["kf_payload"], r["kf_off"]
        x = c.bbwt_inverse(c.mtf_decode(c.rice_kf_decode(kf_dev, kf_off, off), off), off)
        rt = bool(torch.equal(x[:n], d_in[:n]))
        env.barrier()
        d0, d1 = ev_pair(torch)
        d0.record()
        x = c.bbwt_inverse(c.mtf_decode(c.rice_kf_decode(kf_dev, kf_off, off), off), off)
        d1.record()
        env.barrier()
        dec_ms = env.gmax(d0.elapsed_time(d1))
        rec = {"blocks_per_gpu": int(nb), "encode_MBps": round(env.world * n / enc_ms / 1e3, 1), "decode_MBps": round(env.world * n / dec_ms / 1e3, 1),
               "rounds_cyclic": rounds["rounds_cyclic"], "rounds_plain": rounds["rounds_plain"], "roundtrip_bit_exact": env.gall(rt)}
        if fut is not None:
            L, m, kf, k2s = fut.result()
            g_kf = kf_dev[:int(kf_off[1])].cpu().numpy().tobytes()
            g_k2 = r["k2_payload"][:int(r["k2_off"][1])].cpu().numpy().tobytes()
            rec["oracle_block0_identical"] = bool(r["bbwt"][:bs].cpu().numpy().tobytes() == L and r["mtf"][:bs].cpu().numpy().tobytes() == m and g_kf == kf and g_k2 == k2s[0]
                                                  and [int(v) for v in r["k2_sizes"][0]] == [len(p) for p in k2s])
        out["sizes"][f"{kib}KiB"] = rec
        del x
    out["ok"] = all(v["roundtrip_bit_exact"] and v.get("oracle_block0_identical", True) for v in out["sizes"].values())
    del pipe, d_in
    torch.cuda.empty_cache()
    return out


def leg_default_blocks(env, args):
    """The reference's own default block sizes through the fused per-batch call (kolm_encode_blocks / kolm_decode_blocks behind the
    drop-ins): compress_blocks_fixed(data, 2048) (V22 CLI default, V22.py:2650) and kolm_final.compress(data, 8192) (KF.py:866) on the
    S3 mix, with the round trip and a sampled oracle comparison of whole containers (KOLM) / block selections (KOLR)."""
    torch = env.torch
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    from oracle import oracle as O
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=8)
    out = {"workload": "S3 mix at the reference's default block sizes, 1 GPU, host bytes in -> container bytes out"}
    n_r = min(args.default_mib, 120) * MIB                           # <= 65 535 blocks of 2 KiB per KOLR container
    data = synth.s3_mix(max(n_r, args.default_mib * MIB), seed=11, text_seed=0xC0FFEE + 11)
    d_r = data[:n_r].tobytes()
    # oracle expectation for sampled blocks (one 2 KiB block out of each of the eight segment kinds)
    picks = [k * MIB // 2048 + 37 for k in range(8) if (k + 1) * MIB <= n_r]
    futs = [pool.submit(O.encode_block, O.PROFILE_KOLR, d_r[i * 2048:(i + 1) * 2048]) for i in picks]
    V.compress_blocks_fixed(d_r[:4 * MIB], 2048)
    # first call at this size: the engine's contexts, scratch and pinned buffers grow inside it (reported as *_cold); the second is the steady state
    tc = time.perf_counter(); V.compress_blocks_fixed(d_r, 2048); torch.cuda.synchronize(); cold_r = time.perf_counter() - tc
    runs_r = []
    for _ in range(2):                                               # steady state: the faster of two calls (host-side noise is one-sided)
        t0 = time.perf_counter(); blob = V.compress_blocks_fixed(d_r, 2048); torch.cuda.synchronize(); runs_r.append(time.perf_counter() - t0)
    stopped = V._engine().ctx.encode_blocks_stats()["repair_stopped_early"]      # blocks whose Re-Pair rounds ended at the lower bound
    t1 = time.perf_counter(); back = V.decompress(blob); t2 = time.perf_counter()
    names, starts, plens, olens, _, _ = V._parse(blob)
    bad = [i for i, f in zip(picks, futs) if V.KOLR_NAMES[f.result()[0]] != names[i] or blob[starts[i]:starts[i] + plens[i]] != f.result()[1]]
    hist = {}
    for nme in names:
        hist[nme] = hist.get(nme, 0) + 1
    out["kolr_2KiB"] = {"bytes": n_r, "blocks": len(names), "compress_MBps": round(n_r / min(runs_r) / 1e6, 1), "compress_runs_MBps": [round(n_r / t / 1e6, 1) for t in runs_r],
                        "compress_cold_MBps": round(n_r / cold_r / 1e6, 1), "decompress_MBps": round(n_r / (t2 - t1) / 1e6, 1),
                        "roundtrip_bit_exact": back == d_r, "container_bytes": len(blob), "methods": hist, "oracle_blocks": len(picks), "oracle_blocks_identical": not bad,
                        "repair_stopped_early_blocks": int(stopped)}
    d_m = data[:args.default_mib * MIB].tobytes()
    small = d_m[:2 * MIB]
    fut_c = pool.submit(O.kf_compress, small, 8192)                  # the oracle's whole container for a 2 MiB prefix
    KF.compress(d_m[:4 * MIB], 8192)
    tc = time.perf_counter(); KF.compress(d_m, 8192); torch.cuda.synchronize(); cold_m = time.perf_counter() - tc
    runs_m = []
    for _ in range(2):
        t0 = time.perf_counter(); blob = KF.compress(d_m, 8192); torch.cuda.synchronize(); runs_m.append(time.perf_counter() - t0)
    t1 = time.perf_counter(); back = KF.decompress(blob); t2 = time.perf_counter()
    same = KF.compress(small, 8192) == fut_c.result()
    # KOLR again on an 8 MiB prefix: the fused call (Re-Pair stops early) against the stage-by-stage path (every candidate to the end)
    eng = V._engine()
    was, pre = eng.fused, d_r[:min(n_r, 8 * MIB)]
    try:
        part = V.compress_blocks_fixed(pre, 2048)
        eng.fused = False
        same_r = V.compress_blocks_fixed(pre, 2048) == part
    finally:
        eng.fused = was
    out["kolr_2KiB"]["container_8MiB_equals_full_repair_path"] = bool(same_r)
    out["kolm_8KiB"] = {"bytes": len(d_m), "blocks": int.from_bytes(blob[16:18], "little"), "compress_MBps": round(len(d_m) / min(runs_m) / 1e6, 1),
                        "compress_runs_MBps": [round(len(d_m) / t / 1e6, 1) for t in runs_m], "compress_cold_MBps": round(len(d_m) / cold_m / 1e6, 1),
                        "decompress_MBps": round(len(d_m) / (t2 - t1) / 1e6, 1), "roundtrip_bit_exact": back == d_m, "container_bytes": len(blob),
                        "oracle_container_2MiB_identical": bool(same)}
    out["ok"] = bool(out["kolr_2KiB"]["roundtrip_bit_exact"] and not bad and same_r and out["kolm_8KiB"]["roundtrip_bit_exact"] and same)
    return out


def run_ours(args):
    env = Env(args)
    fields, finish_parity = leg_cfg2(env, args)
    parity_fail = False
    legs = {}
    want = set(args.legs.split(",")) if args.legs != "all" else {"cfg3", "cfg4", "cfg5", "default"}
    if args.legs == "none":
        want = set()
    for name, fn, cond in (("cfg3_lz77", leg_cfg3, "cfg3" in want and env.world == 1),
                           ("default_block_sizes", leg_default_blocks, "default" in want and env.world == 1),
                           ("cfg5_block_sweep", leg_cfg5, "cfg5" in want),
                           ("cfg4_full_pipeline", leg_cfg4, "cfg4" in want)):
        if not cond:
            continue
        t0 = time.perf_counter()
        try:
            legs[name] = fn(env, args)
        except Exception as e:                                       # a leg that cannot run is reported, the headline stands
            import traceback
            legs[name] = {"ok": False, "error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}
            env.torch.cuda.empty_cache()
        legs[name]["leg_wall_s"] = round(time.perf_counter() - t0, 1)
    if env.rank == 0:
        line = {"metric": METRIC, "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic"}
        line.update(fields)
        if finish_parity is not None:
            line["cpu_baseline"], line["parity"] = finish_parity()
            parity_fail = not line["parity"]["identical"]
        if env.world == 1 and args.cpu_blocks > 0 and args.cpu_all_cores:
            try:
                line["cpu_baseline_all_cores"] = cpu_port_all_cores()
            except Exception as e:
                line["cpu_baseline_all_cores"] = {"error": str(e)}
        py = python_reference_record()
        if py and env.world == 1:
            line["python_reference"] = py
        line["legs"] = legs
        line["legs_ok"] = all(v.get("ok", False) for v in legs.values())
        print(json.dumps(line))
    parity_fail = not env.gall(not parity_fail)                      # every rank leaves with the same exit code
    env.dist.destroy_process_group()
    return 3 if parity_fail else 0


def ncu_traffic(category, args):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/<tag>_traffic.json, newest tag first; written by
    tools/make_profile_summary.py).  ncu captured only the largest launches of the category, so the figure is their mean; it is
    not measured by this run."""
    kname = {"rerank": "k_rerank", "radix_scatter": "k_radix_scatter", "gather": "k_gather"}.get(category)
    pdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    path = next((os.path.join(pdir, t + "_traffic.json") for t in ("r2z", "r2b", "r2a", "r1d", "r1c") if os.path.isfile(os.path.join(pdir, t + "_traffic.json"))), "")
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
    ap.add_argument("--mib", type=int, default=256, help="cfg 2 corpus MiB per GPU")
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--cpu-blocks", type=int, default=16, help="blocks in the bounded cpu_baseline / parity sample (0 = skip)")
    ap.add_argument("--cpu-all-cores", type=int, default=1, help="also time the oracle port on every host core (N = 1)")
    ap.add_argument("--legs", default="all", help="comma list of cfg3,cfg4,cfg5,default | all | none")
    ap.add_argument("--leg-steps", type=int, default=2, help="timed repetitions inside the cfg 3 / cfg 5 legs")
    ap.add_argument("--cfg3-mib", type=int, default=256)
    ap.add_argument("--cfg4-containers", type=int, default=4, help="containers of the sharded S3 corpus (BASELINE cfg 4: 4 x 1 GiB)")
    ap.add_argument("--cfg4-container-mib", type=int, default=1024)
    ap.add_argument("--cfg4-default-block", type=int, default=8192, help="second pass of the cfg-4 leg at this block size (0 = skip)")
    ap.add_argument("--kolm-mib", type=int, default=256, help="KOLM drop-in leg on this many MiB (N = 1; 0 = skip)")
    ap.add_argument("--cfg5-mib", type=int, default=512, help="cfg 5 corpus MiB per GPU")
    ap.add_argument("--default-mib", type=int, default=120, help="corpus MiB of the default-block-size leg (N = 1)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
