"""Block-sharded multi-GPU encode: blocks are fully independent, so contiguous block ranges balanced by bytes go
to the ranks of a torch.distributed group, each rank runs the per-block hot path on its own GPU, and one exchange gathers the
compressed stream on rank 0 (all_gather of the per-block table, then payload bytes by send/recv into their final order).
NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors for the host-logic tests (with an injected encoder).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

Bounds = Sequence[Tuple[int, int]]
Encoded = List[Tuple[int, bytes]]


def partition_blocks(bounds: Bounds, world: int) -> List[Tuple[int, int]]:
    """Contiguous block ranges [b_r, b_{r+1}) with (nearly) equal byte counts; ranks may get empty ranges."""
    n = len(bounds)
    if n == 0:
        return [(0, 0)] * world
    return _partition_ends(np.fromiter((b for _, b in bounds), dtype=np.int64, count=n), int(bounds[0][0]), world)


def _partition_ends(ends: np.ndarray, start0: int, world: int) -> List[Tuple[int, int]]:
    """partition_blocks on the blocks' end offsets alone (block i = [ends[i-1], ends[i]), the first one starts at start0)."""
    n = len(ends)
    if n == 0:
        return [(0, 0)] * world
    total = int(ends[-1]) - start0
    cuts = [0]
    for r in range(1, world):
        target = start0 + (total * r) // world
        k = int(np.searchsorted(ends, target, side="left"))
        # block k straddles the target: give it to the side that keeps the split closer to the target
        if k < n and (ends[k] - target) <= (target - (ends[k - 1] if k else start0)):
            k += 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _fixed_plan(sizes: Sequence[int], block_size: int, world: int, rank: int):
    """The block table of a corpus of fixed-size blocks by arithmetic — the same ranges partition_blocks(_virtual_bounds(
    corpus_blocks(sizes, block_size)), world) gives, without the per-block lists (a 4 GiB corpus has 524 288 blocks of 8 KiB):
    -> (parts [(first block, end block) per rank], total number of blocks, runs of `rank`: [(container, lo, hi, [(a - lo, b - lo)
    of its blocks])] with [lo, hi) the byte range of the container this rank encodes)."""
    bs = int(block_size)
    counts = [(int(n) + bs - 1) // bs for n in sizes]
    cum = np.concatenate(([0], np.cumsum(np.asarray(counts, dtype=np.int64)))).astype(np.int64)
    nblocks = int(cum[-1])
    ends = np.empty(nblocks, dtype=np.int64)
    base = 0
    for k, n in enumerate(sizes):
        c = counts[k]
        if c:
            ends[cum[k]:cum[k + 1]] = np.minimum(np.arange(1, c + 1, dtype=np.int64) * bs, int(n)) + base
        base += int(n)
    parts = _partition_ends(ends, 0, world)
    b0, b1 = parts[rank]
    runs = []
    for k, n in enumerate(sizes):
        i, j = max(b0, int(cum[k])), min(b1, int(cum[k + 1]))
        if i >= j:
            continue
        lo, hi = (i - int(cum[k])) * bs, min(int(n), (j - int(cum[k])) * bs)
        runs.append((k, lo, hi, [(a - lo, min(hi, a + bs) - lo) for a in range(lo, hi, bs)]))
    return parts, nblocks, runs


def _dev(group) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def sharded_encode(data: bytes, bounds: Bounds, encode_fn: Callable[[bytes, Bounds], Encoded], group=None, dst: int = 0) -> Optional[Encoded]:
    """Every rank holds `data` (or at least its own range) and calls this collectively.
    Returns the per-block (method id, payload) list in block order on rank `dst`, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition_blocks(bounds, world)
    b0, b1 = parts[rank]
    mine: Encoded = encode_fn(data, bounds[b0:b1]) if b1 > b0 else []
    dev = _dev(group)
    # per-block table: method id, payload length
    nloc = b1 - b0
    maxn = max(e - s for s, e in parts)
    table = torch.zeros((max(maxn, 1), 2), dtype=torch.int64, device=dev)
    if nloc:
        table[:nloc, 0] = torch.tensor([m for m, _ in mine], dtype=torch.int64)
        table[:nloc, 1] = torch.tensor([len(p) for _, p in mine], dtype=torch.int64)
    tables = [torch.empty_like(table) for _ in range(world)]
    dist.all_gather(tables, table, group=group)
    blob = b"".join(p for _, p in mine)
    if rank != dst:
        if blob:
            t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
            dist.send(t, dst=dst, group=group)
        return None
    out: Encoded = []
    for r in range(world):
        s, e = parts[r]
        tab = tables[r][:e - s].cpu().numpy()
        nbytes = int(tab[:, 1].sum()) if e > s else 0
        if r == dst:
            raw = blob
        elif nbytes:
            t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            dist.recv(t, src=r, group=group)
            raw = t.cpu().numpy().tobytes()
        else:
            raw = b""
        p = 0
        for k in range(e - s):
            ln = int(tab[k, 1])
            out.append((int(tab[k, 0]), raw[p:p + ln]))
            p += ln
    return out


_PIN = {}


def _pinned(n: int) -> torch.Tensor:
    """Persistent pinned staging buffer (allocating and pinning a GB of host memory costs about a second each time)."""
    t = _PIN.get("buf")
    if t is None or t.numel() < n:
        _PIN["buf"] = t = torch.empty(max(n + (n >> 3), 1 << 20), dtype=torch.uint8).pin_memory()
    return t


_STAGE: dict = {}


def _home_area(buf: torch.Tensor, total: int) -> np.ndarray:
    """Device bytes -> a fresh host array through two persistent pinned staging buffers of 64 MiB: the D2H copy of chunk k+1 runs
    while eight threads move chunk k into the result (whose pages are touched for the first time there).  Pinning a buffer of the
    whole compressed stream instead cost 0.9-1.4 s per GiB on the first call — more than the copy itself — and sat on rank 0's
    critical path of the sharded compress."""
    from .engine import _par_copy
    out = np.empty(total, dtype=np.uint8)
    if total == 0:
        return out
    ch = 64 << 20
    st = _STAGE.get("bufs")
    if st is None:
        st = _STAGE["bufs"] = [torch.empty(ch, dtype=torch.uint8).pin_memory() for _ in range(2)]
    stream = torch.cuda.current_stream(buf.device)
    prev = None
    for k, a in enumerate(range(0, total, ch)):
        n = min(ch, total - a)
        stg = st[k & 1]                                      # last used by chunk k-2, whose host copy finished in the previous iteration
        stg[:n].copy_(buf[a:a + n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        if prev is not None:
            pa, pn, ps, pev = prev
            pev.synchronize()
            _par_copy(out[pa:pa + pn], ps[:pn].numpy())
        prev = (a, n, stg, ev)
    pa, pn, ps, pev = prev
    pev.synchronize()
    _par_copy(out[pa:pa + pn], ps[:pn].numpy())
    return out


class _Clock:
    """Phase timer of the sharded calls: seconds per phase, taken after the device finished the phase's work."""

    def __init__(self, stats, dev):
        import time
        self.stats, self.dev, self.t = stats, dev, time.perf_counter()

    def lap(self, name):
        if self.stats is None:
            return
        import time
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        self.stats[name] = self.stats.get(name, 0.0) + (now - self.t)
        self.t = now


def _exchange(sends, recvs, group):
    """All payload transfers of one gather as ONE group of point-to-point operations (NCCL fuses them into a single launch and
    the receives need no order): sends = [(tensor, peer)], recvs = [(tensor, peer)]."""
    ops = [dist.P2POp(dist.isend, t, peer, group) for t, peer in sends] + [dist.P2POp(dist.irecv, t, peer, group) for t, peer in recvs]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def sharded_encode_area(data: bytes, bounds: Bounds, area_fn, group=None, dst: int = 0, data_base: int = 0, stats: Optional[dict] = None):
    """Array form of sharded_encode for whole containers.  area_fn(data, bounds_slice) -> (method ids, payload lengths, payload
    area); the area is a device tensor under NCCL (it goes GPU -> GPU into its final place in one buffer on `dst`, which then
    comes home in a single copy) and a numpy array / CPU tensor under gloo.  Returns (ids, lengths, area as uint8 numpy array)
    in block order on `dst`, None elsewhere.  `data` may be just this rank's slice of the input: data[0] is byte `data_base` of
    the input the (global) `bounds` refer to.  stats (dict) receives seconds per phase."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition_blocks(bounds, world)
    b0, b1 = parts[rank]
    dev = _dev(group)
    clk = _Clock(stats, dev)
    if b1 > b0:
        mine = bounds[b0:b1] if not data_base else [(a - data_base, b - data_base) for a, b in bounds[b0:b1]]
        mids, lens, area = area_fn(data, mine)
    else:
        mids, lens, area = np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8)
    clk.lap("encode_s")
    if not isinstance(area, torch.Tensor):
        area = torch.from_numpy(np.ascontiguousarray(np.frombuffer(memoryview(area), dtype=np.uint8)).copy())
    area = area.to(dev)
    nloc = b1 - b0
    maxn = max(e - s for s, e in parts)
    table = torch.zeros((max(maxn, 1), 2), dtype=torch.int64, device=dev)
    if nloc:
        table[:nloc, 0] = torch.as_tensor(np.asarray(mids, dtype=np.int64), device=dev)
        table[:nloc, 1] = torch.as_tensor(np.asarray(lens, dtype=np.int64), device=dev)
    tables = [torch.empty_like(table) for _ in range(world)]
    dist.all_gather(tables, table, group=group)
    tabs = [tables[r][:parts[r][1] - parts[r][0]].cpu().numpy() for r in range(world)]
    nbytes = [int(t[:, 1].sum()) if len(t) else 0 for t in tabs]
    clk.lap("table_allgather_s")
    if rank != dst:
        _exchange([(area[:nbytes[rank]].contiguous(), dst)] if nbytes[rank] else [], [], group)
        clk.lap("payload_exchange_s")
        return None
    total = sum(nbytes)
    buf = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    p, recvs = 0, []
    for r in range(world):
        if nbytes[r]:
            if r == dst:
                buf[p:p + nbytes[r]].copy_(area[:nbytes[r]])
            else:
                recvs.append((buf[p:p + nbytes[r]], r))
        p += nbytes[r]
    _exchange([], recvs, group)                               # every rank's area straight into its final place, one group
    clk.lap("payload_exchange_s")
    out_area = _home_area(buf, total) if buf.is_cuda else buf[:total].numpy()
    clk.lap("d2h_s")
    allm = np.concatenate([t[:, 0] for t in tabs]) if tabs else np.zeros(0, np.int64)
    alll = np.concatenate([t[:, 1] for t in tabs]) if tabs else np.zeros(0, np.int64)
    return allm, alll, out_area


def _engine_area(eng, fn):
    """Call eng.<encode_*_area> with the payload area left on the device (NCCL sends it from there)."""
    def run(d, b):
        eng._device_out = True
        try:
            return fn(d, b)
        finally:
            eng._device_out = False
    return run


def compress_kolm(data: bytes, target_block: int = 8192, group=None) -> Optional[bytes]:
    """kolm_final.compress with its blocks sharded over the group's GPUs; the container is returned on rank 0."""
    import struct
    from . import kolm_final as KF
    data = bytes(data)
    cuts = KF.cdc_fast_boundaries(data, target_block // 2, target_block, target_block * 2)
    head = b"KOLM" + struct.pack("<I", target_block & 0xFFFFFFFF) + struct.pack("<Q", len(data)) + struct.pack("<H", len(cuts) & 0xFFFF)
    if not cuts:
        return head if dist.get_rank(group) == 0 else None
    eng = KF._engine()
    enc = sharded_encode_area(data, cuts, _engine_area(eng, eng.encode_kolm_area), group)
    if enc is None:
        return None
    return KF._container(head, cuts, enc[0], enc[1], enc[2])


def _compress_kolr(data: bytes, bounds, mode: int, size_field: int, group=None) -> Optional[bytes]:
    from . import kolm_final_researched_v2_2 as V
    names = V._candidate_names()
    if not bounds:
        return V._assemble(data, bounds, mode, size_field) if dist.get_rank(group) == 0 else None
    eng = V._engine()
    enc = sharded_encode_area(data, bounds, _engine_area(eng, lambda d, b: eng.encode_kolr_area(d, b, names)), group)
    if enc is None:
        return None
    return V._assemble(data, bounds, mode, size_field, encoded=enc)


def compress_kolr_fixed(data: bytes, block_size: int = 8192, group=None) -> Optional[bytes]:
    """compress_blocks_fixed with its blocks sharded over the group's GPUs; the container (TOC built on rank 0) is returned there."""
    from . import kolm_final_researched_v2_2 as V
    data = bytes(data)
    return _compress_kolr(data, V.fixed_boundaries(data, block_size), V.MODE_FIXED, block_size, group)


def compress_kolr_cdc(data: bytes, min_size: int = 4096, avg_size: int = 8192, max_size: int = 16384, group=None) -> Optional[bytes]:
    """compress_blocks_cdc, sharded the same way."""
    from . import kolm_final_researched_v2_2 as V
    data = bytes(data)
    return _compress_kolr(data, V.cdc_fast_boundaries_strict(data, min_size, avg_size, max_size), V.MODE_CDC, avg_size, group)


def _sharded_decode(blob: bytes, names, starts, plens, olens, decode_fn, group=None, dst: int = 0, piece: int = 256 << 20) -> Optional[bytes]:
    """Blocks of one container decoded on the group's GPUs: contiguous block ranges balanced by decoded bytes; every rank holds
    the container, decodes its range with decode_fn(blob, names, starts, plens, olens) -> uint8 tensor of the range's bytes
    (device tensor under NCCL), and the ranges are collected in order on `dst` (send/recv in pieces of `piece` bytes)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ol = np.asarray(olens, dtype=np.int64)
    ends = np.cumsum(ol)
    bounds = [(int(e - l), int(e)) for e, l in zip(ends, ol)]
    parts = partition_blocks(bounds, world)
    b0, b1 = parts[rank]
    dev = _dev(group)
    mine = decode_fn(blob, names[b0:b1], starts[b0:b1], plens[b0:b1], olens[b0:b1]) if b1 > b0 else torch.empty(0, dtype=torch.uint8, device=dev)
    nbytes = [int(ol[s:e].sum()) for s, e in parts]
    if rank != dst:
        for a in range(0, nbytes[rank], piece):
            dist.send(mine[a:min(nbytes[rank], a + piece)].contiguous(), dst=dst, group=group)
        return None
    from .engine import _new_bytes, _par_copy
    total = int(ol.sum())
    out, sink = _new_bytes(total)                            # filled in place: no second copy of the decoded data
    if sink is None:
        sink = np.zeros(total, dtype=np.uint8)
    host = torch.empty(max(1, min(piece, max(nbytes))), dtype=torch.uint8)
    if dev.type == "cuda":
        host = host.pin_memory()
    p = 0
    for r in range(world):
        for a in range(0, nbytes[r], piece):
            n = min(nbytes[r], a + piece) - a
            if r == dst:
                t = mine[a:a + n]
            else:
                t = torch.empty(n, dtype=torch.uint8, device=dev)
                dist.recv(t, src=r, group=group)
            host[:n].copy_(t)
            _par_copy(sink[p:p + n], host[:n].numpy())
            p += n
    return out if total >= 2 else sink.tobytes()


def decompress_kolm(blob: bytes, group=None) -> Optional[bytes]:
    """kolm_final.decompress with the container's blocks decoded across the group's GPUs; the data is returned on rank 0."""
    from . import kolm_final as KF
    blob = bytes(blob)
    names, starts, plens, olens, total_len = KF._parse(blob)
    out = _sharded_decode(blob, names, starts, plens, olens, KF._engine().decode_to_device, group)
    if out is not None and len(out) != total_len:
        raise ValueError(f"Total decoded length mismatch: expected {total_len}, got {len(out)}")
    return out


def decompress_kolr(container: bytes, group=None) -> Optional[bytes]:
    """kolm_final_researched_v2_2.decompress, sharded the same way (every rank parses the TOC)."""
    from . import kolm_final_researched_v2_2 as V
    container = bytes(container)
    names, starts, plens, olens, total_len, pos = V._parse(container)
    out = _sharded_decode(container, names, starts, plens, olens, V._engine().decode_to_device, group)
    if out is not None:
        if len(out) != total_len:
            raise ValueError(f"Length mismatch: got {len(out)}, expect {total_len}")
        if pos != len(container):
            raise ValueError(f"Extra trailing {len(container) - pos} bytes after container end")
    return out


# ------------------------------------------------------------------------------------------------------------------------------
# Corpora of several containers ("multi-container corpora additionally shard by container"; a KOLR container holds
# < 4 GiB and <= 65535 blocks, V22.py:2336-2339).  The blocks of ALL containers form one list in container order; it is cut into
# contiguous ranges balanced by bytes, every rank loads only the bytes of its own blocks, and ONE exchange (table all_gather +
# one group of sends/receives) brings every container's payload area to the assembling rank.
# ------------------------------------------------------------------------------------------------------------------------------
def corpus_blocks(sizes: Sequence[int], block_size: int) -> List[Tuple[int, int, int]]:
    """[(container, start, end)] of the fixed-size blocks of every container (V22.py:314-320), container-major."""
    out = []
    for k, n in enumerate(sizes):                                    # per container in numpy: a 4 GiB corpus has 524 288 blocks of 8 KiB
        a = np.arange(0, n, block_size, dtype=np.int64)
        out.extend(zip([k] * len(a), a.tolist(), np.minimum(a + block_size, n).tolist()))
    return out


def _virtual_bounds(blocks):
    from itertools import accumulate
    lens = [b - a for _, a, b in blocks]
    ends = list(accumulate(lens))
    return list(zip([e - n for e, n in zip(ends, lens)], ends))


def _my_runs(blocks, b0, b1):
    """Blocks [b0, b1) grouped into runs of one container each: [(container, first block index, last + 1)]."""
    runs, i = [], b0
    while i < b1:
        j = i
        while j < b1 and blocks[j][0] == blocks[i][0]:
            j += 1
        runs.append((blocks[i][0], i, j))
        i = j
    return runs


def compress_kolr_fixed_corpus(sizes: Sequence[int], load: Callable[[int, int, int], bytes], block_size: int = 8192, group=None, dst: int = 0,
                               stats: Optional[dict] = None, area_fn=None) -> Optional[List[bytes]]:
    """compress_blocks_fixed (V22.py:2332-2445) of every container of a corpus, block-sharded over the group's GPUs.
    sizes[k] = input length of container k; load(k, a, b) -> the bytes [a, b) of container k's input (called on the rank that
    encodes them, once per container it touches).  area_fn(data, bounds) -> (ids, lengths, payload area) is the per-rank encoder
    (default: this GPU's engine with the reference's candidate list); the containers are returned as a list on `dst`, None
    elsewhere.  stats (dict) receives seconds per phase and the exchanged byte count."""
    from . import kolm_final_researched_v2_2 as V
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _dev(group)
    clk = _Clock(stats, dev)
    parts, nblocks_all, runs = _fixed_plan(sizes, block_size, world, rank)
    b0, b1 = parts[rank]
    if area_fn is None:
        eng, names = V._engine(), V._candidate_names()
        area_fn = _engine_area(eng, lambda d, b: eng.encode_kolr_area(d, b, names))
    mids_l, lens_l, areas = [], [], []
    for k, lo, hi, local in runs:
        d = load(k, lo, hi)
        clk.lap("load_s")
        m, l, ar = area_fn(d, local)
        if not isinstance(ar, torch.Tensor):
            ar = torch.from_numpy(np.ascontiguousarray(np.frombuffer(memoryview(ar), dtype=np.uint8)).copy())
        mids_l.append(np.asarray(m, dtype=np.int64)); lens_l.append(np.asarray(l, dtype=np.int64)); areas.append(ar.to(dev))
        clk.lap("encode_s")
    nloc = b1 - b0
    maxn = max(e - s for s, e in parts)
    table = torch.zeros((max(maxn, 1), 2), dtype=torch.int64, device=dev)
    if nloc:
        table[:nloc, 0] = torch.as_tensor(np.concatenate(mids_l), device=dev)
        table[:nloc, 1] = torch.as_tensor(np.concatenate(lens_l), device=dev)
    tables = [torch.empty_like(table) for _ in range(world)]
    dist.all_gather(tables, table, group=group)
    tabs = [tables[r][:parts[r][1] - parts[r][0]].cpu().numpy() for r in range(world)]
    nbytes = [int(t[:, 1].sum()) if len(t) else 0 for t in tabs]
    clk.lap("table_allgather_s")
    mine = torch.cat(areas) if len(areas) > 1 else (areas[0] if areas else torch.empty(0, dtype=torch.uint8, device=dev))
    if stats is not None:
        stats["exchanged_bytes"] = int(sum(nbytes) - nbytes[dst])
    if rank != dst:
        _exchange([(mine[:nbytes[rank]].contiguous(), dst)] if nbytes[rank] else [], [], group)
        clk.lap("payload_exchange_s")
        return None
    total = sum(nbytes)
    buf = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    p, recvs = 0, []
    for r in range(world):
        if nbytes[r]:
            if r == dst:
                buf[p:p + nbytes[r]].copy_(mine[:nbytes[r]])
            else:
                recvs.append((buf[p:p + nbytes[r]], r))
        p += nbytes[r]
    _exchange([], recvs, group)
    clk.lap("payload_exchange_s")
    area = _home_area(buf, total) if buf.is_cuda else buf[:total].numpy()
    clk.lap("d2h_s")
    allm = np.concatenate([t[:, 0] for t in tabs]) if nblocks_all else np.zeros(0, np.int64)
    alll = np.concatenate([t[:, 1] for t in tabs]) if nblocks_all else np.zeros(0, np.int64)
    ends = np.cumsum(alll)
    out, i = [], 0
    for k, n in enumerate(sizes):
        bounds_k = V._FixedBounds(int(n), block_size)                # container k's blocks, in corpus_blocks order, without the list
        j = i + len(bounds_k)
        a0 = int(ends[i] - alll[i]) if j > i else 0
        a1 = int(ends[j - 1]) if j > i else 0
        out.append(V._assemble(int(n), bounds_k, V.MODE_FIXED, block_size,
                               encoded=(allm[i:j], alll[i:j], area[a0:a1]) if j > i else None))
        i = j
    clk.lap("assemble_s")
    return out


def decompress_kolr_corpus(containers: Optional[Sequence[bytes]], group=None, src: int = 0, gather: bool = True, stats: Optional[dict] = None,
                           decode_fn=None):
    """decompress (V22.py:2451-2550) of every container of a corpus, block-sharded over the group's GPUs.  The containers live on
    rank `src` (None elsewhere): it walks the TOCs, broadcasts the block table, uploads the payload areas once and sends every
    rank the payload span of its block range (one group of sends/receives); the ranks decode their ranges.
    gather=True : the decoded bytes travel back to `src` (one group), which returns [bytes per container]; others return None.
    gather=False: every rank returns [(container, first byte, end byte, uint8 tensor of those decoded bytes)] for its own range
                  (the sharded output SURVEY §8e describes when no single-device result is needed).
    decode_fn(span tensor, names, starts, plens, olens) -> uint8 tensor (default: this GPU's engine)."""
    from . import kolm_final_researched_v2_2 as V
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _dev(group)
    clk = _Clock(stats, dev)
    box = [None]
    if rank == src:
        metas, p = [], 0
        for c in containers:
            c = bytes(c) if not isinstance(c, bytes) else c
            names, starts, plens, olens, total_len, pos = V._parse(c)
            if pos != len(c):
                raise ValueError(f"Extra trailing {len(c) - pos} bytes after container end")
            if sum(olens) != total_len:
                raise ValueError(f"Length mismatch: got {sum(olens)}, expect {total_len}")
            a0 = starts[0] if starts else pos
            metas.append(dict(names=names, rel=[x - a0 for x in starts], plens=plens, olens=olens, area0=a0, area_len=pos - a0, gbase=p))
            p += pos - a0
        box = [[{k: v for k, v in m.items() if k != "area0"} for m in metas]]
    dist.broadcast_object_list(box, src=src, group=group)
    table = box[0]
    clk.lap("toc_s")
    # flat, container-major block table as arrays (a 4 GiB corpus at 8 KiB blocks has 524 288 rows: no per-block Python)
    counts = [len(m["olens"]) for m in table]
    cum = np.concatenate(([0], np.cumsum(np.asarray(counts, dtype=np.int64)))).astype(np.int64)
    nall = int(cum[-1])

    def _flat(key, add=None):
        if not nall:
            return np.zeros(0, dtype=np.int64)
        return np.concatenate([np.asarray(m[key], dtype=np.int64) + (m[add] if add else 0) for m in table])
    olen, plen, gstart = _flat("olens"), _flat("plens"), _flat("rel", "gbase")
    ostart = np.concatenate([np.cumsum(np.asarray(m["olens"], dtype=np.int64)) - np.asarray(m["olens"], dtype=np.int64) for m in table]) if nall else olen
    names: List[str] = []
    for m in table:
        names.extend(m["names"])
    parts = _partition_ends(np.cumsum(olen), 0, world)
    spans = []
    for s0, s1 in parts:                                             # payload span (in the concatenated areas) of every rank's range
        if s1 > s0:
            spans.append((int(gstart[s0:s1].min()), int((gstart[s0:s1] + plen[s0:s1]).max())))
        else:
            spans.append((0, 0))
    b0, b1 = parts[rank]
    lo, hi = spans[rank]
    if rank == src:
        total_area = sum(m["area_len"] for m in table)
        allp = torch.empty(max(total_area, 1), dtype=torch.uint8, device=dev)
        for c, m in zip(containers, metas):
            if m["area_len"]:
                view = np.frombuffer(c, dtype=np.uint8, count=m["area_len"], offset=m["area0"])
                from .engine import _ro_tensor
                allp[m["gbase"]:m["gbase"] + m["area_len"]].copy_(_ro_tensor(view))
        clk.lap("h2d_s")
        _exchange([(allp[spans[r][0]:spans[r][1]], r) for r in range(world) if r != src and spans[r][1] > spans[r][0]], [], group)
        span = allp[lo:hi]
    else:
        span = torch.empty(max(hi - lo, 1), dtype=torch.uint8, device=dev)
        _exchange([], [(span[:hi - lo], src)] if hi > lo else [], group)
    clk.lap("payload_scatter_s")
    if decode_fn is None:
        decode_fn = V._engine().decode_to_device
    pieces = []
    for k in range(len(table)):                                      # my blocks, one run per container
        i, j = max(b0, int(cum[k])), min(b1, int(cum[k + 1]))
        if i >= j:
            continue
        y = decode_fn(span, names[i:j], (gstart[i:j] - lo).tolist(), plen[i:j].tolist(), olen[i:j].tolist())
        pieces.append((k, int(ostart[i]), int(ostart[j - 1] + olen[j - 1]), y))
    clk.lap("decode_s")
    if not gather:
        return pieces
    nbytes = [int(olen[s0:s1].sum()) for s0, s1 in parts]
    mine = torch.cat([y for *_, y in pieces]) if len(pieces) > 1 else (pieces[0][3] if pieces else torch.empty(0, dtype=torch.uint8, device=dev))
    if rank != src:
        _exchange([(mine[:nbytes[rank]].contiguous(), src)] if nbytes[rank] else [], [], group)
        clk.lap("output_gather_s")
        return None
    total = sum(nbytes)
    buf = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    p, recvs = 0, []
    for r in range(world):
        if nbytes[r]:
            if r == src:
                buf[p:p + nbytes[r]].copy_(mine[:nbytes[r]])
            else:
                recvs.append((buf[p:p + nbytes[r]], r))
        p += nbytes[r]
    _exchange([], recvs, group)
    clk.lap("output_gather_s")
    from .engine import _new_bytes, _par_copy
    outs, p = [], 0
    host = None
    for m in table:
        n = sum(m["olens"])
        ob, sink = _new_bytes(n)
        if n:
            if buf.is_cuda:
                if host is None or host.numel() < n:
                    host = torch.empty(n, dtype=torch.uint8).pin_memory()
                host[:n].copy_(buf[p:p + n], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                srcv = host[:n].numpy()
            else:
                srcv = buf[p:p + n].numpy()
            if sink is None:
                ob = srcv.tobytes()
            else:
                _par_copy(sink, srcv)
        outs.append(ob)
        p += n
    clk.lap("d2h_s")
    return outs
