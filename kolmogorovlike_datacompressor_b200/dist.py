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
    ends = np.array([b for _, b in bounds], dtype=np.int64)
    total = int(ends[-1]) - int(bounds[0][0])
    cuts = [0]
    for r in range(1, world):
        target = bounds[0][0] + (total * r) // world
        k = int(np.searchsorted(ends, target, side="left"))
        # block k straddles the target: give it to the side that keeps the split closer to the target
        if k < n and (ends[k] - target) <= (target - (ends[k - 1] if k else bounds[0][0])):
            k += 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


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


def sharded_encode_area(data: bytes, bounds: Bounds, area_fn, group=None, dst: int = 0):
    """Array form of sharded_encode for whole containers.  area_fn(data, bounds_slice) -> (method ids, payload lengths, payload
    area); the area is a device tensor under NCCL (it goes GPU -> GPU into its final place in one buffer on `dst`, which then
    comes home in a single copy) and a numpy array / CPU tensor under gloo.  Returns (ids, lengths, area as uint8 numpy array)
    in block order on `dst`, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition_blocks(bounds, world)
    b0, b1 = parts[rank]
    dev = _dev(group)
    if b1 > b0:
        mids, lens, area = area_fn(data, bounds[b0:b1])
    else:
        mids, lens, area = np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8)
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
    if rank != dst:
        if nbytes[rank]:
            dist.send(area[:nbytes[rank]].contiguous(), dst=dst, group=group)
        return None
    total = sum(nbytes)
    buf = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    p = 0
    for r in range(world):
        if nbytes[r]:
            if r == dst:
                buf[p:p + nbytes[r]].copy_(area[:nbytes[r]])
            else:
                dist.recv(buf[p:p + nbytes[r]], src=r, group=group)
        p += nbytes[r]
    if buf.is_cuda:
        host = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        host.copy_(buf, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out_area = host[:total].numpy()
    else:
        out_area = buf[:total].numpy()
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
