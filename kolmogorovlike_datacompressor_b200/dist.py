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


def compress_kolm(data: bytes, target_block: int = 8192, group=None) -> Optional[bytes]:
    """kolm_final.compress with its blocks sharded over the group's GPUs; the container is returned on rank 0."""
    import struct
    from . import kolm_final as KF
    data = bytes(data)
    cuts = KF.cdc_fast_boundaries(data, target_block // 2, target_block, target_block * 2)
    enc = sharded_encode(data, cuts, lambda d, b: KF._engine().encode_kolm(d, b), group)
    if enc is None:
        return None
    out = bytearray(b"KOLM")
    out += struct.pack("<I", target_block & 0xFFFFFFFF) + struct.pack("<Q", len(data)) + struct.pack("<H", len(cuts) & 0xFFFF)
    for (a, b), (mid, payload) in zip(cuts, enc):
        out.append(mid & 0xFF)
        out += struct.pack("<II", (b - a) & 0xFFFFFFFF, len(payload) & 0xFFFFFFFF)
        out += payload
    return bytes(out)
