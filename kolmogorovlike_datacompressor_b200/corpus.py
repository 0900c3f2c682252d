"""Corpora larger than one container.

Both container formats cap what one container can hold: `KOLR` packs total_len as u32 and nblocks as u16 and
`struct.pack` raises beyond 4 GiB-1 / 65 535 blocks (kolm_final_researched_v2-2.py:2221-2222, 2338-2339); `KOLM` stores a u64
total but its block count wraps silently at 65 536 and the per-block lengths are u32 (kolm_final.py:889-890).  The BASELINE 4 GiB
and 16 GiB corpora are therefore sequences of containers.  This module only cuts the input into spans that respect the limits
and calls the drop-in modules span by span — each container is exactly what the reference produces for that span.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

MAX_BYTES = 0xFFFFFFFF          # u32 total_len (KOLR); also keeps every KOLM orig_len / payload_len inside u32
MAX_BLOCKS = 0xFFFF             # u16 nblocks


def container_spans(total_len: int, min_block: int, max_bytes: int = MAX_BYTES, max_blocks: int = MAX_BLOCKS) -> List[Tuple[int, int]]:
    """[a, b) spans covering [0, total_len): each at most max_bytes long and at most max_blocks blocks even if every block has
    the smallest size the chunker may emit (`min_block`: the fixed block size, or the CDC minimum).  Spans end on multiples of
    min_block so that fixed-size blocking of a span equals fixed-size blocking of the whole input."""
    if total_len < 0 or min_block < 1:
        raise ValueError("bad span parameters")
    cap = min(max_bytes, max_blocks * min_block)
    cap -= cap % min_block
    if cap < min_block:
        raise ValueError("block size exceeds the container limits")
    return [(a, min(total_len, a + cap)) for a in range(0, total_len, cap)] or [(0, 0)]


def compress_corpus(data: bytes, block_size: int = 1 << 20, profile: str = "kolr", cdc: bool = False,
                    max_bytes: int = MAX_BYTES, max_blocks: int = MAX_BLOCKS) -> List[bytes]:
    """One container per span.  profile 'kolr': compress_blocks_fixed(span, block_size) or, with cdc=True,
    compress_blocks_cdc(span, block_size // 2, block_size, 2 * block_size); profile 'kolm': kolm_final.compress(span, block_size)
    (always content-defined, blocks of block_size/2 .. 2*block_size)."""
    view = memoryview(data)
    if profile == "kolr":
        from . import kolm_final_researched_v2_2 as V
        spans = container_spans(len(data), block_size // 2 if cdc else block_size, max_bytes, max_blocks)
        if cdc:
            return [V.compress_blocks_cdc(bytes(view[a:b]), block_size // 2, block_size, 2 * block_size) for a, b in spans]
        return [V.compress_blocks_fixed(bytes(view[a:b]), block_size) for a, b in spans]
    if profile == "kolm":
        from . import kolm_final as K
        return [K.compress(bytes(view[a:b]), block_size) for a, b in container_spans(len(data), max(1, block_size // 2), max_bytes, max_blocks)]
    raise ValueError("profile must be 'kolr' or 'kolm'")


def decompress_corpus(containers: Iterable[bytes], profile: str = "kolr") -> bytes:
    if profile == "kolr":
        from . import kolm_final_researched_v2_2 as M
    elif profile == "kolm":
        from . import kolm_final as M
    else:
        raise ValueError("profile must be 'kolr' or 'kolm'")
    return b"".join(M.decompress(c) for c in containers)
