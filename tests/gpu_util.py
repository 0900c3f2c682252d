"""Helpers for the -m gpu parity tests (call through the C-ABI via the stage wrappers)."""
from __future__ import annotations

import numpy as np
import torch

from kolmogorovlike_datacompressor_b200 import stages

_ctx = {}


def ctx(max_bytes=1 << 26, max_blocks=1 << 14):
    key = (max_bytes, max_blocks)
    if key not in _ctx:
        _ctx.clear()
        torch.cuda.empty_cache()
        _ctx[key] = stages.Context(max_bytes, max_blocks)
    return _ctx[key]


def batch(blocks):
    off = np.zeros(len(blocks) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in blocks])
    data = b"".join(blocks)
    t = torch.frombuffer(bytearray(data or b"\0"), dtype=torch.uint8).cuda()
    return t, off


def unbatch(t, off):
    a = t.cpu().numpy().tobytes()
    return [a[off[i]:off[i + 1]] for i in range(len(off) - 1)]
