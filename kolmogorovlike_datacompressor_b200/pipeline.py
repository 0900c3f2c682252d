"""BlockPipeline — the per-block transform hot path on one GPU, end to end through the C-ABI.

    BBWT (Lyndon + rotation sort) -> MTF -> { KF model-2 token coder | V22 Rice(k=2) x 5 variants }

`encode_device` works on device-resident inputs/outputs (no copies); `encode_host` is the call a
drop-in user makes: host bytes in (pinned staging -> H2D), payload bytes out (D2H to pinned).
No CPU fallback: construction fails without a CUDA device / the native library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .stages import Context

K2_FLAGS = (0, 1, 4, 8, 16)   # V22 methods 2..6 (kolm_final_researched_v2-2.py:2156-2160)


class BlockPipeline:
    def __init__(self, max_batch_bytes: int, max_blocks: int, device: Optional[int] = None, profile_kf: bool = True,
                 profile_k2: bool = True):
        self.ctx = Context(max_batch_bytes, max_blocks, device)
        self.device = torch.device("cuda", self.ctx.device)
        self.cap = int(max_batch_bytes)
        self.max_blocks = int(max_blocks)
        self.profile_kf, self.profile_k2 = profile_kf, profile_k2
        import os
        self.wave_bytes = int(os.environ.get("KOLM_WAVE_MIB", "65536")) << 20
        with torch.cuda.device(self.device):
            self.d_in = torch.empty(self.cap + 64, dtype=torch.uint8, device=self.device)
            self.d_bbwt = torch.empty(self.cap + 64, dtype=torch.uint8, device=self.device)
            self.d_mtf = torch.empty(self.cap + 64, dtype=torch.uint8, device=self.device)
            # worst cases: KF <= ~2.2 bytes/byte (gamma of 255), K2 <= 8.25 bytes/byte + padded group
            self.d_kf = torch.empty(3 * self.cap + 16 * self.max_blocks + 64, dtype=torch.uint8, device=self.device) if profile_kf else None
            self.d_k2 = torch.empty(9 * self.cap + 16 * self.max_blocks + 64, dtype=torch.uint8, device=self.device) if profile_k2 else None
        self.h_in = torch.empty(self.cap, dtype=torch.uint8).pin_memory()
        self.h_out = None

    # ------------------------------------------------------------------
    def encode_device(self, x: torch.Tensor, off: Sequence[int], k2_flags: int = 0):
        """x: uint8 CUDA tensor with the blocks back to back.  Returns a dict of device tensors + host offsets."""
        c = self.ctx
        r = {}
        # optional waves of blocks (KOLM_WAVE_MIB); measured slower than one launch set over the whole batch, kept as a knob
        off = np.asarray(off, dtype=np.int64)
        nb = len(off) - 1
        b0 = 0
        while b0 < nb:
            b1 = b0 + 1
            while b1 < nb and off[b1 + 1] - off[b0] <= self.wave_bytes:
                b1 += 1
            c.bbwt_forward(x, off[b0:b1 + 1], out=self.d_bbwt)
            b0 = b1
        c.mtf_encode(self.d_bbwt, off, out=self.d_mtf)
        if self.profile_kf:
            r["kf_payload"], r["kf_off"], r["kf_params"] = c.rice_kf_encode(self.d_mtf, off, out=self.d_kf, want_params=True)
        if self.profile_k2:
            r["k2_payload"], r["k2_off"], r["k2_sizes"] = c.rice_k2_encode(self.d_mtf, off, k2_flags, out=self.d_k2)
        r["bbwt"], r["mtf"] = self.d_bbwt, self.d_mtf
        return r

    def encode_host(self, data, off: Sequence[int], k2_flags: int = 0):
        """data: bytes / numpy uint8 / CPU uint8 tensor.  H2D, encode, D2H of the payloads (all inside this call)."""
        n = int(off[-1])
        if isinstance(data, torch.Tensor):
            src = data
        else:
            src = torch.from_numpy(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        if not src.is_pinned():
            self.h_in[:n].copy_(src[:n])
            src = self.h_in
        with torch.cuda.device(self.device):
            self.d_in[:n].copy_(src[:n], non_blocking=True)
            r = self.encode_device(self.d_in, off, k2_flags)
            out = {}
            total = 0
            for key in ("kf", "k2"):
                if key + "_payload" in r:
                    m = int(r[key + "_off"][-1])
                    total += m
            if self.h_out is None or self.h_out.numel() < total:
                self.h_out = torch.empty(max(total, 1) + (total >> 2), dtype=torch.uint8).pin_memory()
            p = 0
            for key in ("kf", "k2"):
                if key + "_payload" in r:
                    m = int(r[key + "_off"][-1])
                    self.h_out[p:p + m].copy_(r[key + "_payload"][:m], non_blocking=True)
                    out[key] = (p, m)
                    p += m
            torch.cuda.current_stream().synchronize()
        res = {"h2d_bytes": n, "d2h_bytes": total}
        for key, (p0, m) in out.items():
            res[key + "_payload"] = self.h_out[p0:p0 + m]
            res[key + "_off"] = r[key + "_off"]
        if "kf_params" in r:
            res["kf_params"] = r["kf_params"]
        if "k2_sizes" in r:
            res["k2_sizes"] = r["k2_sizes"]
        return res

    # ------------------------------------------------------------------
    def profile(self, enable: bool):
        _lib.lib().kolm_profile_enable(self.ctx._h, 1 if enable else 0)

    def profile_reset(self):
        _lib.lib().kolm_profile_reset(self.ctx._h)

    def profile_read(self):
        L = _lib.lib()
        L.kolm_profile_name.restype = C.c_char_p
        n = L.kolm_profile_categories()
        ms = (C.c_double * n)()
        la = (C.c_int64 * n)()
        ab = (C.c_int64 * n)()
        _lib.check(L.kolm_profile_read(self.ctx._h, ms, la, ab))
        return {L.kolm_profile_name(i).decode(): dict(ms=ms[i], launches=la[i], alg_bytes=ab[i]) for i in range(n)}
