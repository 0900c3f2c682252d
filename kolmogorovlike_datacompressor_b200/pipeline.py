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
        self.h_kf = self.h_k2 = None
        with torch.cuda.device(self.device):
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()

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
        if self.profile_kf and self.profile_k2:                     # one cost read for both references' coders
            (r["kf_payload"], r["kf_off"], r["kf_params"], r["k2_payload"], r["k2_off"], r["k2_sizes"]) = c.rice_dual_encode(
                self.d_mtf, off, k2_flags, kf_out=self.d_kf, k2_out=self.d_k2)
        elif self.profile_kf:
            r["kf_payload"], r["kf_off"], r["kf_params"] = c.rice_kf_encode(self.d_mtf, off, out=self.d_kf, want_params=True)
        elif self.profile_k2:
            r["k2_payload"], r["k2_off"], r["k2_sizes"] = c.rice_k2_encode(self.d_mtf, off, k2_flags, out=self.d_k2)
        r["bbwt"], r["mtf"] = self.d_bbwt, self.d_mtf
        return r

    def encode_host(self, data, off: Sequence[int], k2_flags: int = 0, chunks: Optional[int] = None):
        """data: bytes / numpy uint8 / CPU uint8 tensor (pinned tensors are used in place).  Host in, host out: the batch is
        cut into `chunks` consecutive groups of blocks; the H2D copy of group i+1 and the D2H copy of group i-1's payloads run
        on their own streams while group i is being encoded."""
        off = np.asarray(off, dtype=np.int64)
        n, nb = int(off[-1]), len(off) - 1
        if isinstance(data, torch.Tensor):
            src = data
        else:
            arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
            src = torch.from_numpy(arr if arr.flags.writeable else arr.copy())
        if not src.is_pinned():
            self.h_in[:n].copy_(src[:n])
            src = self.h_in
        if chunks is None:
            chunks = 1          # measured on B200: splitting the batch costs as much as the overlap hides (tools/e2e_chunks.py)
        chunks = max(1, min(chunks, nb))
        cuts = [int(round(i * nb / chunks)) for i in range(chunks + 1)]
        if self.h_kf is None and self.profile_kf:
            self.h_kf = torch.empty(max(n, 1 << 16), dtype=torch.uint8).pin_memory()
        if self.h_k2 is None and self.profile_k2:
            self.h_k2 = torch.empty(max(n, 1 << 16), dtype=torch.uint8).pin_memory()
        kf_off = np.zeros(nb + 1, dtype=np.int64)
        k2_off = np.zeros(nb + 1, dtype=np.int64)
        kf_params, k2_sizes = [], []
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            ev_in = []
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_stream(main)
                for ci in range(chunks):
                    a, b = int(off[cuts[ci]]), int(off[cuts[ci + 1]])
                    self.d_in[a:b].copy_(src[a:b], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(self.s_in)
                    ev_in.append(e)
            kpos = k2pos = 0
            for ci in range(chunks):
                b0, b1 = cuts[ci], cuts[ci + 1]
                sub = off[b0:b1 + 1]
                main.wait_event(ev_in[ci])
                c = self.ctx
                c.bbwt_forward(self.d_in, sub, out=self.d_bbwt)
                c.mtf_encode(self.d_bbwt, sub, out=self.d_mtf)
                done = []

                def ship():
                    # payloads go home on the output stream while the next kernels run
                    with torch.cuda.stream(self.s_out):
                        self.s_out.wait_stream(main)
                        for hbuf, hpos, p_, m_ in done:
                            hbuf[hpos:hpos + m_].copy_(p_[:m_], non_blocking=True)
                    done.clear()
                if self.profile_kf:
                    base = (3 * int(sub[0]) + 16 * b0 + 15) & ~15
                    p, o, prm = c.rice_kf_encode(self.d_mtf, sub, out=self.d_kf[base:], want_params=True)
                    m = int(o[-1])
                    kf_off[b0:b1 + 1] = kpos + o
                    kf_params.append(prm)
                    self.h_kf = self._grow(self.h_kf, kpos, m)
                    done.append((self.h_kf, kpos, p, m))
                    kpos += m
                    ship()
                if self.profile_k2:
                    base = (9 * int(sub[0]) + 16 * b0 + 15) & ~15
                    p, o, sz = c.rice_k2_encode(self.d_mtf, sub, k2_flags, out=self.d_k2[base:])
                    m = int(o[-1])
                    k2_off[b0:b1 + 1] = k2pos + o
                    k2_sizes.append(sz)
                    self.h_k2 = self._grow(self.h_k2, k2pos, m)
                    done.append((self.h_k2, k2pos, p, m))
                    k2pos += m
                ship()
            self.s_out.synchronize()
            main.synchronize()
        res = {"h2d_bytes": n, "d2h_bytes": kpos + k2pos, "chunks": chunks}
        if self.profile_kf:
            res["kf_payload"], res["kf_off"], res["kf_params"] = self.h_kf[:kpos], kf_off, np.concatenate(kf_params)
        if self.profile_k2:
            res["k2_payload"], res["k2_off"], res["k2_sizes"] = self.h_k2[:k2pos], k2_off, np.concatenate(k2_sizes)
        return res

    # ------------------------------------------------------------------
    def encode_host_many(self, batches, k2_flags: int = 0):
        """Streaming form of `encode_host` for corpora larger than one batch: `batches` yields (data, off) pairs, the results
        come back in order (same dict as `encode_host`).  Two device input buffers and two pinned payload buffers alternate, so
        the H2D copy of batch i+1 and the D2H copy of batch i's payloads run on their own streams underneath the kernels of the
        neighbouring batch; every batch is still copied in and its payloads copied out.  A result (its tensors alias the pinned
        buffers) must be consumed before the generator is advanced."""
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        if not hasattr(self, "_slots"):
            with torch.cuda.device(self.device):
                self._slots = [dict(d_in=self.d_in, h_in=self.h_in, h_kf=None, h_k2=None),
                               dict(d_in=torch.empty(self.cap + 64, dtype=torch.uint8, device=self.device), h_in=None, h_kf=None, h_k2=None)]
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()

            def h2d(batch, slot):
                data, off = batch
                off = np.asarray(off, dtype=np.int64)
                n = int(off[-1])
                if isinstance(data, torch.Tensor):
                    src = data
                else:
                    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
                    src = torch.from_numpy(arr if arr.flags.writeable else arr.copy())
                sl = self._slots[slot]
                if not src.is_pinned():
                    if sl["h_in"] is None:
                        sl["h_in"] = torch.empty(self.cap, dtype=torch.uint8).pin_memory()
                    sl["h_in"][:n].copy_(src[:n])
                    src = sl["h_in"]
                with torch.cuda.stream(self.s_in):
                    self.s_in.wait_stream(main)             # the previous user of this input buffer has finished
                    sl["d_in"][:n].copy_(src[:n], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.s_in)
                return off, n, ev

            def run(slot, off, n, ev):
                sl, c = self._slots[slot], self.ctx
                nb = len(off) - 1
                main.wait_event(ev)
                c.bbwt_forward(sl["d_in"], off, out=self.d_bbwt)
                c.mtf_encode(self.d_bbwt, off, out=self.d_mtf)
                main.wait_stream(self.s_out)                # the previous batch's payloads have left d_kf / d_k2
                res = {"h2d_bytes": n, "d2h_bytes": 0, "chunks": 1}
                for on, key, enc in ((self.profile_kf, "kf", lambda: c.rice_kf_encode(self.d_mtf, off, out=self.d_kf, want_params=True)),
                                     (self.profile_k2, "k2", lambda: c.rice_k2_encode(self.d_mtf, off, k2_flags, out=self.d_k2))):
                    if not on:
                        continue
                    pay, o, extra = enc()
                    m = int(o[-1])
                    hk = "h_" + key
                    if sl[hk] is None or sl[hk].numel() < m:
                        sl[hk] = torch.empty(max(m + (m >> 3), n, 1 << 16), dtype=torch.uint8).pin_memory()
                    with torch.cuda.stream(self.s_out):
                        self.s_out.wait_stream(main)
                        sl[hk][:m].copy_(pay[:m], non_blocking=True)
                    res[key + "_payload"], res[key + "_off"] = sl[hk][:m], o
                    res["kf_params" if key == "kf" else "k2_sizes"] = extra
                    res["d2h_bytes"] += m
                done = torch.cuda.Event()
                done.record(self.s_out)
                return res, done

            slot = 0
            off, n, ev = h2d(cur, slot)
            pending = None
            while cur is not None:
                nxt = next(it, None)
                if nxt is not None:
                    nxt_args = h2d(nxt, slot ^ 1)           # travels while this batch is being encoded
                out = run(slot, off, n, ev)
                if pending is not None:
                    pending[1].synchronize()
                    yield pending[0]
                pending, cur, slot = out, nxt, slot ^ 1
                if nxt is not None:
                    off, n, ev = nxt_args
            pending[1].synchronize()
            yield pending[0]

    def _grow(self, hbuf: torch.Tensor, used: int, more: int) -> torch.Tensor:
        """Pinned host buffer with room for used+more bytes (payloads usually shrink; incompressible data can expand)."""
        if used + more <= hbuf.numel():
            return hbuf
        self.s_out.synchronize()
        nb = torch.empty(max(used + more, 2 * hbuf.numel()), dtype=torch.uint8).pin_memory()
        nb[:used].copy_(hbuf[:used])
        return nb

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
