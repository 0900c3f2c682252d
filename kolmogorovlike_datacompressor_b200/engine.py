"""Host-side engine shared by the two drop-in modules: batching of blocks, per-block candidate evaluation on
the GPU (through the C-ABI stage operators), exact model selection, payload collection.

Mirrors the reference's per-block loops:
    KOLM  _encode_block            kolm_final.py:821-864            (ids 0..3, smallest payload, lowest id on ties)
    KOLR  selection loops          kolm_final_researched_v2-2.py:2233-2252 / 2350-2369  (ids = list index, strict '<')
There is no CPU fallback: everything below needs the CUDA library and a device.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Dict, List, Optional, Sequence, Tuple

import os

import numpy as np
import torch

from . import _lib
from .stages import Context

KOLR_NAMES = ["raw", "xor", "bbwt", "bbwt_bp", "bbwt_nib", "bbwt_br", "bbwt_gray", "lz77", "lfsr_pred", "repair", "v2_new"]
K2_FLAG_OF = {"bbwt": 0, "bbwt_bp": 1, "bbwt_nib": 4, "bbwt_br": 8, "bbwt_gray": 16}
K2_SLOT = {0: 0, 1: 1, 4: 2, 8: 3, 16: 4}
_BIG = np.iinfo(np.int64).max
_KOLR_IDS = {n: i for i, n in enumerate(KOLR_NAMES)}
_KOLM_IDS = {"raw": 0, "kf_xor": 1, "kf_bbwt": 2, "kf_lz77": 3}


def raise_like_reference(e: _lib.KolmError):
    """Map C-ABI error codes onto the exception types the reference raises."""
    if e.code == -4:
        raise EOFError(str(e)) from e
    if e.code == -7:
        raise IndexError(str(e)) from e
    if e.code in (-5, -2):
        raise ValueError(str(e)) from e
    raise e


_COPY_POOL = None


def _par_copy(dst: np.ndarray, src: np.ndarray) -> None:
    """dst[:] = src with several threads for large buffers: a fresh result buffer is all page faults on first touch (~2 GB/s
    from one thread), numpy releases the GIL inside the copy."""
    n = int(dst.shape[0])
    if n < (16 << 20):
        dst[:] = src
        return
    global _COPY_POOL
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=8, thread_name_prefix="kolm-copy")
    step = max(4 << 20, (n + 7) // 8)
    futs = [_COPY_POOL.submit(np.copyto, dst[a:min(n, a + step)], src[a:min(n, a + step)]) for a in range(0, n, step)]
    for f in futs:
        f.result()


def _ro_tensor(arr: np.ndarray) -> torch.Tensor:
    """CPU uint8 tensor sharing the memory of a read-only numpy view (the caller only reads it).  torch.from_numpy warns on
    read-only arrays; the array interface of a ctypes view of the same address does not, and no filter state is touched."""
    import ctypes
    n = int(arr.shape[0])
    if n == 0:
        return torch.empty(0, dtype=torch.uint8)
    addr = arr.__array_interface__["data"][0]
    view = np.ctypeslib.as_array((ctypes.c_ubyte * n).from_address(addr))
    t = torch.from_numpy(view)
    t._kolm_keepalive = arr                                     # the bytes object must outlive the tensor
    return t


def _new_bytes(n: int):
    """(bytes object of n bytes, writable uint8 view of its storage).  The object is filled in place before anyone else can see
    it (CPython: a fresh `bytes(n)` is not shared, interned or hashed), which saves one full copy of every decompressed output."""
    out = bytes(n)
    if n < 2:
        return out, None
    import ctypes
    addr = ctypes.cast(ctypes.c_char_p(out), ctypes.c_void_p).value
    return out, np.ctypeslib.as_array((ctypes.c_ubyte * n).from_address(addr))


def cdc_boundaries(which: str, data: bytes, mn: int, avg: int, mx: int) -> List[Tuple[int, int]]:
    n = len(data)
    if n == 0:
        return []
    L = _lib.lib()
    cap = n // max(1, mn) + 4
    ends = np.zeros(cap, dtype=np.int64)
    buf = C.cast(C.c_char_p(data), C.c_void_p)              # bytes are immutable and contiguous: no copy
    k = getattr(L, "kolm_cdc_" + which)(buf, n, mn, avg, mx, ends.ctypes.data_as(C.POINTER(C.c_int64)), cap)
    if k < 0:
        if k == -2:
            raise ValueError("Require 0 < min_size <= avg_size <= max_size and avg_size >= 64")
        raise _lib.KolmError(int(k))
    out, a = [], 0
    for e in ends[:k]:
        out.append((a, int(e)))
        a = int(e)
    return out


class Engine:
    """One GPU, one context that grows on demand."""

    _shared: Dict[int, "Engine"] = {}

    def __init__(self, device: Optional[int] = None, batch_bytes: int = 128 << 20):
        if not torch.cuda.is_available():
            raise RuntimeError("kolmogorovlike_datacompressor_b200: no CUDA device — the GPU path has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.batch_bytes = int(batch_bytes)
        self.ctx: Optional[Context] = None
        self.cap_bytes = 0
        self.cap_blocks = 0
        # Re-Pair candidate: blocks up to kolm_repair_max_block() bytes run in shared memory, longer ones through the incremental
        # kernel (exact, but a serial chain of ~10^5 rounds per MiB: tens of MB/s).  KOLM_REPAIR_MAX_BLOCK / this attribute caps
        # the block length for which the candidate is evaluated; capped blocks are skipped with a RuntimeWarning and the container
        # can then differ from the reference's wherever Re-Pair would have won.  Default: no cap (byte-exact).
        env = os.environ.get("KOLM_REPAIR_MAX_BLOCK")
        self.repair_max = int(env) if env else (1 << 30)
        self.enable_v2_new = False     # method 10 as an encode candidate (dead in the shipped reference; see kolm_final_researched_v2_2.G_ENABLE_V2_NEW)
        self._pin: Optional[torch.Tensor] = None
        self._device_out = False       # encode_*_area return the payload area as a device tensor (set by dist.* around its calls)
        # the LZ77 candidate runs beside the BBWT chain: own context, own (non-blocking) stream, one worker thread — its kernels
        # are latency bound (class walks, one parse CTA per block) and fill the gaps of the sort rounds (KOLM_LZ_ASYNC=0: inline)
        self.lz_async = os.environ.get("KOLM_LZ_ASYNC", "1") != "0"
        # kolm_encode_blocks / kolm_decode_blocks: one C-ABI call per batch, sizes / offsets / method ids stay on the device
        # (KOLM_FUSED=0: the stage-by-stage path below, kept for A/B runs and for the candidate lists the fused call does not cover)
        self.fused = os.environ.get("KOLM_FUSED", "1") != "0"
        self.ctx2: Optional[Context] = None
        self.cap2_bytes = 0
        self.cap2_blocks = 0
        self._side: Optional[torch.cuda.Stream] = None
        self._worker = None
        # Re-Pair on long blocks (> kolm_repair_max_block() bytes: the incremental kernel, one CTA per block and seconds per MiB)
        # is run AHEAD of the batch loop over groups of up to repair_group_bytes on a third (Re-Pair-only, 8 bytes of scratch per
        # byte) context / stream / worker thread: a group offers enough blocks for four CTAs on every SM and a work queue that
        # evens out fast and slow blocks (every group ends with a tail as long as its slowest block, so the larger the better),
        # and the BBWT / LZ77 chains of all batches run beside it (KOLM_REPAIR_GROUP_MIB, 0: per batch on the LZ77 side stream)
        self.repair_group_bytes = min(int(os.environ.get("KOLM_REPAIR_GROUP_MIB", "1792")), 1920) << 20
        self.ctx3: Optional[Context] = None
        self.cap3_bytes = 0
        self.cap3_blocks = 0
        self._side3: Optional[torch.cuda.Stream] = None
        self._worker3 = None

    def _home(self, dev: torch.Tensor, n: int) -> np.ndarray:
        """Device bytes -> host through a persistent pinned staging buffer (pageable D2H runs at ~2 GB/s, pinned at PCIe speed)."""
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        if self._pin is None or self._pin.numel() < n:
            self._pin = torch.empty(max(n, 1 << 20) + (n >> 3), dtype=torch.uint8).pin_memory()
        self._pin[:n].copy_(dev[:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pin[:n].numpy()

    @classmethod
    def shared(cls, device: Optional[int] = None) -> "Engine":
        d = torch.cuda.current_device() if device is None else int(device)
        if d not in cls._shared:
            cls._shared[d] = Engine(d)
        return cls._shared[d]

    # ------------------------------------------------------------------
    def cdc_boundaries(self, which: str, data: bytes, mn: int, avg: int, mx: int, piece: int = 64 << 20) -> List[Tuple[int, int]]:
        """Content-defined chunk boundaries with the per-byte scan on the GPU: the data goes up in pieces
        (32 bytes of history each), `kolm_cdc_candidates` returns the positions whose window hash passes the mask test and
        `kolm_cdc_walk_*` (host C++) follows the chain over them.  Bit-identical to the host functions; degenerate inputs whose
        candidate list would not be sparse (e.g. constant data where every position passes) use the host scan."""
        n = len(data)
        if n == 0:
            return []
        if which == "v22" and (not (0 < mn <= avg <= mx) or avg < 64):
            raise ValueError("Require 0 < min_size <= avg_size <= max_size and avg_size >= 64")
        L = _lib.lib()
        self._ensure(1 << 20, 16)
        variant = 0 if which == "kf" else 1
        k = max(6, min(20, int(avg).bit_length() - 1))
        kl = k if variant == 0 else (k - 2 if k > 2 else 1)
        arr = np.frombuffer(data, dtype=np.uint8)
        lists = []
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            sp = C.c_void_p(stream.cuda_stream)
            pin = torch.empty(min(n, piece) + 32, dtype=torch.uint8).pin_memory()
            dbuf = torch.empty(min(n, piece) + 32, dtype=torch.uint8, device="cuda")
            cap = max(4096, 4 * ((min(n, piece) >> kl) + 1))
            dout = torch.empty(cap + 1, dtype=torch.int64, device="cuda")
            for a in range(0, n, piece):
                b = min(n, a + piece)
                hist = min(a, 32)
                m = b - a + hist
                pin[:m].numpy()[:] = arr[a - hist:b]
                dbuf[:m].copy_(pin[:m], non_blocking=True)
                cnt = C.c_int64(0)
                rc = L.kolm_cdc_candidates(self.ctx._h, C.c_void_p(dbuf.data_ptr()), hist, m, a - hist, variant, avg,
                                           C.c_void_p(dout.data_ptr()), cap, C.byref(cnt), sp)
                if rc == -3:                                 # not sparse: the host scan is the right tool
                    return cdc_boundaries(which, data, mn, avg, mx)
                _lib.check(rc)
                if cnt.value:
                    lists.append(dout[1:1 + cnt.value].cpu().numpy().view(np.uint64))
        cand = np.ascontiguousarray(np.concatenate(lists)) if lists else np.zeros(0, dtype=np.uint64)
        capn = n // max(1, mn) + 4
        ends = np.zeros(capn, dtype=np.int64)
        buf = C.cast(C.c_char_p(data), C.c_void_p)
        r = getattr(L, "kolm_cdc_walk_" + which)(buf, n, mn, avg, mx, C.c_void_p(cand.ctypes.data), cand.size,
                                                 ends.ctypes.data_as(C.POINTER(C.c_int64)), capn)
        if r < 0:
            if r == -2:
                raise ValueError("Require 0 < min_size <= avg_size <= max_size and avg_size >= 64")
            raise _lib.KolmError(int(r))
        e = ends[:r]
        return list(zip([0] + e[:-1].tolist(), e.tolist()))

    def _ensure(self, nbytes: int, nblocks: int):
        if self.ctx is None or nbytes > self.cap_bytes or nblocks > self.cap_blocks:
            if self.ctx is not None:
                self.ctx.close()
                self.ctx = None
                torch.cuda.empty_cache()
            self.cap_bytes = max(nbytes, min(self.batch_bytes, 1 << 22), self.cap_bytes)
            self.cap_blocks = max(nblocks, 1024, self.cap_blocks)
            with torch.cuda.device(self.device):
                self.ctx = Context(self.cap_bytes, self.cap_blocks, self.device)

    def _lz_submit(self, x: torch.Tensor, off: np.ndarray, window: int, max_len: int, with_repair: bool = False):
        """Start lz77_encode(x, off) and return a callable that yields (payload, offsets) — with_repair: (payload, offsets,
        (repair payload, repair offsets)), the Re-Pair candidate runs after it on the same side stream.  x must have been
        produced on the current stream; the side stream waits for it."""
        nbytes, nblocks = int(off[-1] - off[0]), len(off) - 1
        if not self.lz_async:
            res = self.ctx.lz77_encode(x, off, window, max_len)
            if with_repair:
                res = res + (self.ctx.repair_encode(x, off),)
            return lambda: res
        if self.ctx2 is None or nbytes > self.cap2_bytes or nblocks > self.cap2_blocks:
            if self.ctx2 is not None:
                self.ctx2.close()
                self.ctx2 = None
            self.cap2_bytes = max(nbytes, min(self.batch_bytes, 1 << 22), self.cap2_bytes)
            self.cap2_blocks = max(nblocks, 1024, self.cap2_blocks)
            self.ctx2 = Context(self.cap2_bytes, self.cap2_blocks, self.device)
        if self._side is None:
            from concurrent.futures import ThreadPoolExecutor
            self._side = torch.cuda.Stream(device=self.device)
            self._worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="kolm-lz")
        ready = torch.cuda.Event()
        ready.record()
        ctx2, side, dev = self.ctx2, self._side, self.device

        def work():
            with torch.cuda.device(dev), torch.cuda.stream(side):
                side.wait_event(ready)
                res = ctx2.lz77_encode(x, off, window, max_len)     # returns after its stream finished (payload offsets come home)
                return res + (ctx2.repair_encode(x, off),) if with_repair else res
        fut = self._worker.submit(work)
        return fut.result

    def _repair_ahead(self, data, bounds: Sequence[Tuple[int, int]]):
        """Queue repair_encode over the whole input in groups of <= repair_group_bytes on the Re-Pair context.  Returns a function
        (i, j) -> (sizes int64[j-i], absolute device addresses uint64[j-i]) that waits for the groups holding blocks [i, j); the
        payload tensors stay alive as long as the returned function does."""
        groups = list(self._batches(bounds, self.repair_group_bytes))
        need_bytes = max(bounds[j - 1][1] - bounds[i][0] for i, j in groups)
        need_blocks = max(j - i for i, j in groups)
        if self._side3 is None:
            from concurrent.futures import ThreadPoolExecutor
            self._side3 = torch.cuda.Stream(device=self.device)
            self._worker3 = ThreadPoolExecutor(max_workers=1, thread_name_prefix="kolm-repair")
        dev, side = self.device, self._side3

        def work(i, j):
            with torch.cuda.device(dev), torch.cuda.stream(side):
                if self.ctx3 is None or need_bytes > self.cap3_bytes or need_blocks > self.cap3_blocks:
                    if self.ctx3 is not None:
                        self.ctx3.close()
                        self.ctx3 = None
                    self.cap3_bytes = max(need_bytes, 1 << 22, self.cap3_bytes)
                    self.cap3_blocks = max(need_blocks, 1024, self.cap3_blocks)
                    self.ctx3 = Context(self.cap3_bytes, self.cap3_blocks, dev, repair_only=True)
                a, b = bounds[i][0], bounds[j - 1][1]
                off = np.array([bounds[k][0] - a for k in range(i, j)] + [b - a], dtype=np.int64)
                x = self._upload(data, a, b)
                try:
                    p, o = self.ctx3.repair_encode(x, off)             # returns after its stream finished
                except _lib.KolmError as e:
                    if e.code == -3:                                   # KOLM_E_CAPACITY: not one slab of the incremental kernel fits
                        raise RuntimeError("Re-Pair candidate: a block of %d bytes needs more device memory than is free; set "
                                           "KOLM_REPAIR_MAX_BLOCK (Engine.repair_max) to skip the candidate on such blocks — the "
                                           "container can then differ from the reference's where Re-Pair would win" % int(np.diff(off).max())) from e
                    raise
                p = p[:max(int(o[-1]), 4)].clone()                      # the capacity is 4x the input: keep what was used
                side.synchronize()
                return p, np.asarray(o, dtype=np.int64)
        futs = [(i, j, self._worker3.submit(work, i, j)) for i, j in groups]

        def take(i, j):
            sizes = np.zeros(j - i, dtype=np.int64)
            addr = np.zeros(j - i, dtype=np.uint64)
            for gi, gj, f in futs:
                lo, hi = max(i, gi), min(j, gj)
                if lo >= hi:
                    continue
                pt, o = f.result()
                sizes[lo - i:hi - i] = np.diff(o)[lo - gi:hi - gi]
                addr[lo - i:hi - i] = np.uint64(pt.data_ptr()) + o[lo - gi:hi - gi].astype(np.uint64)
            return sizes, addr

        def release():
            """After the container: a Re-Pair context sized for GiB groups and its slab pool (up to 100 GB) go back to the
            device; small ones stay for the next call."""
            for _, _, f in futs:
                f.exception()                                          # wait; errors surface where the results are taken
            if self.ctx3 is not None and self.cap3_bytes > (256 << 20):
                def close():
                    with torch.cuda.device(dev):
                        side.synchronize()
                        self.ctx3.close()
                        self.ctx3 = None
                        self.cap3_bytes = self.cap3_blocks = 0
                self._worker3.submit(close).result()
        take.release = release
        return take

    def _batches(self, bounds: Sequence[Tuple[int, int]], limit: Optional[int] = None):
        """Consecutive blocks grouped so that a batch holds <= limit (default batch_bytes) bytes (a single larger block forms its own batch)."""
        limit = self.batch_bytes if limit is None else max(1, int(limit))
        i, n = 0, len(bounds)
        while i < n:
            j, tot = i, 0
            while j < n and (j == i or tot + (bounds[j][1] - bounds[j][0]) <= limit) and j - i < (1 << 20):
                tot += bounds[j][1] - bounds[j][0]
                j += 1
            yield i, j
            i = j

    def _upload(self, data, a: int, b: int) -> torch.Tensor:
        if isinstance(data, torch.Tensor) and data.is_cuda:        # already on the device (dist.*_corpus: payload spans arrive over NCCL)
            if b - a >= 4 and a + max(4, b - a + 4) <= data.numel():
                return data[a:a + max(4, b - a + 4)]
            t = torch.zeros(max(4, b - a + 4), dtype=torch.uint8, device=data.device)
            t[:b - a].copy_(data[a:b])
            return t
        t = torch.empty(max(4, b - a + 4), dtype=torch.uint8, device=torch.device("cuda", self.device))
        if b > a:
            # numpy view of the (read-only) bytes -> torch: no warning to filter (warnings.catch_warnings is not thread-safe and
            # _upload runs on the main thread and on the side-stream workers at once)
            arr = np.frombuffer(data, dtype=np.uint8, count=b - a, offset=a)
            t[:b - a].copy_(_ro_tensor(arr))
        return t

    @staticmethod
    def _host(t: torch.Tensor, n: int) -> bytes:
        return t[:n].cpu().numpy().tobytes() if n else b""

    # ------------------------------------------------------------------
    # KOLM profile
    def _gather_home(self, c: Context, addr: np.ndarray, lens: np.ndarray, keep: bool = True):
        """Winners' payloads -> one device buffer in block order -> host memory.  Returns a uint8 numpy array; with keep=False it
        is a view of the pinned staging buffer, valid until the next engine call (the caller joins it into the container at once).
        With self._device_out set (dist.*: the payloads travel GPU -> GPU over NCCL) the device tensor is returned instead."""
        total = int(lens.sum())
        dev = torch.empty(max(total, 4) + 16, dtype=torch.uint8, device=torch.device("cuda", self.device))
        c.gather_payloads(addr, lens, dev)
        if self._device_out:
            torch.cuda.current_stream().synchronize()                # the candidates' payload tensors die when the caller moves on
            return dev[:total]
        v = self._home(dev, total)
        return v.copy() if keep else v

    def _area_home(self, dev: torch.Tensor, total: int, keep: bool = True):
        """Payload area of a fused call -> host (or the device tensor itself for dist.*, see _gather_home)."""
        if self._device_out:
            torch.cuda.current_stream().synchronize()
            return dev[:total]
        v = self._home(dev, total)
        return v.copy() if keep else v

    @staticmethod
    def _cat(areas):
        if not areas:
            return np.zeros(0, np.uint8)
        if len(areas) == 1:
            return areas[0]
        return torch.cat(areas) if isinstance(areas[0], torch.Tensor) else np.concatenate(areas)

    def encode_kolm_area(self, data: bytes, bounds: Sequence[Tuple[int, int]]):
        """-> (method ids int64[nb], payload lengths int64[nb], payload area uint8[sum]) for the KOLM candidates."""
        mids_all, lens_all, areas = [], [], []
        for i, j in self._batches(bounds):
            a, b = bounds[i][0], bounds[j - 1][1]
            nb = j - i
            off = np.array([bounds[k][0] - a for k in range(i, j)] + [b - a], dtype=np.int64)
            lens = np.diff(off)
            self._ensure(b - a, nb)
            if self.fused:
                with torch.cuda.device(self.device):
                    x = self._upload(data, a, b)
                    out, poff, mids = self.ctx.encode_blocks(_lib.KOLM_PROFILE_KOLM, x, off)
                    areas.append(self._area_home(out, int(poff[-1]), keep=j < len(bounds)))
                mids_all.append(mids.astype(np.int64))
                lens_all.append(np.diff(poff))
                continue
            with torch.cuda.device(self.device):
                c = self.ctx
                x = self._upload(data, a, b)
                lz = self._lz_submit(x, off, 255, 127)
                sx = c.residual_sizes(x, off)[:, 0]
                L = c.bbwt_forward(x, off)
                m = c.mtf_encode(L, off)
                kfp, kfo = c.rice_kf_encode(m, off)
                lzp, lzo = lz()
                sizes = np.stack([lens, sx, np.diff(kfo), np.diff(lzo)], axis=1)
                mids, _ = c.select_blocks(sizes)                     # first minimum == lowest id on ties (KF.py:857)
                base = np.zeros((nb, 4), dtype=np.uint64)
                base[:, 0] = np.uint64(x.data_ptr()) + off[:-1].astype(np.uint64)
                if (mids == 1).any():
                    xp, xo = c.residual_encode(x, off, 0)
                    base[:, 1] = np.uint64(xp.data_ptr()) + xo[:-1].astype(np.uint64)
                base[:, 2] = np.uint64(kfp.data_ptr()) + kfo[:-1].astype(np.uint64)
                base[:, 3] = np.uint64(lzp.data_ptr()) + lzo[:-1].astype(np.uint64)
                plen = sizes[np.arange(nb), mids].astype(np.int64)
                areas.append(self._gather_home(c, base[np.arange(nb), mids], plen, keep=j < len(bounds)))
            mids_all.append(mids.astype(np.int64))
            lens_all.append(plen)
        if not areas:
            return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8)
        return np.concatenate(mids_all), np.concatenate(lens_all), self._cat(areas)

    def encode_kolm(self, data: bytes, bounds: Sequence[Tuple[int, int]]) -> List[Tuple[int, bytes]]:
        mids, lens, area = self.encode_kolm_area(data, bounds)
        ends = np.cumsum(lens)
        return [(int(mids[k]), area[ends[k] - lens[k]:ends[k]].tobytes()) for k in range(len(mids))]

    def kolm_model_payloads(self, block: bytes, mid: int) -> bytes:
        """One model on one block (the reference's _ENCODERS[mid])."""
        off = np.array([0, len(block)], dtype=np.int64)
        self._ensure(len(block), 1)
        with torch.cuda.device(self.device):
            c = self.ctx
            x = self._upload(block, 0, len(block))
            if mid == 0:
                return bytes(block)
            if mid == 1:
                p, o = c.residual_encode(x, off, 0)
            elif mid == 2:
                p, o = c.rice_kf_encode(c.mtf_encode(c.bbwt_forward(x, off), off), off)
            elif mid == 3:
                p, o = c.lz77_encode(x, off, 255, 127)
            else:
                raise KeyError(mid)
            return self._host(p, int(o[-1]))

    # ------------------------------------------------------------------
    # KOLR profile.  `names` = candidate names in id order (the reference's _select_encoders() list).
    def encode_kolr_area(self, data: bytes, bounds: Sequence[Tuple[int, int]], names: Sequence[str]):
        """-> (method ids int64[nb], payload lengths int64[nb], payload area uint8[sum]) for the KOLR candidate list `names`."""
        mids_all, lens_all, areas = [], [], []
        v2 = self.enable_v2_new and "v2_new" in names
        longest = max((b - a for a, b in bounds), default=0)
        rp_take = None
        if ("repair" in names and self.lz_async and self.repair_group_bytes > 0 and
                _lib.lib().kolm_repair_max_block() < longest <= self.repair_max):
            rp_take = self._repair_ahead(data, bounds)
        for i, j in self._batches(bounds, self.batch_bytes // 8 if v2 else None):   # v2_new sorts 8 planes per block: same scratch per batch
            a, b = bounds[i][0], bounds[j - 1][1]
            nb = j - i
            off = np.array([bounds[k][0] - a for k in range(i, j)] + [b - a], dtype=np.int64)
            lens = np.diff(off)
            if v2:                                                   # the eight bit planes of every block are sorted as one batch
                self._ensure(max(8 * (b - a) + 64, 1024 * nb), 8 * nb)
            else:
                self._ensure(b - a, nb)
            want_rp_f = "repair" in names and int(lens.max(initial=0)) <= self.repair_max
            if self.fused and not v2 and list(names) == KOLR_NAMES:      # the reference's full candidate list: ids are list indices
                if "repair" in names and not want_rp_f:
                    warnings.warn("Re-Pair candidate skipped: block longer than the configured cap of %d bytes (Engine.repair_max / "
                                  "KOLM_REPAIR_MAX_BLOCK); the container can differ from the reference's where Re-Pair would win" % self.repair_max,
                                  RuntimeWarning, stacklevel=3)
                with torch.cuda.device(self.device):
                    x = self._upload(data, a, b)
                    ext, mask = None, 0x3FF
                    if not want_rp_f:
                        mask = 0x1FF
                    elif rp_take is not None:                        # Re-Pair of long blocks ran ahead on its own context
                        rs_, ra_ = rp_take(i, j)
                        ext = (9, rs_, ra_)
                    out, poff, mids = self.ctx.encode_blocks(_lib.KOLM_PROFILE_KOLR, x, off, cand_mask=mask, ext=ext)
                    areas.append(self._area_home(out, int(poff[-1]), keep=j < len(bounds)))
                mids_all.append(mids.astype(np.int64))
                lens_all.append(np.diff(poff))
                continue
            with torch.cuda.device(self.device):
                c = self.ctx
                x = self._upload(data, a, b)
                cols = []
                v2p = v2o = None
                if v2:
                    v2p, v2o = c.v2new_encode(x, off)
                want_rp = "repair" in names and int(lens.max(initial=0)) <= self.repair_max
                if "repair" in names and not want_rp:
                    warnings.warn("Re-Pair candidate skipped: block longer than the configured cap of %d bytes (Engine.repair_max / "
                                  "KOLM_REPAIR_MAX_BLOCK); the container can differ from the reference's where Re-Pair would win" % self.repair_max,
                                  RuntimeWarning, stacklevel=3)
                rp_side = want_rp and rp_take is None and "lz77" in names and self.lz_async   # both latency-bound candidates share the side stream
                lz = self._lz_submit(x, off, 4096, 0, with_repair=rp_side) if "lz77" in names else None
                need_res = any(n in ("xor", "lfsr_pred") for n in names)
                rs = c.residual_sizes(x, off) if need_res else None
                need_bbwt = any(n in K2_FLAG_OF for n in names)
                m = k2p = k2o = k2s = None
                if need_bbwt:
                    m = c.mtf_encode(c.bbwt_forward(x, off), off)
                    k2p, k2o, k2s = c.rice_k2_encode(m, off, 0)
                lzp = lzo = rpp = rpo = None
                if lz is not None:
                    got = lz()
                    lzp, lzo = got[0], got[1]
                    if rp_side:
                        rpp, rpo = got[2]
                rp_sizes = rp_addr = None
                if want_rp and rp_take is not None:
                    rp_sizes, rp_addr = rp_take(i, j)
                elif want_rp and rpo is None:
                    rpp, rpo = c.repair_encode(x, off)
                for nme in names:
                    if nme == "raw":
                        cols.append(lens)
                    elif nme == "xor":
                        cols.append(rs[:, 1])
                    elif nme == "lfsr_pred":
                        cols.append(rs[:, 2])
                    elif nme in K2_FLAG_OF:
                        cols.append(k2s[:, K2_SLOT[K2_FLAG_OF[nme]]])
                    elif nme == "lz77":
                        cols.append(np.diff(lzo))
                    elif nme == "repair" and rp_sizes is not None:
                        cols.append(rp_sizes)
                    elif nme == "repair" and rpo is not None:
                        cols.append(np.diff(rpo))
                    elif nme == "v2_new" and v2o is not None:
                        cols.append(np.diff(v2o))
                    else:                                            # v2_new raises NameError in the shipped reference; skipped repair
                        cols.append(np.full(nb, _BIG, dtype=np.int64))
                sizes = np.stack(cols, axis=1)
                mids, _ = c.select_blocks(sizes)                     # strict '<' in the reference == first minimum
                base = np.zeros((nb, len(names)), dtype=np.uint64)
                keep = [x]                                           # keep every payload tensor alive until the gather ran
                for mid in sorted(set(int(v) for v in mids)):
                    nme = names[mid]
                    if nme == "raw":
                        base[:, mid] = np.uint64(x.data_ptr()) + off[:-1].astype(np.uint64)
                        continue
                    if nme == "xor":
                        p, o = c.residual_encode(x, off, 1)
                    elif nme == "lfsr_pred":
                        p, o = c.residual_encode(x, off, 2)
                    elif nme == "bbwt":
                        p, o = k2p, k2o
                    elif nme in K2_FLAG_OF:
                        p, o, _ = c.rice_k2_encode(m, off, K2_FLAG_OF[nme])
                    elif nme == "lz77":
                        p, o = lzp, lzo
                    elif nme == "repair" and rp_addr is not None:
                        base[:, mid] = rp_addr                       # the groups' payload tensors live in rp_take
                        continue
                    elif nme == "repair":
                        p, o = rpp, rpo
                    elif nme == "v2_new":
                        p, o = v2p, v2o
                    else:
                        raise RuntimeError("unreachable candidate " + nme)
                    keep.append(p)
                    base[:, mid] = np.uint64(p.data_ptr()) + o[:-1].astype(np.uint64)
                plen = sizes[np.arange(nb), mids].astype(np.int64)
                areas.append(self._gather_home(c, base[np.arange(nb), mids], plen, keep=j < len(bounds)))
            mids_all.append(mids.astype(np.int64))
            lens_all.append(plen)
        if rp_take is not None:
            rp_take.release()
        if not areas:
            return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8)
        return np.concatenate(mids_all), np.concatenate(lens_all), self._cat(areas)

    def encode_kolr(self, data: bytes, bounds: Sequence[Tuple[int, int]], names: Sequence[str]) -> List[Tuple[int, bytes]]:
        mids, lens, area = self.encode_kolr_area(data, bounds, names)
        ends = np.cumsum(lens)
        return [(int(mids[k]), area[ends[k] - lens[k]:ends[k]].tobytes()) for k in range(len(mids))]

    def kolr_model_payload(self, block: bytes, name: str) -> bytes:
        off = np.array([0, len(block)], dtype=np.int64)
        self._ensure(len(block), 1)
        with torch.cuda.device(self.device):
            c = self.ctx
            x = self._upload(block, 0, len(block))
            if name == "raw":
                return bytes(block)
            if name == "xor":
                p, o = c.residual_encode(x, off, 1)
            elif name == "lfsr_pred":
                p, o = c.residual_encode(x, off, 2)
            elif name in K2_FLAG_OF:
                p, o, _ = c.rice_k2_encode(c.mtf_encode(c.bbwt_forward(x, off), off), off, K2_FLAG_OF[name])
            elif name == "lz77":
                p, o = c.lz77_encode(x, off, 4096, 0)
            elif name == "repair":
                p, o = c.repair_encode(x, off)
            elif name == "v2_new" and self.enable_v2_new:
                self._ensure(max(8 * len(block) + 64, 1024), 8)
                p, o = self.ctx.v2new_encode(x, off)
            else:
                raise NameError("name 'os' is not defined")          # v2_new: what the shipped reference raises
            return self._host(p, int(o[-1]))

    # ------------------------------------------------------------------
    # decode: blocks = [(method name, payload, orig_len)]
    def decode_area(self, blocks: Sequence[Tuple[str, bytes, int]]) -> np.ndarray:
        """All blocks decoded back to back (uint8 array of sum(orig_len) bytes); blocks = [(method name, payload, orig_len)]."""
        names = [b[0] for b in blocks]
        plens = np.array([len(b[1]) for b in blocks], dtype=np.int64)
        starts = np.zeros(len(blocks), dtype=np.int64)
        if len(blocks):
            starts[1:] = np.cumsum(plens)[:-1]
        out = self._decode_spans(b"".join(b[1] for b in blocks), names, starts, plens, np.array([b[2] for b in blocks], dtype=np.int64))
        return np.frombuffer(out, dtype=np.uint8)

    def decode_container(self, blob: bytes, names: Sequence[str], starts, plens, orig_lens) -> bytes:
        """Container decode: payload k is blob[starts[k] : starts[k] + plens[k]] (container order).  The payload bytes of an
        output batch travel to the device in ONE copy; each method group is compacted there (`kolm_gather_payloads`), decoded
        as one batch and copied to its blocks' final offsets; one D2H per output batch, straight into the result."""
        return self._decode_spans(blob, list(names), np.asarray(starts, dtype=np.int64), np.asarray(plens, dtype=np.int64),
                                  np.asarray(orig_lens, dtype=np.int64))

    def decode_to_device(self, blob, names: Sequence[str], starts, plens, orig_lens) -> torch.Tensor:
        """Like decode_container, but the decoded bytes stay on the device (one uint8 tensor; dist.decompress_* sends it over NCCL)."""
        ols = np.asarray(orig_lens, dtype=np.int64)
        out = torch.empty(max(int(ols.sum()), 1), dtype=torch.uint8, device=torch.device("cuda", self.device))

        def keep(base_off, dev_out, tot):
            out[base_off:base_off + tot].copy_(dev_out[:tot])
        self._decode_spans(blob, list(names), np.asarray(starts, dtype=np.int64), np.asarray(plens, dtype=np.int64), ols, on_batch=keep)
        return out[:int(ols.sum())]

    def _decode_spans(self, blob, names: List[str], starts: np.ndarray, plens: np.ndarray, ols: np.ndarray, on_batch=None) -> bytes:
        nb = len(names)
        if nb == 0:
            return b""
        ends = np.cumsum(ols)
        # untrusted header: every coder here expands a payload byte into a bounded number of output bytes except the run /
        # match / grammar coders, so only sanity-bound the allocation (the reference builds its output incrementally and
        # fails on the first bad block instead of a MemoryError)
        # lengths are allocated up front only when plausible; an implausible header (the lengths are untrusted) is decoded batch by
        # batch into a growing buffer, so a bad block fails with the reference's exception instead of a MemoryError
        blob_len = blob.numel() if isinstance(blob, torch.Tensor) else len(blob)
        lazy = on_batch is None and int(ends[-1]) > max(1 << 30, 4096 * blob_len)
        grow = bytearray() if lazy else None
        result, sink = (b"", None) if (on_batch is not None or lazy) else _new_bytes(int(ends[-1]))
        if sink is None and on_batch is None and not lazy:           # 0 or 1 byte: CPython shares these objects, build them the ordinary way
            sink = np.zeros(int(ends[-1]), dtype=np.uint8)
        i = 0
        uniq = set(names)                                            # 61 440 blocks at the default block size: no per-block Python here
        any_kf = any(nm.startswith("kf_") for nm in uniq)
        any_v2 = "v2_new" in uniq
        ends_l = np.asarray(ends, dtype=np.int64)
        while i < nb:                                                # output batches of <= batch_bytes (at least one block)
            done = int(ends_l[i - 1]) if i else 0
            j = max(i + 1, int(np.searchsorted(ends_l, done + self.batch_bytes, side="right")))
            tot = int(ends_l[j - 1]) - done
            base_off = int(ends[i] - ols[i])
            span0, span1 = int(starts[i]), int((starts[i:j] + plens[i:j]).max())
            with torch.cuda.device(self.device):
                dev_out = torch.empty(max(tot, 4) + 16, dtype=torch.uint8, device=torch.device("cuda", self.device))
                span = self._upload(blob, span0, max(span0, span1))
                if self.fused and not (any_v2 and "v2_new" in names[i:j]):
                    # one kolm_decode_blocks call for the output batch (groups by method inside)
                    kolm_prof = any_kf and any(nm.startswith("kf_") for nm in names[i:j])
                    ids = _KOLM_IDS if kolm_prof else _KOLR_IDS
                    try:
                        mids = np.array([ids[nm] for nm in names[i:j]], dtype=np.uint8)
                    except KeyError as ke:
                        raise NotImplementedError("no decoder for method '%s'" % ke.args[0])
                    ooff = np.zeros(j - i + 1, dtype=np.int64)
                    ooff[1:] = np.cumsum(ols[i:j])
                    self._ensure(max(tot, int(plens[i:j].sum()), 1), j - i)
                    try:
                        self.ctx.decode_blocks(_lib.KOLM_PROFILE_KOLM if kolm_prof else _lib.KOLM_PROFILE_KOLR, span, starts[i:j] - span0, plens[i:j], mids, ooff, out=dev_out)
                    except _lib.KolmError as err:
                        blk = getattr(err, "block", -1)
                        if blk >= 0 and names[i + blk] == "raw" and err.code == -5:
                            raise AssertionError("Payload length mismatch for RAW") from err
                        if err.code == -4 and not kolm_prof:
                            raise ValueError(str(err)) from err     # V22's readers raise ValueError on truncation (v2-2.py:131-132, 1437-1447)
                        raise_like_reference(err)
                    if tot and on_batch is not None:
                        on_batch(base_off, dev_out, tot)
                        torch.cuda.current_stream().synchronize()
                    elif tot and lazy:
                        grow += self._home(dev_out, tot).tobytes()
                    elif tot:
                        _par_copy(sink[base_off:base_off + tot], self._home(dev_out, tot))
                    i = j
                    continue
                groups: Dict[Tuple[str, int], List[int]] = {}
                v2_tot, v2_part = 0, 0
                for idx in range(i, j):
                    part = 0
                    if names[idx] == "v2_new":                       # eight planes per block are sorted: keep these sub-batches 8x smaller
                        if v2_tot and v2_tot + int(ols[idx]) > self.batch_bytes // 8:
                            v2_part, v2_tot = v2_part + 1, 0
                        v2_tot += int(ols[idx])
                        part = v2_part
                    groups.setdefault((names[idx], part), []).append(idx)
                for (nme, _part), sub in groups.items():
                    sub = np.asarray(sub, dtype=np.int64)
                    sol, spl = ols[sub], plens[sub]
                    off = np.zeros(len(sub) + 1, dtype=np.int64)
                    off[1:] = np.cumsum(sol)
                    dst = np.uint64(dev_out.data_ptr()) + (ends[sub] - sol - base_off).astype(np.uint64)
                    src_pay = np.uint64(span.data_ptr()) + (starts[sub] - span0).astype(np.uint64)
                    ptot = int(spl.sum())
                    self._ensure(max(int(off[-1]), ptot, 1), len(sub))
                    c = self.ctx
                    if nme == "raw":
                        if not bool((spl == sol).all()):             # untrusted container: never an `assert` (gone under python -O)
                            raise AssertionError("Payload length mismatch for RAW")
                        c.copy_blocks(src_pay, dst, sol)
                        torch.cuda.current_stream().synchronize()
                        continue
                    pt = torch.empty(max(ptot, 4) + 16, dtype=torch.uint8, device=span.device)
                    poff = c.gather_payloads(src_pay, spl, pt)
                    try:
                        if nme in ("kf_xor", "xor", "lfsr_pred"):
                            y = c.residual_decode(pt, poff, off, {"kf_xor": 0, "xor": 1, "lfsr_pred": 2}[nme])
                        elif nme == "kf_bbwt":
                            y = c.bbwt_inverse(c.mtf_decode(c.rice_kf_decode(pt, poff, off), off), off)
                        elif nme in K2_FLAG_OF:
                            y = c.bbwt_inverse(c.mtf_decode(c.rice_k2_decode(pt, poff, off, K2_FLAG_OF[nme]), off), off)
                        elif nme == "kf_lz77":
                            y = c.lz77_decode(pt, poff, off, 0)
                        elif nme == "lz77":
                            y = c.lz77_decode(pt, poff, off, 4096)
                        elif nme == "repair":
                            y = c.repair_decode(pt, poff, off)
                        elif nme == "v2_new":
                            self._ensure(8 * int(off[-1]) + 64, 8 * len(sub))   # the eight bit planes of every block form one batch
                            c = self.ctx
                            y = c.v2new_decode(pt, poff, off)
                        else:
                            raise NotImplementedError("no decoder for method '%s'" % nme)
                    except _lib.KolmError as err:
                        if err.code == -4 and not nme.startswith("kf_"):
                            raise ValueError(str(err)) from err     # V22's readers raise ValueError on truncation (v2-2.py:131-132, 1437-1447)
                        raise_like_reference(err)
                    src = np.uint64(y.data_ptr()) + off[:-1].astype(np.uint64)
                    self.ctx.copy_blocks(src, dst, sol)
                    torch.cuda.current_stream().synchronize()        # y / pt may be freed when the loop moves on
                if tot and on_batch is not None:
                    on_batch(base_off, dev_out, tot)
                    torch.cuda.current_stream().synchronize()
                elif tot and lazy:
                    grow += self._home(dev_out, tot).tobytes()
                elif tot:
                    _par_copy(sink[base_off:base_off + tot], self._home(dev_out, tot))
            i = j
        if on_batch is not None:
            return b""
        if lazy:
            return bytes(grow)
        return result if len(result) >= 2 else sink.tobytes()

    def decode_blocks(self, blocks: Sequence[Tuple[str, bytes, int]]) -> List[bytes]:
        area = self.decode_area(blocks)
        res, p = [], 0
        for _, _, ol in blocks:
            res.append(area[p:p + ol].tobytes())
            p += ol
        return res
