"""Batched stage operators on torch CUDA uint8 tensors — thin wrappers over the C-ABI.

A batch is a 1-D uint8 CUDA tensor holding nblocks blocks back to back plus a Python/numpy list of
nblocks+1 offsets.  Function names follow the reference's stage functions:
bbwt_forward, bbwt_inverse, mtf_encode, mtf_decode, duval_lyndon, ...
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def _offsets(off: Sequence[int]):
    a = np.ascontiguousarray(np.asarray(off, dtype=np.int64))
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


class Context:
    """Owns the device scratch for batches up to (max_batch_bytes, max_blocks) on one device."""

    def __init__(self, max_batch_bytes: int, max_blocks: int, device: Optional[int] = None, repair_only: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("kolmogorovlike_datacompressor_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_batch_bytes = int(max_batch_bytes)
        self.max_blocks = int(max_blocks)
        h = C.c_void_p()
        # repair_only: KOLM_CTX_REPAIR_ONLY — only repair_encode works on such a context (8 instead of 38 bytes of scratch per byte)
        _lib.check(_lib.lib().kolm_create_ex(self.device, max(1, self.max_batch_bytes), max(1, self.max_blocks), 1 if repair_only else 0, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().kolm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _n2n(self, fn: str, x: torch.Tensor, off, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
        oa, op = _offsets(off)
        assert int(oa[-1]) <= x.numel()
        if out is None:
            out = torch.empty_like(x)
        _lib.check(getattr(_lib.lib(), fn)(self._h, C.c_void_p(x.data_ptr()), op, len(oa) - 1, C.c_void_p(out.data_ptr()), self._stream()), fn)
        return out

    def duval_lyndon_flags(self, x, off):
        """1 where a Lyndon factor starts (kolm_final.py:200-225)."""
        return self._n2n("kolm_lyndon", x, off)

    def bbwt_forward(self, x, off, out=None):
        return self._n2n("kolm_bbwt_fwd", x, off, out)

    def bbwt_inverse(self, x, off, out=None):
        return self._n2n("kolm_bbwt_inv", x, off, out)

    def mtf_encode(self, x, off, out=None):
        return self._n2n("kolm_mtf_enc", x, off, out)

    def mtf_decode(self, x, off, out=None):
        return self._n2n("kolm_mtf_dec", x, off, out)

    def counters(self):
        a = (C.c_int64 * 4)()
        _lib.lib().kolm_last_counters(self._h, a)
        return dict(rounds_plain=a[0], rounds_cyclic=a[1], launches=a[2], records_sorted=a[3])

    # ---- entropy coders -------------------------------------------------
    def rice_kf_encode(self, mtf, off, out: Optional[torch.Tensor] = None, want_params=False):
        """KF model-2 token stream of each block's MTF sequence -> (payload tensor, out_off ndarray[, params])."""
        oa, op = _offsets(off)
        nb = len(oa) - 1
        if out is None:
            out = torch.empty(int(oa[-1] - oa[0]) * 2 + 16 * nb + 64, dtype=torch.uint8, device=mtf.device)
        out_off = np.zeros(nb + 1, dtype=np.int64)
        params = np.zeros(4 * max(1, nb), dtype=np.int32)
        _lib.check(_lib.lib().kolm_rice_kf_enc(self._h, C.c_void_p(mtf.data_ptr()), op, nb, C.c_void_p(out.data_ptr()), out.numel(),
                                               out_off.ctypes.data_as(C.POINTER(C.c_int64)), params.ctypes.data_as(C.POINTER(C.c_int)),
                                               self._stream()), "kolm_rice_kf_enc")
        return (out, out_off, params.reshape(-1, 4)[:nb]) if want_params else (out, out_off)

    def rice_kf_decode(self, payload, pay_off, off, out=None):
        pa, pp = _offsets(pay_off)
        oa, op = _offsets(off)
        nb = len(oa) - 1
        if out is None:
            out = torch.empty(int(oa[-1]), dtype=torch.uint8, device=payload.device)
        _lib.check(_lib.lib().kolm_rice_kf_dec(self._h, C.c_void_p(payload.data_ptr()), pp, op, nb, C.c_void_p(out.data_ptr()), self._stream()),
                   "kolm_rice_kf_dec")
        return out

    def rice_k2_encode(self, mtf, off, flags: int, out: Optional[torch.Tensor] = None):
        """V22 models 2-6 -> (payload tensor, out_off ndarray, sizes ndarray[nb,5])."""
        oa, op = _offsets(off)
        nb = len(oa) - 1
        if out is None:
            out = torch.empty(int(oa[-1] - oa[0]) * 9 + 16 * nb + 64, dtype=torch.uint8, device=mtf.device)
        out_off = np.zeros(nb + 1, dtype=np.int64)
        sizes = np.zeros(5 * max(1, nb), dtype=np.int64)
        _lib.check(_lib.lib().kolm_rice_k2_enc(self._h, C.c_void_p(mtf.data_ptr()), op, nb, int(flags), C.c_void_p(out.data_ptr()), out.numel(),
                                               out_off.ctypes.data_as(C.POINTER(C.c_int64)), sizes.ctypes.data_as(C.POINTER(C.c_int64)),
                                               self._stream()), "kolm_rice_k2_enc")
        return out, out_off, sizes.reshape(-1, 5)[:nb]

    def rice_dual_encode(self, mtf, off, k2_flags: int, kf_out: Optional[torch.Tensor] = None, k2_out: Optional[torch.Tensor] = None):
        """KF model 2 and the V22 Rice variants of one MTF batch from ONE cost read (kolm_rice_dual_enc) ->
        (kf payload, kf_off, kf params[nb,4], k2 payload, k2_off, k2 sizes[nb,5])."""
        oa, op = _offsets(off)
        nb = len(oa) - 1
        n = int(oa[-1] - oa[0])
        if kf_out is None:
            kf_out = torch.empty(3 * n + 16 * nb + 64, dtype=torch.uint8, device=mtf.device)
        if k2_out is None:
            k2_out = torch.empty(9 * n + 16 * nb + 64, dtype=torch.uint8, device=mtf.device)
        kf_off = np.zeros(nb + 1, dtype=np.int64)
        k2_off = np.zeros(nb + 1, dtype=np.int64)
        params = np.zeros(4 * max(1, nb), dtype=np.int32)
        sizes = np.zeros(5 * max(1, nb), dtype=np.int64)
        _lib.check(_lib.lib().kolm_rice_dual_enc(self._h, C.c_void_p(mtf.data_ptr()), op, nb, int(k2_flags),
                                                 C.c_void_p(kf_out.data_ptr()), kf_out.numel(), kf_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 params.ctypes.data_as(C.POINTER(C.c_int)),
                                                 C.c_void_p(k2_out.data_ptr()), k2_out.numel(), k2_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 sizes.ctypes.data_as(C.POINTER(C.c_int64)), self._stream()), "kolm_rice_dual_enc")
        return kf_out, kf_off, params.reshape(-1, 4)[:nb], k2_out, k2_off, sizes.reshape(-1, 5)[:nb]

    def rice_k2_decode(self, payload, pay_off, off, flags: int, out=None):
        pa, pp = _offsets(pay_off)
        oa, op = _offsets(off)
        nb = len(oa) - 1
        if out is None:
            out = torch.empty(int(oa[-1]), dtype=torch.uint8, device=payload.device)
        _lib.check(_lib.lib().kolm_rice_k2_dec(self._h, C.c_void_p(payload.data_ptr()), pp, op, nb, int(flags), C.c_void_p(out.data_ptr()),
                                               self._stream()), "kolm_rice_k2_dec")
        return out

    # ---- LZ77 / residual coders -----------------------------------------
    def _enc(self, fn, x, off, extra, cap_mul, out=None):
        oa, op = _offsets(off)
        nb = len(oa) - 1
        if out is None:
            out = torch.empty(int(oa[-1] - oa[0]) * cap_mul + 16 * nb + 64, dtype=torch.uint8, device=x.device)
        out_off = np.zeros(nb + 1, dtype=np.int64)
        _lib.check(getattr(_lib.lib(), fn)(self._h, C.c_void_p(x.data_ptr()), op, nb, *extra, C.c_void_p(out.data_ptr()), out.numel(),
                                           out_off.ctypes.data_as(C.POINTER(C.c_int64)), self._stream()), fn)
        return out, out_off

    def _dec(self, fn, payload, pay_off, off, extra, out=None):
        pa, pp = _offsets(pay_off)
        oa, op = _offsets(off)
        nb = len(oa) - 1
        if out is None:
            out = torch.empty(max(1, int(oa[-1])), dtype=torch.uint8, device=payload.device)
        _lib.check(getattr(_lib.lib(), fn)(self._h, C.c_void_p(payload.data_ptr()), pp, op, nb, *extra, C.c_void_p(out.data_ptr()),
                                           self._stream()), fn)
        return out

    def lz77_encode(self, x, off, window: int, max_len: int, out=None):
        """encode_model_lz77 (window 255, max_len 127) / encode_lz77 (window 4096, max_len 0 = unbounded)."""
        return self._enc("kolm_lz77_enc", x, off, (C.c_uint32(window), C.c_uint32(max_len)), 2, out)

    def lz77_decode(self, payload, pay_off, off, window_check: int = 0, out=None):
        return self._dec("kolm_lz77_dec", payload, pay_off, off, (C.c_uint32(window_check),), out)

    def residual_sizes(self, x, off):
        """Exact payload sizes [nb, 3] of the XOR / delta / LFSR-predictor candidates."""
        oa, op = _offsets(off)
        nb = len(oa) - 1
        sizes = np.zeros(3 * max(1, nb), dtype=np.int64)
        _lib.check(_lib.lib().kolm_residual_sizes(self._h, C.c_void_p(x.data_ptr()), op, nb, sizes.ctypes.data_as(C.POINTER(C.c_int64)),
                                                  self._stream()), "kolm_residual_sizes")
        return sizes.reshape(-1, 3)[:nb]

    def residual_encode(self, x, off, kind: int, out=None):
        return self._enc("kolm_residual_enc", x, off, (C.c_int(kind),), 2, out)

    def residual_decode(self, payload, pay_off, off, kind: int, out=None):
        return self._dec("kolm_residual_dec", payload, pay_off, off, (C.c_int(kind),), out)

    def repair_encode(self, x, off, out=None):
        """repair_compress per block (V22.py:1841-1911): blocks <= kolm_repair_max_block() bytes in shared memory, longer ones through
        the incremental kernel (exact at any length)."""
        return self._enc("kolm_repair_enc", x, off, (), 4 if int(np.asarray(off)[-1]) < (1 << 28) else 5, out)

    def repair_decode(self, payload, pay_off, off, out=None):
        return self._dec("kolm_repair_dec", payload, pay_off, off, (), out)

    def v2new_encode(self, x, off, out=None):
        """encode_new_pipeline per block with parallel=False semantics (method 10).  Context of 8x the batch, like the decoder."""
        return self._enc("kolm_v2new_enc", x, off, (), 2, out)

    def v2new_decode(self, payload, pay_off, off, out=None):
        """decode_new_pipeline per block (method 10).  The context must hold 8x the batch (the bit planes are one batch)."""
        return self._dec("kolm_v2new_dec", payload, pay_off, off, (), out)

    # ---- fused hot path ---------------------------------------------------
    def _scratch(self, nbytes: int, device) -> torch.Tensor:
        """Device scratch of the fused calls, kept between calls (256-byte aligned: torch allocations are 512-byte aligned)."""
        t = getattr(self, "_fused_scratch", None)
        if t is None or t.numel() < nbytes or t.device != device:
            self._fused_scratch = None
            t = self._fused_scratch = torch.empty(int(nbytes) + (int(nbytes) >> 4) + 4096, dtype=torch.uint8, device=device)
        return t

    def encode_blocks(self, profile: int, x: torch.Tensor, off, cand_mask: int = 0, ext=None, out: Optional[torch.Tensor] = None,
                      want_sizes: bool = False):
        """kolm_encode_blocks: _encode_block (profile 1, KF.py:821-864) / the KOLR selection loops (profile 2, V22.py:2350-2369) for a
        whole batch, every candidate on the device -> (payload area tensor, payload_off int64[nb+1], method ids uint8[nb]
        [, sizes int64[nb, ncand]]).  ext = (candidate id, sizes int64[nb], device addresses uint64[nb]) for a candidate that was
        computed elsewhere."""
        oa, op = _offsets(off)
        nb = len(oa) - 1
        n = int(oa[-1] - oa[0])
        L = _lib.lib()
        scratch = self._scratch(L.kolm_encode_blocks_scratch(int(profile), max(n, 1), max(nb, 1)), x.device)
        if out is None:
            out = torch.empty(n + 16 * nb + 64, dtype=torch.uint8, device=x.device)      # a winner is never larger than the block (raw)
        poff = np.zeros(nb + 1, dtype=np.int64)
        mids = np.zeros(max(nb, 1), dtype=np.uint8)
        ncand = 4 if int(profile) == 1 else 10
        sizes = np.zeros(max(nb, 1) * ncand, dtype=np.int64) if want_sizes else None
        ext_id, ext_sizes, ext_addr = -1, None, None
        if ext is not None:
            ext_id = int(ext[0])
            ext_sizes = np.ascontiguousarray(ext[1], dtype=np.int64)
            ext_addr = np.ascontiguousarray(ext[2], dtype=np.uint64)
        _lib.check(L.kolm_encode_blocks(self._h, int(profile), C.c_void_p(x.data_ptr()), op, nb, int(cand_mask), ext_id,
                                        ext_sizes.ctypes.data_as(C.POINTER(C.c_int64)) if ext_sizes is not None else None,
                                        C.c_void_p(ext_addr.ctypes.data) if ext_addr is not None else None,
                                        C.c_void_p(scratch.data_ptr()), scratch.numel(), C.c_void_p(out.data_ptr()), out.numel(),
                                        poff.ctypes.data_as(C.POINTER(C.c_int64)), C.c_void_p(mids.ctypes.data),
                                        sizes.ctypes.data_as(C.POINTER(C.c_int64)) if sizes is not None else None, self._stream()), "kolm_encode_blocks")
        res = (out, poff, mids[:nb])
        return res + (sizes.reshape(-1, ncand)[:nb],) if want_sizes else res

    def encode_blocks_stats(self):
        """kolm_encode_blocks_stats: about the last encode_blocks call of this context"""
        a = (C.c_int64 * 2)()
        _lib.check(_lib.lib().kolm_encode_blocks_stats(self._h, a), "kolm_encode_blocks_stats")
        return dict(repair_stopped_early=int(a[0]))

    def decode_blocks(self, profile: int, payload: torch.Tensor, pay_start, pay_len, method_ids, out_off, out: Optional[torch.Tensor] = None):
        """kolm_decode_blocks: the decode loop of decompress (KF.py:925-949 / V22.py:2530-2540) for a whole batch; block b's payload is
        payload[pay_start[b] : pay_start[b] + pay_len[b]].  Raises KolmError with .block = the lowest failing block index."""
        ps = np.ascontiguousarray(pay_start, dtype=np.int64)
        pl = np.ascontiguousarray(pay_len, dtype=np.int64)
        oa, op = _offsets(out_off)
        nb = len(oa) - 1
        mids = np.ascontiguousarray(method_ids, dtype=np.uint8)
        L = _lib.lib()
        nout = int(oa[-1])
        scratch = self._scratch(L.kolm_decode_blocks_scratch(max(int(pl.sum()), 1), max(nout - int(oa[0]), 1), max(nb, 1)), payload.device)
        if out is None:
            out = torch.empty(max(nout, 1) + 16, dtype=torch.uint8, device=payload.device)
        bad = C.c_int(-1)
        rc = L.kolm_decode_blocks(self._h, int(profile), C.c_void_p(payload.data_ptr()), ps.ctypes.data_as(C.POINTER(C.c_int64)),
                                  pl.ctypes.data_as(C.POINTER(C.c_int64)), C.c_void_p(mids.ctypes.data), op, nb,
                                  C.c_void_p(scratch.data_ptr()), scratch.numel(), C.c_void_p(out.data_ptr()), C.byref(bad), self._stream())
        if rc != 0:
            e = _lib.KolmError(rc, "kolm_decode_blocks")
            e.block = bad.value
            raise e
        return out

    def select_blocks(self, sizes: np.ndarray):
        """_encode_block / the KOLR selection loops: sizes int64[nblocks, ncand] -> (winner index int64[nblocks], its size
        int64[nblocks]); first minimum, i.e. the lowest id on ties (strict '<' in the reference).  Compared on the device."""
        sizes = np.ascontiguousarray(sizes, dtype=np.int64)
        nb, nc = sizes.shape
        mids = np.zeros(nb, dtype=np.int32)
        best = np.zeros(nb, dtype=np.int64)
        if nb:
            _lib.check(_lib.lib().kolm_select_blocks(self._h, sizes.ctypes.data_as(C.POINTER(C.c_int64)), nb, nc,
                                                     mids.ctypes.data_as(C.POINTER(C.c_int32)), best.ctypes.data_as(C.POINTER(C.c_int64)),
                                                     self._stream()), "kolm_select_blocks")
        return mids.astype(np.int64), best

    def gather_payloads(self, src_addr: np.ndarray, lens: np.ndarray, out: torch.Tensor):
        """Winning payloads (device addresses + lengths per block) -> back to back in `out`; returns out_off."""
        nb = len(lens)
        src_addr = np.ascontiguousarray(src_addr, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        out_off = np.zeros(nb + 1, dtype=np.int64)
        _lib.check(_lib.lib().kolm_gather_payloads(self._h, src_addr.ctypes.data_as(C.POINTER(C.c_uint64)), lens.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   nb, C.c_void_p(out.data_ptr()), out_off.ctypes.data_as(C.POINTER(C.c_int64)), self._stream()),
                   "kolm_gather_payloads")
        return out_off

    def copy_blocks(self, src_addr: np.ndarray, dst_addr: np.ndarray, lens: np.ndarray):
        src_addr = np.ascontiguousarray(src_addr, dtype=np.uint64)
        dst_addr = np.ascontiguousarray(dst_addr, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        _lib.check(_lib.lib().kolm_copy_blocks(self._h, src_addr.ctypes.data_as(C.POINTER(C.c_uint64)), dst_addr.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               lens.ctypes.data_as(C.POINTER(C.c_int64)), len(lens), self._stream()), "kolm_copy_blocks")
