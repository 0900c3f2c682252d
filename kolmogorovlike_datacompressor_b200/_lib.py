"""ctypes loader of libkolm_b200.so (the C-ABI in include/kolm_abi.h).  Fails loudly when the CUDA
library is missing: there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

KOLM_PROFILE_KOLM = 1
KOLM_PROFILE_KOLR = 2

EXPORTS = [
    "kolm_abi_version", "kolm_strerror", "kolm_last_cuda_error", "kolm_scratch_bytes", "kolm_create", "kolm_create_ex", "kolm_destroy",
    "kolm_lyndon", "kolm_bbwt_fwd", "kolm_bbwt_inv", "kolm_mtf_enc", "kolm_mtf_dec",
    "kolm_rice_kf_enc", "kolm_rice_kf_dec", "kolm_rice_k2_enc", "kolm_rice_k2_dec", "kolm_last_counters",
    "kolm_lz77_enc", "kolm_lz77_dec", "kolm_residual_sizes", "kolm_residual_enc", "kolm_residual_dec",
    "kolm_repair_enc", "kolm_repair_dec", "kolm_repair_max_block", "kolm_v2new_enc", "kolm_v2new_dec", "kolm_cdc_kf", "kolm_cdc_v22", "kolm_cdc_candidates", "kolm_cdc_walk_kf", "kolm_cdc_walk_v22", "kolm_select_blocks", "kolm_gather_payloads", "kolm_copy_blocks",
    "kolm_rice_dual_enc", "kolm_encode_blocks_scratch", "kolm_encode_blocks", "kolm_encode_blocks_stats", "kolm_decode_blocks_scratch", "kolm_decode_blocks",
    "kolm_profile_categories", "kolm_profile_name", "kolm_profile_enable", "kolm_profile_reset", "kolm_profile_read",
]


class KolmError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = lib().kolm_strerror(code).decode()
        if code == -1:
            msg += ": " + lib().kolm_last_cuda_error().decode()
        super().__init__(f"libkolm_b200: {msg} ({code}) {detail}")


def so_path() -> str:
    return _build.SO


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = so_path()
    if not os.path.exists(path) or _build.stale():
        try:
            _build.build()
        except Exception as e:  # no nvcc / compile error: refuse to run
            if not os.path.exists(path):
                raise RuntimeError(f"libkolm_b200.so is missing and could not be built ({e}); "
                                   "the GPU path has no fallback — run `python -m kolmogorovlike_datacompressor_b200.build`") from e
    L = C.CDLL(path)
    L.kolm_strerror.restype = C.c_char_p
    L.kolm_last_cuda_error.restype = C.c_char_p
    L.kolm_scratch_bytes.restype = C.c_size_t
    L.kolm_scratch_bytes.argtypes = [C.c_size_t, C.c_int]
    L.kolm_create.argtypes = [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.kolm_create_ex.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_uint, C.POINTER(C.c_void_p)]
    L.kolm_destroy.argtypes = [C.c_void_p]
    p, i64p, ip = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)
    L.kolm_lyndon.argtypes = [p, p, i64p, C.c_int, p, p]
    for f in ("kolm_bbwt_fwd", "kolm_bbwt_inv", "kolm_mtf_enc", "kolm_mtf_dec"):
        getattr(L, f).argtypes = [p, p, i64p, C.c_int, p, p]
    L.kolm_rice_kf_enc.argtypes = [p, p, i64p, C.c_int, p, C.c_size_t, i64p, ip, p]
    L.kolm_rice_kf_dec.argtypes = [p, p, i64p, i64p, C.c_int, p, p]
    L.kolm_rice_k2_enc.argtypes = [p, p, i64p, C.c_int, C.c_int, p, C.c_size_t, i64p, i64p, p]
    L.kolm_rice_k2_dec.argtypes = [p, p, i64p, i64p, C.c_int, C.c_int, p, p]
    L.kolm_encode_blocks_scratch.restype = C.c_size_t
    L.kolm_encode_blocks_scratch.argtypes = [C.c_int, C.c_size_t, C.c_int]
    L.kolm_encode_blocks.argtypes = [p, C.c_int, p, i64p, C.c_int, C.c_uint32, C.c_int, i64p, p, p, C.c_size_t, p, C.c_size_t, i64p, p, i64p, p]
    L.kolm_decode_blocks_scratch.restype = C.c_size_t
    L.kolm_decode_blocks_scratch.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
    L.kolm_decode_blocks.argtypes = [p, C.c_int, p, i64p, i64p, p, i64p, C.c_int, p, C.c_size_t, p, ip, p]
    L.kolm_rice_dual_enc.argtypes = [p, p, i64p, C.c_int, C.c_int, p, C.c_size_t, i64p, ip, p, C.c_size_t, i64p, i64p, p]
    L.kolm_last_counters.argtypes = [p, i64p]
    L.kolm_encode_blocks_stats.argtypes = [p, i64p]
    L.kolm_lz77_enc.argtypes = [p, p, i64p, C.c_int, C.c_uint32, C.c_uint32, p, C.c_size_t, i64p, p]
    L.kolm_lz77_dec.argtypes = [p, p, i64p, i64p, C.c_int, C.c_uint32, p, p]
    for f in ("kolm_cdc_kf", "kolm_cdc_v22"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, i64p, C.c_int64]
    L.kolm_cdc_candidates.argtypes = [p, p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64, p, C.c_int64, i64p, p]
    for f in ("kolm_cdc_walk_kf", "kolm_cdc_walk_v22"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, p, C.c_int64, i64p, C.c_int64]
    L.kolm_select_blocks.argtypes = [p, i64p, C.c_int, C.c_int, C.POINTER(C.c_int32), i64p, p]
    L.kolm_gather_payloads.argtypes = [p, C.POINTER(C.c_uint64), i64p, C.c_int, p, i64p, p]
    L.kolm_copy_blocks.argtypes = [p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), i64p, C.c_int, p]
    L.kolm_repair_enc.argtypes = [p, p, i64p, C.c_int, p, C.c_size_t, i64p, p]
    L.kolm_repair_dec.argtypes = [p, p, i64p, i64p, C.c_int, p, p]
    L.kolm_v2new_dec.argtypes = [p, p, i64p, i64p, C.c_int, p, p]
    L.kolm_v2new_enc.argtypes = [p, p, i64p, C.c_int, p, C.c_size_t, i64p, p]
    L.kolm_residual_sizes.argtypes = [p, p, i64p, C.c_int, i64p, p]
    L.kolm_residual_enc.argtypes = [p, p, i64p, C.c_int, C.c_int, p, C.c_size_t, i64p, p]
    L.kolm_residual_dec.argtypes = [p, p, i64p, i64p, C.c_int, C.c_int, p, p]
    L.kolm_profile_name.restype = C.c_char_p
    L.kolm_profile_name.argtypes = [C.c_int]
    L.kolm_profile_enable.argtypes = [p, C.c_int]
    L.kolm_profile_reset.argtypes = [p]
    L.kolm_profile_read.argtypes = [p, C.POINTER(C.c_double), i64p, i64p]
    _LIB = L
    return L


def check(code: int, detail: str = ""):
    if code != 0:
        raise KolmError(code, detail)
