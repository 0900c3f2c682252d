"""ctypes binding of oracle/_build/libkolm_oracle.so — TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package.  bytes in, bytes out.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkolm_oracle.so")
_lib = None

PROFILE_KOLM = 1
PROFILE_KOLR = 2
RES_XOR, RES_DELTA, RES_LFSR = 0, 1, 2
V22_NAMES = ["raw", "xor", "bbwt", "bbwt_bp", "bbwt_nib", "bbwt_br", "bbwt_gray", "lz77", "lfsr_pred", "repair"]


class OracleError(ValueError):
    def __init__(self, code):
        super().__init__({-1: "truncated", -2: "bad value", -3: "capacity", -4: "bad argument", -5: "index"}.get(code, str(code)))
        self.code = code


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kolm_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        for name in dir_exports():
            getattr(_lib, name).restype = C.c_int64
    return _lib


def dir_exports():
    return ["ko_duval", "ko_bbwt_forward_literal", "ko_bbwt_forward", "ko_bbwt_inverse", "ko_mtf_encode", "ko_mtf_decode",
            "ko_kf_rice_pack", "ko_kf_rice_unpack", "ko_v22_rice_pack", "ko_v22_rice_unpack", "ko_lz77_encode", "ko_lz77_decode",
            "ko_residual_encode", "ko_residual_decode", "ko_repair_compress", "ko_repair_decompress", "ko_kf_cdc", "ko_v22_cdc",
            "ko_lz77_encode_fast", "ko_encode_block_fast", "ko_v2new_encode", "ko_v2new_decode", "ko_v2new_encode_forced", "ko_encode_model", "ko_decode_model", "ko_encode_block", "ko_kf_compress", "ko_kf_decompress"]


def _buf(n):
    return (C.c_uint8 * max(1, n))()


def _in(b):
    b = bytes(b)
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b or b"\0"), len(b)


def _chk(r):
    if r < 0:
        raise OracleError(r)
    return r


def _n2n(fn, data):
    p, n = _in(data)
    out = _buf(n)
    _chk(getattr(lib(), fn)(p, C.c_int64(n), out))
    return bytes(out[:n])


def duval(data):
    p, n = _in(data)
    st = (C.c_uint32 * max(1, n))()
    k = _chk(lib().ko_duval(p, C.c_int64(n), st))
    return list(st[:k])


def bbwt_forward(data): return _n2n("ko_bbwt_forward", data)
def bbwt_forward_literal(data): return _n2n("ko_bbwt_forward_literal", data)
def bbwt_inverse(data): return _n2n("ko_bbwt_inverse", data)
def mtf_encode(data): return _n2n("ko_mtf_encode", data)
def mtf_decode(data): return _n2n("ko_mtf_decode", data)


def kf_rice_pack(mtf, with_params=False):
    p, n = _in(mtf)
    cap = 2 * n + 64 + n * 0  # gamma/rice worst case is bounded by choose-min; keep generous
    cap = max(cap, 8 * n + 64)
    out = _buf(cap)
    prm = (C.c_int * 4)()
    r = _chk(lib().ko_kf_rice_pack(p, C.c_int64(n), out, C.c_int64(cap), prm))
    res = bytes(out[:r])
    return (res, dict(k0=prm[0], k1=prm[1], use_rice_zero=bool(prm[2]), use_rice_nz=bool(prm[3]))) if with_params else res


def kf_rice_unpack(payload, orig_len):
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_kf_rice_unpack(p, C.c_int64(n), C.c_int64(orig_len), out))
    return bytes(out[:orig_len])


def v22_rice_pack(mtf, flags):
    p, n = _in(mtf)
    cap = 9 * n + 64
    out = _buf(cap)
    r = _chk(lib().ko_v22_rice_pack(p, C.c_int64(n), C.c_int(flags), out, C.c_int64(cap)))
    return bytes(out[:r])


def v22_rice_unpack(payload, flags, orig_len):
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_v22_rice_unpack(p, C.c_int64(n), C.c_int(flags), C.c_int64(orig_len), out))
    return bytes(out[:orig_len])


def lz77_encode(data, window, cap_len):
    p, n = _in(data)
    cap = 2 * n + 16
    out = _buf(cap)
    r = _chk(lib().ko_lz77_encode(p, C.c_int64(n), C.c_uint32(window), C.c_uint32(cap_len), out, C.c_int64(cap)))
    return bytes(out[:r])


def lz77_encode_fast(data, window, cap_len):
    """lz77_encode through trigram chains (exact; for long blocks).  tests/test_oracle_lz77_fast.py ties it to lz77_encode."""
    p, n = _in(data)
    cap = 2 * n + 16
    out = _buf(cap)
    r = _chk(lib().ko_lz77_encode_fast(p, C.c_int64(n), C.c_uint32(window), C.c_uint32(cap_len), out, C.c_int64(cap)))
    return bytes(out[:r])


def lz77_decode(payload, orig_len, window_check=0):
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_lz77_decode(p, C.c_int64(n), C.c_int64(orig_len), C.c_uint32(window_check), out))
    return bytes(out[:orig_len])


def residual_encode(data, kind):
    p, n = _in(data)
    cap = 2 * n + 16
    out = _buf(cap)
    r = _chk(lib().ko_residual_encode(p, C.c_int64(n), C.c_int(kind), out, C.c_int64(cap)))
    return bytes(out[:r])


def residual_decode(payload, orig_len, kind):
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_residual_decode(p, C.c_int64(n), C.c_int64(orig_len), C.c_int(kind), out))
    return bytes(out[:orig_len])


def repair_compress(data):
    p, n = _in(data)
    cap = 3 * n + 64
    out = _buf(cap)
    r = _chk(lib().ko_repair_compress(p, C.c_int64(n), out, C.c_int64(cap)))
    return bytes(out[:r])


def repair_compress_fast(data):
    """repair_compress with incrementally kept counts: for long blocks, where the literal recount needs minutes.  The tests
    require it to equal repair_compress wherever that one is affordable."""
    p, n = _in(data)
    cap = 4 * n + 64
    out = _buf(cap)
    r = _chk(lib().ko_repair_compress_fast(p, C.c_int64(n), out, C.c_int64(cap)))
    return bytes(out[:r])


def repair_decompress(payload, orig_len):
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_repair_decompress(p, C.c_int64(n), C.c_int64(orig_len), out))
    return bytes(out[:orig_len])


def v2new_encode(data, force=None):
    """encode_new_pipeline with circuit_map_automaton_forward(parallel=False) (V22.py:1498-1576).
    force=(mode, param): skip the model search and use that model (test hook, mirrors make_golden_v2new.py)."""
    p, n = _in(data)
    cap = 2 * n + 64
    out = _buf(cap)
    if force is None:
        r = _chk(lib().ko_v2new_encode(p, C.c_int64(n), out, C.c_int64(cap)))
    else:
        r = _chk(lib().ko_v2new_encode_forced(p, C.c_int64(n), C.c_int(force[0]), C.c_uint32(force[1]), out, C.c_int64(cap)))
    return bytes(out[:r])


def v2new_decode(payload, orig_len):
    """decode_new_pipeline (V22.py:1578-1648)."""
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_v2new_decode(p, C.c_int64(n), C.c_int64(orig_len), out))
    return bytes(out[:orig_len])


def gear(which):
    g = (C.c_uint32 * 256)()
    getattr(lib(), "ko_%s_gear" % which)(g)
    return list(g)


def _cdc(fn, data, mn, avg, mx):
    p, n = _in(data)
    cap = n // max(1, mn) + 4
    ends = (C.c_int64 * cap)()
    k = _chk(getattr(lib(), fn)(p, C.c_int64(n), C.c_int64(mn), C.c_int64(avg), C.c_int64(mx), ends, C.c_int64(cap)))
    out, a = [], 0
    for e in ends[:k]:
        out.append((a, e))
        a = e
    return out


def kf_cdc(data, mn, avg, mx): return _cdc("ko_kf_cdc", data, mn, avg, mx)
def v22_cdc(data, mn, avg, mx): return _cdc("ko_v22_cdc", data, mn, avg, mx)


def encode_model(profile, mid, data):
    p, n = _in(data)
    cap = 9 * n + 64
    out = _buf(cap)
    r = _chk(lib().ko_encode_model(C.c_int(profile), C.c_int(mid), p, C.c_int64(n), out, C.c_int64(cap)))
    return bytes(out[:r])


def decode_model(profile, mid, payload, orig_len):
    p, n = _in(payload)
    out = _buf(orig_len)
    _chk(lib().ko_decode_model(C.c_int(profile), C.c_int(mid), p, C.c_int64(n), C.c_int64(orig_len), out))
    return bytes(out[:orig_len])


def encode_block(profile, data, models_mask=None, fast=False):
    """-> (method, payload, sizes[list]).  fast: LZ77 and Re-Pair through lz77_encode_fast / repair_compress_fast (long blocks)."""
    nm = 4 if profile == PROFILE_KOLM else 10
    if models_mask is None:
        models_mask = (1 << nm) - 1
    p, n = _in(data)
    cap = 9 * n + 64
    out = _buf(cap)
    method = C.c_int()
    sizes = (C.c_int64 * 16)()
    r = _chk((lib().ko_encode_block_fast if fast else lib().ko_encode_block)(C.c_int(profile), p, C.c_int64(n), C.c_uint32(models_mask), out, C.c_int64(cap), C.byref(method), sizes))
    return method.value, bytes(out[:r]), [s if s >= 0 else None for s in sizes[:nm]]


def kf_compress(data, target_block=8192):
    p, n = _in(data)
    cap = 2 * n + 9 * (n // max(1, target_block // 2) + 2) + 64
    out = _buf(cap)
    r = _chk(lib().ko_kf_compress(p, C.c_int64(n), C.c_uint32(target_block), out, C.c_int64(cap)))
    return bytes(out[:r])


def kf_decompress(blob, cap):
    p, n = _in(blob)
    out = _buf(cap)
    r = _chk(lib().ko_kf_decompress(p, C.c_int64(n), out, C.c_int64(cap)))
    return bytes(out[:r])
