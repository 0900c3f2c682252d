"""Host logic of the GPU-assisted chunker: `kolm_cdc_walk_kf/_v22` over a candidate list must reproduce
`cdc_fast_boundaries` (kolm_final.py:161-194) / `cdc_fast_boundaries_strict` (kolm_final_researched_v2-2.py:210-309) exactly.
The candidate list is emulated here with numpy (window hash of the last 32 bytes at every position), so the test needs no GPU;
tests/test_gpu_cdc.py feeds the same walkers from the CUDA kernel."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from kolmogorovlike_datacompressor_b200 import _lib, synth


def window_candidates(data: bytes, which: str, avg: int) -> np.ndarray:
    g = np.array(O.gear(which), dtype=np.uint64)
    a = np.frombuffer(data, dtype=np.uint8)
    n = a.size
    h = np.zeros(n, dtype=np.uint64)
    gv = g[a]
    for d in range(32):                                      # h[p] = sum_d G[data[p-d]] << d  (mod 2^32)
        h[d:] += gv[:n - d] << np.uint64(d)
    h &= np.uint64(0xFFFFFFFF)
    k = max(6, min(20, int(avg).bit_length() - 1))
    if which == "kf":
        ml = ms = (1 << k) - 1
    else:
        ks, kl = (k + 2 if k + 2 <= 20 else 20), (k - 2 if k > 2 else 1)
        ms, ml = (1 << ks) - 1, (1 << kl) - 1
    pos = np.nonzero((h & np.uint64(ml)) == 0)[0].astype(np.uint64)
    strict = ((h[pos.astype(np.int64)] & np.uint64(ms)) == 0).astype(np.uint64)
    out = (pos << np.uint64(1)) | strict
    rng = np.random.default_rng(1)
    rng.shuffle(out)                                         # the kernel reports candidates in arbitrary order
    return np.ascontiguousarray(out)


def walk(which, data, mn, avg, mx):
    L = _lib.lib()
    cand = window_candidates(data, which, avg)
    cap = len(data) // max(1, mn) + 4
    ends = np.zeros(cap, dtype=np.int64)
    r = getattr(L, "kolm_cdc_walk_" + which)(C.cast(C.c_char_p(data), C.c_void_p), len(data), mn, avg, mx, C.c_void_p(cand.ctypes.data), cand.size,
                                             ends.ctypes.data_as(C.POINTER(C.c_int64)), cap)
    assert r >= 0, r
    return ends[:r].tolist()


def corpora():
    rng = np.random.default_rng(7)
    yield "text", synth.s1_text(300000, seed=3).tobytes()
    yield "mixed", synth.s2_mixed(1 << 20).tobytes()[200000:700000]
    yield "random", rng.integers(0, 256, 250000, dtype=np.uint8).tobytes()
    yield "zeros", bytes(100000)
    yield "short", b"abc" * 50
    yield "period", (b"0123456789abcdef" * 9000)[:131077]


@pytest.mark.parametrize("params", [(32, 64, 128), (64, 64, 64), (100, 300, 1000), (512, 1024, 2048), (4096, 8192, 16384), (1, 8192, 100000)])
def test_walkers_match_the_host_scan_and_the_oracle(params):
    mn, avg, mx = params
    for name, data in corpora():
        assert walk("kf", data, mn, avg, mx) == [e for _, e in O.kf_cdc(data, mn, avg, mx)], (name, "kf")
        assert walk("v22", data, mn, avg, mx) == [e for _, e in O.v22_cdc(data, mn, avg, mx)], (name, "v22")
