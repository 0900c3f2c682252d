"""Corrupt-payload fuzz of every device decoder: mutated and truncated payloads must come back as one of the
reference's error classes (KOLM_E_TRUNCATED / CORRUPT / INDEX) or as a successful decode — never as a CUDA fault.  Meant to be
run under compute-sanitizer as well (tools/sanitize.sh): a decoder that reads or writes out of bounds on a malformed payload
shows up there even when the call "succeeds"."""
import random

import numpy as np
import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _mutations(payload: bytes, rnd, count):
    out = []
    n = len(payload)
    for _ in range(count):
        b = bytearray(payload)
        kind = rnd.randrange(5)
        if kind == 0 and n:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(n)] ^= 1 << rnd.randrange(8)
        elif kind == 1 and n:
            b = b[:rnd.randrange(n)]
        elif kind == 2 and n:
            i = rnd.randrange(n)
            b[i:i + rnd.randrange(1, 9)] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 9)))
        elif kind == 3:
            b += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 6)))
        else:
            b = bytearray(rnd.randrange(256) for _ in range(rnd.randrange(0, max(2, n))))
        out.append(bytes(b))
    return out


def _run(decode, payloads, olens):
    """decode(payload tensor, pay_off, off) on a batch of independent (payload, orig_len) pairs."""
    import torch
    from kolmogorovlike_datacompressor_b200 import _lib
    po = np.zeros(len(payloads) + 1, dtype=np.int64)
    po[1:] = np.cumsum([len(p) for p in payloads])
    off = np.zeros(len(olens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(olens)
    t = torch.frombuffer(bytearray(b"".join(payloads) + bytes(16)), dtype=torch.uint8).cuda()
    try:
        decode(t, po, off)
    except _lib.KolmError as e:
        assert e.code in (-4, -5, -7), (e.code, str(e))          # truncated / corrupt / index: the reference's exception classes
    torch.cuda.synchronize()                                      # a sticky CUDA fault would surface here


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_every_decoder(seed):
    import gpu_util as G
    c = G.ctx()
    rnd = random.Random(seed)
    cases = datasets.small_cases()
    blocks = [cases[k] for k in ("text", "random_small_alpha", "repetitive_text", "runs18", "byte_counter", "zero_2k", "checker_0", "sine_1000", "banana", "one")]
    decs = []
    for blk in blocks:
        m = O.mtf_encode(O.bbwt_forward(blk))
        n = len(blk)
        decs.append((lambda t, po, off: c.rice_kf_decode(t, po, off), O.kf_rice_pack(m), n))
        for fl in (0, 1, 16):
            decs.append((lambda t, po, off, fl=fl: c.rice_k2_decode(t, po, off, fl), O.v22_rice_pack(m, fl), n))
        decs.append((lambda t, po, off: c.lz77_decode(t, po, off, 0), O.lz77_encode(blk, 255, 127), n))
        decs.append((lambda t, po, off: c.lz77_decode(t, po, off, 4096), O.lz77_encode(blk, 4096, 0), n))
        for kind in (0, 1, 2):
            decs.append((lambda t, po, off, kind=kind: c.residual_decode(t, po, off, kind), O.residual_encode(blk, kind), n))
        decs.append((lambda t, po, off: c.repair_decode(t, po, off), O.repair_compress(blk), n))
    for dec, payload, n in decs:
        muts = _mutations(payload, rnd, 6)
        # one batch per decoder: a bad payload fails the call, so mutated payloads are decoded one at a time and then all together
        for mp in muts[:2]:
            _run(dec, [mp], [n])
        _run(dec, muts, [n] * len(muts))
        # wrong declared length with an intact payload
        _run(dec, [payload], [n + rnd.randrange(1, 50)])
        if n > 1:
            _run(dec, [payload], [rnd.randrange(0, n)])


def test_fuzz_v2new_decoder():
    import gpu_util as G
    rnd = random.Random(5)
    cases = datasets.small_cases()
    c = G.ctx(max_bytes=1 << 22, max_blocks=1 << 10)
    for k in ("text", "zero_2k", "banana"):
        blk = cases[k]
        payload = O.v2new_encode(blk)
        for mp in _mutations(payload, rnd, 8):
            _run(lambda t, po, off: c.v2new_decode(t, po, off), [mp], [len(blk)])


def test_fuzz_containers():
    """The drop-ins on mutated containers: a Python exception of the reference's kinds, or a result; no CUDA fault."""
    import torch
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    rnd = random.Random(9)
    data = datasets.medium_cases()["pattern_mix_24k"][:12000] + datasets.small_cases()["text"]
    for mod, blob in ((KF, KF.compress(data, 2048)), (V, V.compress_blocks_fixed(data, 2048))):
        assert mod.decompress(blob) == data
        for mp in _mutations(blob, rnd, 25):
            try:
                mod.decompress(mp)
            except (ValueError, EOFError, IndexError, AssertionError, KeyError, OverflowError, MemoryError) as e:
                pass
            except Exception as e:                                 # struct.error and friends are what the reference raises as well
                assert type(e).__name__ in ("error", "KolmError") and getattr(e, "code", -5) != -1, repr(e)
            torch.cuda.synchronize()
