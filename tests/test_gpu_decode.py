"""GPU parity: inverse BBWT and the Rice/gamma bit parsers vs the CPU oracle; full stage round trips."""
import random

import numpy as np
import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu
FLAGS = [0, 1, 4, 8, 16]


def _blocks():
    c = {**datasets.small_cases(), **datasets.medium_cases()}
    rnd = random.Random(9)
    for i in range(20):
        n = rnd.choice([1, 2, 3, 8, 64, 100, 4096, 4097, 9000, 16384])
        alpha = rnd.choice([1, 2, 4, 256])
        c["rnd%d" % i] = bytes(rnd.randrange(alpha) for _ in range(n))
    for name in datasets.FIXTURES:
        c["fx_" + name] = datasets.fixture(name)[1000:66536]
    return c


def test_bbwt_inverse_matches_oracle():
    import gpu_util as G
    cases = _blocks()
    names = sorted(cases)
    fwd = [O.bbwt_forward(cases[k]) for k in names]
    t, off = G.batch(fwd)
    out = G.unbatch(G.ctx().bbwt_inverse(t, off), off)
    for k, o in zip(names, out):
        assert o == cases[k], k
    # arbitrary byte strings are valid BBWT images too (the transform is a bijection)
    raw = [cases[k] for k in names]
    t, off = G.batch(raw)
    out = G.unbatch(G.ctx().bbwt_inverse(t, off), off)
    for k, b, o in zip(names, raw, out):
        assert o == O.bbwt_inverse(b), k


def test_kf_rice_decode_roundtrip():
    import gpu_util as G
    cases = _blocks()
    names = sorted(cases)
    mtf = [O.mtf_encode(O.bbwt_forward(cases[k])) for k in names]
    pays = [O.kf_rice_pack(m) for m in mtf]
    pt, poff = G.batch(pays)
    _, off = G.batch(mtf)
    out = G.unbatch(G.ctx().rice_kf_decode(pt, poff, off), off)
    for k, m, o in zip(names, mtf, out):
        assert o == m, k


@pytest.mark.parametrize("flags", FLAGS)
def test_k2_rice_decode_roundtrip(flags):
    import gpu_util as G
    cases = {k: v for k, v in _blocks().items() if flags != 1 or len(v) % 8 == 0}
    names = sorted(cases)
    mtf = [O.mtf_encode(O.bbwt_forward(cases[k])) for k in names]
    pays = [O.v22_rice_pack(m, flags) for m in mtf]
    pt, poff = G.batch(pays)
    _, off = G.batch(mtf)
    out = G.unbatch(G.ctx().rice_k2_decode(pt, poff, off, flags), off)
    for k, m, o in zip(names, mtf, out):
        assert o == m, (k, flags)


def test_decode_error_codes():
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200._lib import KolmError
    m = O.mtf_encode(O.bbwt_forward(datasets.small_cases()["text"]))
    pay = O.kf_rice_pack(m)
    pt, poff = G.batch([pay[:len(pay) // 2]])
    _, off = G.batch([m])
    with pytest.raises(KolmError) as e:
        G.ctx().rice_kf_decode(pt, poff, off)
    assert e.value.code == -4                      # truncated -> reference EOFError
    odd = O.mtf_encode(O.bbwt_forward(b"abcdefghijk"))   # 11 bytes: bit-plane decode raises IndexError in the reference
    pt, poff = G.batch([O.v22_rice_pack(odd, 1)])
    _, off = G.batch([odd])
    with pytest.raises(KolmError) as e:
        G.ctx().rice_k2_decode(pt, poff, off, 1)
    assert e.value.code == -7


def test_full_stage_roundtrip_1mib_blocks():
    """Size-independent property at the BASELINE block size: decode(encode(x)) == x through every GPU stage."""
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    n = 4 << 20
    data = np.concatenate([synth.s1_text(2 << 20), synth.s2_mixed(2 << 20)])
    off = np.arange(0, n + 1, 1 << 20, dtype=np.int64)
    import torch
    t = torch.from_numpy(data).cuda()
    c = G.ctx()
    L = c.bbwt_forward(t, off)
    m = c.mtf_encode(L, off)
    pay, poff = c.rice_kf_encode(m, off)
    m2 = c.rice_kf_decode(pay, poff, off)
    assert torch.equal(m2[:n], m[:n])
    pay2, poff2, sizes = c.rice_k2_encode(m, off, 16)
    m3 = c.rice_k2_decode(pay2, poff2, off, 16)
    assert torch.equal(m3[:n], m[:n])
    L2 = c.mtf_decode(m2, off)
    assert torch.equal(L2[:n], L[:n])
    x = c.bbwt_inverse(L2, off)
    assert torch.equal(x[:n], t[:n])
    # and the first block agrees with the oracle bit for bit
    blk = data[:1 << 20].tobytes()
    assert L[:1 << 20].cpu().numpy().tobytes() == O.bbwt_forward(blk)
