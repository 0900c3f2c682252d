"""GPU parity: move-to-front encode / decode vs the CPU oracle (bit-exact)."""
import random

import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _cases():
    c = dict(datasets.small_cases())
    c.update(datasets.medium_cases())
    rnd = random.Random(11)
    for i in range(30):
        n = rnd.choice([1, 2, 31, 32, 33, 100, 4095, 4096, 4097, 9000, 20000])
        alpha = rnd.choice([1, 2, 5, 40, 256])
        c["rnd%d" % i] = bytes(rnd.randrange(alpha) for _ in range(n))
    # BWT-like inputs (runs) at several tile counts
    for name in ("checker", "sine", "pattern"):
        c["bbwt_" + name] = O.bbwt_forward(datasets.fixture(name)[:70000])
    return c


def test_mtf_encode_matches_oracle():
    import gpu_util as G
    cases = _cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    out = G.unbatch(G.ctx().mtf_encode(t, off), off)
    for k, b, o in zip(names, blocks, out):
        assert o == O.mtf_encode(b), k


def test_mtf_decode_matches_oracle_and_roundtrips():
    import gpu_util as G
    cases = _cases()
    names = sorted(cases)
    blocks = [O.mtf_encode(cases[k]) for k in names]
    t, off = G.batch(blocks)
    out = G.unbatch(G.ctx().mtf_decode(t, off), off)
    for k, b, o in zip(names, blocks, out):
        assert o == cases[k], k
    # arbitrary index streams (not produced by an encoder) decode like the reference too
    rnd = random.Random(5)
    raw = [bytes(rnd.randrange(256) for _ in range(n)) for n in (1, 33, 5000, 12345)]
    t, off = G.batch(raw)
    out = G.unbatch(G.ctx().mtf_decode(t, off), off)
    for b, o in zip(raw, out):
        assert o == O.mtf_decode(b)
