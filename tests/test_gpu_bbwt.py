"""GPU parity: Lyndon factorisation + BBWT forward vs the CPU oracle (bit-exact)."""
import random

import numpy as np
import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _cases():
    c = dict(datasets.small_cases())
    c.update(datasets.medium_cases())
    rnd = random.Random(7)
    for i in range(40):
        n = rnd.choice([1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 100, 257, 1000, 4095, 4096, 4097, 5000, 9000])
        alpha = rnd.choice([1, 2, 3, 4, 16, 256])
        c["rnd%d" % i] = bytes(rnd.randrange(alpha) for _ in range(n))
    c["zeros_20k"] = bytes(20000)
    c["abab_9001"] = (b"ab" * 4501)[:9001]
    c["aab_period"] = b"aab" * 3000
    return c


def test_lyndon_flags_match_duval():
    import gpu_util as G
    cases = _cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    flags = G.unbatch(G.ctx().duval_lyndon_flags(t, off), off)
    for k, b, f in zip(names, blocks, flags):
        want = bytearray(len(b))
        for s in O.duval(b):
            want[s] = 1
        assert f == bytes(want), k


def test_bbwt_forward_ragged_batch():
    import gpu_util as G
    cases = _cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    out = G.unbatch(G.ctx().bbwt_forward(t, off), off)
    for k, b, o in zip(names, blocks, out):
        assert o == O.bbwt_forward(b), k


def test_bbwt_forward_single_blocks():
    import gpu_util as G
    for k, b in sorted(datasets.small_cases().items()):
        t, off = G.batch([b])
        out = G.unbatch(G.ctx().bbwt_forward(t, off), off)[0]
        assert out == O.bbwt_forward(b), k


def test_bbwt_forward_fixture_blocks_64k():
    import gpu_util as G
    blocks = []
    for name in datasets.FIXTURES:
        d = datasets.fixture(name)
        blocks += [d[i:i + 65536] for i in range(0, min(len(d), 4 * 65536), 65536)]
    t, off = G.batch(blocks)
    out = G.unbatch(G.ctx().bbwt_forward(t, off), off)
    for i, (b, o) in enumerate(zip(blocks, out)):
        assert o == O.bbwt_forward(b), i
