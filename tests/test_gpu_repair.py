"""GPU parity: Re-Pair grammar candidate vs the CPU oracle (bit-exact), encode and decode."""
import random

import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _cases():
    c = {k: v for k, v in datasets.small_cases().items()}
    rnd = random.Random(33)
    for i in range(20):
        n = rnd.choice([1, 2, 3, 4, 7, 8, 9, 100, 1000, 2048, 4096, 8191, 8192])
        alpha = rnd.choice([1, 2, 3, 4, 16, 256])
        c["rnd%d" % i] = bytes(rnd.randrange(alpha) for _ in range(n))
    c["aaaa_odd"] = b"a" * 4097
    c["aabaab"] = b"aab" * 1365
    for name in datasets.FIXTURES:
        d = datasets.fixture(name)
        for k in range(3):
            c["fx_%s_%d" % (name, k)] = d[70000 * k + 500:70000 * k + 500 + 2048]
    c["text_8k"] = datasets.medium_cases()["text_big"][:8192]
    return c


def test_repair_encode_decode():
    import gpu_util as G
    cases = _cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    out, out_off = G.ctx().repair_encode(t, off)
    got = out.cpu().numpy().tobytes()
    for i, k in enumerate(names):
        assert got[out_off[i]:out_off[i + 1]] == O.repair_compress(blocks[i]), k
    dec = G.unbatch(G.ctx().repair_decode(out, out_off, off), off)
    for k, b, d in zip(names, blocks, dec):
        assert d == b, k


def test_repair_rejects_large_blocks():
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200._lib import KolmError
    t, off = G.batch([bytes(9000)])
    with pytest.raises(KolmError) as e:
        G.ctx().repair_encode(t, off)
    assert e.value.code == -6
