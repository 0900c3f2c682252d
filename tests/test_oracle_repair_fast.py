"""The oracle's incremental Re-Pair (for long blocks) must equal its literal restatement of repair_compress (V22.py:1841-1911),
which the golden vectors pin to the reference — checked wherever the literal one is affordable.  CPU only."""
import random
import time

from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import datasets
from oracle import oracle as O


def _cases():
    c = dict(datasets.small_cases())
    rnd = random.Random(91)
    text = datasets.medium_cases()["text_big"]
    c["text_20k"] = text
    c["sine_24k"] = datasets.fixture("sine")[1000:25000]
    c["pattern_32k"] = datasets.fixture("pattern")[60000:60000 + 32768]
    c["gradient_12k"] = datasets.fixture("gradient")[5000:17000]
    c["checker_33k"] = datasets.fixture("checker")[:33000]
    c["zeros_9k"] = bytes(9001)
    c["zeros_even"] = bytes(16384)
    c["aab"] = b"aab" * 3400
    c["abab"] = b"ab" * 4600
    c["aaa_tail"] = b"xy" * 5000 + b"aaa"
    c["aaaaa"] = b"aaaaa"
    for a in (1, 2, 3, 4, 16, 256):
        for n in (2, 3, 7, 100, 3000):
            c["rnd_a%d_n%d" % (a, n)] = bytes(rnd.randrange(a) for _ in range(n))
    c["runs_mixed"] = b"".join(bytes([rnd.randrange(3)]) * rnd.randrange(1, 40) for _ in range(900))
    return c


def test_fast_equals_literal_on_fixed_cases():
    for name, d in sorted(_cases().items()):
        assert O.repair_compress_fast(d) == O.repair_compress(d), name


@settings(max_examples=300, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.integers(1, 5), st.lists(st.integers(0, 255), min_size=0, max_size=700), st.integers(0, 4))
def test_fast_equals_literal_on_generated_blocks(alpha_bits, raw, rep):
    d = bytes(v & ((1 << alpha_bits) - 1) for v in raw)
    d = d + d[: len(d) // 2] * rep                         # repeats far apart: pairs that come back after replacements
    assert O.repair_compress_fast(d) == O.repair_compress(d)


def test_fast_is_usable_at_one_mib():
    from kolmogorovlike_datacompressor_b200 import synth
    d = synth.s1_text(1 << 20).tobytes()
    t = time.time()
    p = O.repair_compress_fast(d)
    assert time.time() - t < 120
    assert O.repair_decompress(p, len(d)) == d
