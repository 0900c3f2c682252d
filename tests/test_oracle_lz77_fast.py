"""The oracle's trigram-chain LZ77 (for long blocks) must equal its literal restatement of encode_model_lz77 / encode_lz77
(KF.py:567-617, V22.py:1686-1763: every distance scanned, nearest among longest, greedy), which the golden vectors pin to the
reference — checked for the three parameter sets of BASELINE cfg 3 wherever the literal scan is affordable.  CPU only."""
import random
import time

from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import datasets
from oracle import oracle as O

PARAMS = ((255, 127), (4096, 0), (65536, 0))


def _cases():
    c = dict(datasets.small_cases())
    rnd = random.Random(77)
    c["text_20k"] = datasets.medium_cases()["text_big"]
    c["sine_24k"] = datasets.fixture("sine")[1000:25000]
    c["pattern_32k"] = datasets.fixture("pattern")[60000:60000 + 32768]
    c["gradient_12k"] = datasets.fixture("gradient")[5000:17000]
    c["checker_33k"] = datasets.fixture("checker")[:33000]
    c["zeros_9k"] = bytes(9001)
    c["period_300"] = bytes(rnd.randrange(256) for _ in range(300)) * 40          # longer than the KF window, shorter than V22's
    c["period_5000"] = bytes(rnd.randrange(4) for _ in range(5000)) * 4           # longer than V22's window
    c["aab"] = b"aab" * 3400
    c["two_byte_tail"] = b"abcabcab"                                             # a 2-byte match at the end stays two literals
    c["overlap"] = b"x" + b"ab" * 700 + b"x" + b"ab" * 300
    for a in (1, 2, 3, 16, 256):
        for n in (1, 2, 3, 4, 7, 100, 3000):
            c["rnd_a%d_n%d" % (a, n)] = bytes(rnd.randrange(a) for _ in range(n))
    c["runs_mixed"] = b"".join(bytes([rnd.randrange(3)]) * rnd.randrange(1, 300) for _ in range(300))
    return c


def test_fast_equals_literal_on_fixed_cases():
    for name, d in sorted(_cases().items()):
        for w, cap in PARAMS:
            assert O.lz77_encode_fast(d, w, cap) == O.lz77_encode(d, w, cap), (name, w, cap)


@settings(max_examples=300, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(st.integers(1, 4), st.lists(st.integers(0, 255), min_size=0, max_size=600), st.integers(0, 5), st.sampled_from(PARAMS + ((7, 5), (300, 3))))
def test_fast_equals_literal_on_generated_blocks(alpha_bits, raw, rep, prm):
    d = bytes(v & ((1 << alpha_bits) - 1) for v in raw)
    d = d + d[len(d) // 3:] * rep                          # far repeats, overlapping candidates, ties between distances
    assert O.lz77_encode_fast(d, *prm) == O.lz77_encode(d, *prm)


def test_fast_is_usable_at_one_mib():
    from kolmogorovlike_datacompressor_b200 import synth
    t = time.time()
    for seg in range(4):                                   # gradient, sine, pattern, checker megabytes of the S2 corpus
        d = synth.s2_mixed(4 << 20)[seg << 20:(seg + 1) << 20].tobytes()
        for w, cap in PARAMS:
            p = O.lz77_encode_fast(d, w, cap)
            assert O.lz77_decode(p, len(d), 0 if cap else w) == d
    assert time.time() - t < 240
