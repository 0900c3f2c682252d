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


def _small_alphabet_cases():
    rnd = random.Random(11)
    c = {}
    text = datasets.medium_cases()["text_big"]
    c["text_20k"] = text
    c["text_70k"] = (text * 4)[:70001]
    c["dna_50k"] = bytes(rnd.choice(b"ACGT") for _ in range(50000))
    c["binary_9k"] = bytes(rnd.randrange(2) for _ in range(9001))
    c["ones_5k"] = b"\x01" * 5000
    c["desc_letters"] = bytes(range(122, 96, -1)) * 200                  # many Lyndon factors (descending runs)
    c["period7"] = b"abcabca" * 3000
    c["two"] = b"ba"
    c["one"] = b"z"
    c["empty"] = b""
    c["six_bit_64"] = bytes(32 + rnd.randrange(64) for _ in range(30000))   # exactly the largest compressed alphabet (6 bits -> h0 = 5)
    return c


def test_bbwt_forward_deep_bootstrap_batch():
    """Batches whose alphabets all fit 6 bits take the deep bootstrap (two 32-bit sorts, k_rerank<2>): ragged batch vs the oracle."""
    import gpu_util as G
    cases = _small_alphabet_cases()
    names = sorted(cases)
    blocks = [cases[k] for k in names]
    t, off = G.batch(blocks)
    out = G.unbatch(G.ctx().bbwt_forward(t, off), off)
    for k, b, o in zip(names, blocks, out):
        assert o == O.bbwt_forward(b), k


def test_bbwt_forward_deep_bootstrap_forced():
    """KOLM_DEEP_BOOT=2 (deep bootstrap for every alphabet, read once per process): the ragged mixed batch in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, KOLM_DEEP_BOOT="2")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(here, "test_gpu_bbwt.py"), "-k",
                        "ragged_batch or single_blocks or fixture_blocks"], env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("knobs", [
    {"KOLM_LOCAL_ROUNDS": "0"},                          # the round-1 engine (gather + global passes every round)
    {"KOLM_LS_DIV": "0"},                                # local refinement rounds to the end, no hand-over to the compacted rounds
    {"KOLM_LS_DIV": "1"},                                # hand-over right after the first local round
    {"KOLM_LYNDON_FAST": "0"},                           # plain-suffix sort (ISA -> Lyndon starts) through the local rounds as well
    {"KOLM_LYNDON_FAST": "0", "KOLM_LOCAL_ROUNDS": "0"},
    {"KOLM_DEEP_BOOT": "0"},                             # first local round at depth 4..6: most records unsettled, many big groups
], ids=lambda k: ",".join("%s=%s" % kv for kv in sorted(k.items())))
def test_sort_engine_variants_agree_with_the_oracle(knobs):
    """Every way through the rotation / suffix sort (knobs are read once per process, hence the child process) on the ragged,
    periodic and fixture batches, Lyndon flags included."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, **knobs)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(here, "test_gpu_bbwt.py"), "-k",
                        "lyndon_flags or ragged_batch or fixture_blocks or deep_bootstrap_batch"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
