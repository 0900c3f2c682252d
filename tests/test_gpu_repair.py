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


def _big_cases():
    rnd = random.Random(77)
    text = datasets.medium_cases()["text_big"]
    c = {}
    c["text_20k"] = text
    c["text_50k"] = (text * 3)[:50001]
    c["sine_40k"] = datasets.fixture("sine")[1000:41000]
    c["pattern_64k"] = datasets.fixture("pattern")[60000:60000 + 65536]
    c["gradient_30k"] = datasets.fixture("gradient")[5000:35000]
    c["checker_33k"] = datasets.fixture("checker")[:33000]
    c["zeros_9k"] = bytes(9001)
    c["zeros_even"] = bytes(16384)
    c["aab_10k"] = b"aab" * 3400
    c["abab_9k"] = b"ab" * 4600
    c["alpha2_12k"] = bytes(rnd.randrange(2) for _ in range(12000))
    c["alpha4_10k"] = bytes(rnd.randrange(4) for _ in range(10000))
    c["alpha16_9k"] = bytes(rnd.randrange(16) for _ in range(9000))
    c["random_9k"] = bytes(rnd.randrange(256) for _ in range(9000))
    c["runs_mixed"] = b"".join(bytes([rnd.randrange(3)]) * rnd.randrange(1, 40) for _ in range(900))
    c["aaa_tail"] = b"xy" * 5000 + b"aaa"                       # count 2, one replacement: the reference stops there
    return c


def test_repair_large_blocks_incremental_kernel():
    """Blocks of 9 .. 24 KB: the 16 KiB shape of the shared-memory kernel and, beyond it, the incremental kernel: payloads equal the literal oracle's."""
    import gpu_util as G
    cases = _big_cases()
    cases["small_mixed_in"] = b"abracadabra" * 50               # small and large blocks in one batch
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


def test_repair_one_mib_roundtrip():
    """Full-size property check (the literal oracle needs minutes here): 1 MiB blocks round-trip through the grammar."""
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    d = synth.s3_mix(4 << 20).tobytes()
    blocks = [d[i << 20:(i + 1) << 20] for i in (0, 2, 3)]     # text, gradient, sine
    t, off = G.batch(blocks)
    out, out_off = G.ctx().repair_encode(t, off)
    dec = G.unbatch(G.ctx().repair_decode(out, out_off, off), off)
    assert dec == blocks
    assert all(out_off[i + 1] - out_off[i] < len(blocks[i]) for i in range(3))


def test_kolr_selection_with_repair_ahead_of_the_batches(monkeypatch):
    """Long blocks: the engine runs Re-Pair ahead of its batch loop in groups on a third context.  (method, payload) of every
    block must equal the oracle's selection over all ten candidates — small groups and batches so that groups span batches."""
    import numpy as np
    from kolmogorovlike_datacompressor_b200 import synth
    from kolmogorovlike_datacompressor_b200.engine import Engine
    from kolmogorovlike_datacompressor_b200.kolm_final_researched_v2_2 import KOLR_NAMES
    rnd = random.Random(5)
    bs = 20 << 10                                               # beyond the largest shape of the shared-memory kernel (16 KiB)
    parts = [synth.s1_text(3 * bs).tobytes(), bytes(rnd.randrange(4) for _ in range(bs)), b"abcabcabd" * (2 * bs // 9),
             synth.s2_mixed(2 * bs).tobytes(), bytes(rnd.randrange(256) for _ in range(bs // 2))]
    data = b"".join(parts)
    bounds = [(a, min(a + bs, len(data))) for a in range(0, len(data), bs)]
    eng = Engine(batch_bytes=3 * bs)
    eng.repair_group_bytes = 4 * bs
    names = [n for n in KOLR_NAMES if n != "v2_new"]
    got = eng.encode_kolr(data, bounds, names)
    assert eng.ctx3 is not None                                  # the ahead path ran
    two = eng.encode_kolr(data, bounds, ["raw", "repair"])      # Re-Pair against raw only: its payloads reach the container
    wins = 0
    for k, (a, b) in enumerate(bounds):
        mid, payload, _ = O.encode_block(O.PROFILE_KOLR, data[a:b])
        assert got[k][0] == mid and got[k][1] == payload, k
        mid2, payload2, _ = O.encode_block(O.PROFILE_KOLR, data[a:b], models_mask=(1 << 0) | (1 << 9))
        assert two[k][0] == (1 if mid2 == 9 else 0) and two[k][1] == payload2, k
        wins += mid2 == 9
    assert wins >= 4


def test_kolr_container_where_repair_wins_on_long_blocks():
    """The sine fixture at 32 KiB blocks: Re-Pair (method 9) is the reference's choice on such blocks.  Every block of the drop-in's
    container carries the oracle's method and payload, and the container round-trips."""
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    from kolmogorovlike_datacompressor_b200.engine import KOLR_NAMES
    d = datasets.fixture("sine")[:4 * 32768 + 1000]
    blob = V.compress_blocks_fixed(d, 32768)
    names, starts, plens, orig_lens, total_len, _ = V._parse(blob)
    assert total_len == len(d) and len(names) == 5
    nine = 0
    for k, (nm, st, pl, ol) in enumerate(zip(names, starts, plens, orig_lens)):
        mid, payload, _ = O.encode_block(O.PROFILE_KOLR, d[k * 32768:k * 32768 + ol])
        assert nm == KOLR_NAMES[mid] and blob[st:st + pl] == payload, k
        nine += mid == 9
    assert nine >= 3
    assert V.decompress(blob) == d


def test_repair_one_mib_blocks_equal_the_incremental_oracle():
    """BASELINE's block size: the payloads of six 1 MiB blocks of the S3 mix (text, gradient, sine, patterns, checker, random) equal
    the oracle's incremental Re-Pair byte for byte (tests/test_oracle_repair_fast.py ties that one to the literal restatement)."""
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200 import synth
    mix = synth.s3_mix(8 << 20)
    blocks = [mix[t << 20:(t + 1) << 20].tobytes() for t in (0, 2, 3, 4, 5, 7)]
    t, off = G.batch(blocks)
    out, oo = G.ctx().repair_encode(t, off)
    got = out.cpu().numpy().tobytes()
    for i, b in enumerate(blocks):
        assert got[oo[i]:oo[i + 1]] == O.repair_compress_fast(b), i
