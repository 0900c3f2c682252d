"""The lower bounds behind the early stop of the Re-Pair candidate in kolm_encode_blocks (csrc/repair.cu, k_repair_enc):

    final payload >= 7 + bytes of the rules made so far + sum over the DISTINCT adjacent pairs (x, y) of the current sequence of uleb(x)
                     + (m - 1 - D) / (f - 1)   (m symbols, D distinct pairs, f the largest pair count)
    final payload >= 6 + bytes of the rules made so far + (3 n1 + 4 n2) / f     (n1 / n2 one- / two-byte symbols in the sequence,
                                                                                 f = the largest pair count, f >= 3; blocks <= 8 KiB)

checked round by round on a plain restatement of repair_compress (kolm_final_researched_v2-2.py:1841-1911) whose payload must
equal the oracle's.  CPU only: this is the arithmetic the kernel relies on, not the kernel."""
import random

import pytest

import datasets
from oracle import oracle as O


def _uleb(v):
    return 1 if v < 128 else 2 if v < 16384 else 3 if v < (1 << 21) else 4


def _uleb_bytes(v):
    out = bytearray()
    while v >= 128:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def repair_with_bounds(block: bytes):
    """-> (payload, [pair bound before every round, and before the final break], [(f, f * payload bound 2) before every round])"""
    seq = list(block)
    rules = []
    bounds = []
    bounds2 = []
    while True:
        freq = {}
        for i in range(len(seq) - 1):                       # _count_pairs: overlapping adjacencies
            p = (seq[i], seq[i + 1])
            freq[p] = freq.get(p, 0) + 1
        if seq:
            # the kernel's sharpened form: D distinct pairs need D boundaries; K later rules and a final sequence of F symbols have
            # K + F - 1 of them, every rule shortens the sequence by at most f (the largest count now, which never rises), every
            # boundary costs a byte or more on its left and a rule one more on its right
            f = max(freq.values(), default=0)
            more = -(-(len(seq) - 1 - len(freq)) // (f - 1)) if f >= 2 else 0
            bounds.append(7 + sum(_uleb(a) + _uleb(b) for a, b in rules) + sum(_uleb(x) for x, _ in freq) + more)
        best, bf = None, 1
        for p, f in freq.items():
            if f > bf or (f == bf and best is not None and p < best):
                best, bf = p, f
        if best is None or bf < 2:
            break
        n1 = sum(1 for v in seq if v < 128)
        n2 = len(seq) - n1
        gain = n1 * max(bf - 3, 0) + n2 * (2 * bf - 4)
        bounds2.append((bf, bf * (6 + sum(_uleb(a) + _uleb(b) for a, b in rules) + n1 + 2 * n2) - gain))
        new, out, i, rep = 256 + len(rules), [], 0, 0
        while i < len(seq):
            if i + 1 < len(seq) and (seq[i], seq[i + 1]) == best:
                out.append(new)
                i += 2
                rep += 1
            else:
                out.append(seq[i])
                i += 1
        if rep < 2:
            break
        rules.append(best)
        seq = out
    pay = bytearray(b"RP") + _uleb_bytes(256) + _uleb_bytes(len(rules))
    for a, b in rules:
        pay += _uleb_bytes(a) + _uleb_bytes(b)
    pay += _uleb_bytes(len(seq))
    for s in seq:
        pay += _uleb_bytes(s)
    return bytes(pay), bounds, bounds2


def _cases():
    c = datasets.small_cases()
    rnd = random.Random(11)
    out = [v[:600] for k, v in sorted(c.items()) if v]
    for n in (1, 2, 3, 17, 200, 700):
        out.append(bytes(rnd.getrandbits(8) for _ in range(n)))
        out.append(bytes(rnd.choice(b"ab") for _ in range(n)))
        out.append(bytes(rnd.choice(b"\x80\x81\xfe\xff") for _ in range(n)))
        out.append(bytes((i * 7) & 0xFF for i in range(n)))
        out.append(b"\xaa" * n)
    # compressible structure, where a wrong bound would show first: word soups, periodic strings with defects, nested repeats,
    # runs of one symbol (a == b rounds), high bytes only (two-byte terminals)
    for k in range(40):
        words = [bytes(rnd.getrandbits(8) | (0x80 if k % 3 == 0 else 0) for _ in range(rnd.randint(1, 7))) for _ in range(rnd.randint(2, 10))]
        out.append(b"".join(rnd.choice(words) for _ in range(rnd.randint(5, 300)))[:1200])
    for k in range(20):
        unit = bytes(rnd.getrandbits(8) for _ in range(rnd.randint(1, 12)))
        s = bytearray(unit * rnd.randint(2, 120))[:1000]
        for _ in range(rnd.randint(0, 4)):
            s[rnd.randrange(len(s))] ^= 1 << rnd.randrange(8)
        out.append(bytes(s))
    for k in range(10):
        a, b = bytes([rnd.getrandbits(8)]), bytes([rnd.getrandbits(8)])
        out.append((a * rnd.randint(1, 40) + b * rnd.randint(1, 40)) * rnd.randint(1, 12))
        x = bytes(rnd.getrandbits(8) for _ in range(3))
        for _ in range(rnd.randint(2, 6)):
            x = x + x[:rnd.randint(1, len(x))] + x
        out.append(x[:1500])
    return out


def test_bound_never_exceeds_the_final_payload():
    checked = 0
    for blk in _cases():
        pay, bounds, bounds2 = repair_with_bounds(blk)
        assert pay == O.repair_compress(blk), len(blk)
        assert bounds and all(b <= len(pay) for b in bounds), (len(blk), max(bounds), len(pay))
        assert all(fb <= f * len(pay) for f, fb in bounds2), (len(blk), len(pay))
        assert all(bounds2[i][0] >= bounds2[i + 1][0] for i in range(len(bounds2) - 1))      # the largest count never rises
        checked += len(bounds)
    assert checked > 1000


def test_bound_is_tight_enough_to_matter():
    """on incompressible bytes the bound passes the raw size long before the last round"""
    rnd = random.Random(5)
    blk = bytes(rnd.getrandbits(8) for _ in range(2048))
    pay, bounds, bounds2 = repair_with_bounds(blk)
    first = next(i for i, b in enumerate(bounds) if b >= len(blk))
    assert first == 0 and len(bounds) > 20 and len(pay) > len(blk)
    # 8 KiB of the S3 mix's sine segment (high bytes, few distinct pairs at first): the plain pair bound needs 848 of the 3696
    # rounds; its sharpened form and the byte-count bound under a hundred
    from kolmogorovlike_datacompressor_b200 import synth
    blk = synth.s3_mix(8 << 20)[3 << 20:(3 << 20) + 8192].tobytes()
    pay, bounds, bounds2 = repair_with_bounds(blk)
    assert pay == O.repair_compress(blk) and len(pay) > len(blk)
    assert all(b <= len(pay) for b in bounds) and all(fb <= f * len(pay) for f, fb in bounds2)
    t1 = next((i for i, b in enumerate(bounds) if b >= len(blk)), len(bounds))
    t2 = next((i for i, (f, fb) in enumerate(bounds2) if fb >= f * len(blk)), len(bounds2))
    assert t2 < 100 and t1 < 100 and len(bounds) > 3000, (t1, t2, len(bounds))
