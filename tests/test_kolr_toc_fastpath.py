"""KOLR TOC walk (kolm_final_researched_v2-2.py:2451-2530): the vectorised Elias-Fano part of `_parse` against its scalar loops
(which are the reference's, checked against the live reference in test_oracle_vs_reference_live.py) on containers with many
blocks, FIXED and CDC mode, including empty payloads and a damaged TOC."""
import random

import numpy as np
import pytest

from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V


def _container(rnd, nblocks, mode, spread):
    if mode == V.MODE_FIXED:
        size_field = 2048
        lens = [2048] * (nblocks - 1) + [rnd.randrange(1, 2049)]
    else:
        size_field = 8192
        lens = [rnd.randrange(4096, 16385) for _ in range(nblocks)]
    bounds, p = [], 0
    for n in lens:
        bounds.append((p, p + n))
        p += n
    mids = [rnd.randrange(len(V.KOLR_NAMES)) if rnd.random() < 0.3 else 1 for _ in range(nblocks)]
    plens = [rnd.randrange(0, spread) for _ in range(nblocks)]
    area = np.frombuffer(bytes(rnd.randrange(256) for _ in range(sum(plens))), dtype=np.uint8)
    return V._assemble(p, bounds, mode, size_field, encoded=(mids, plens, area)), mids, plens, lens


@pytest.mark.parametrize("mode", [V.MODE_FIXED, V.MODE_CDC])
@pytest.mark.parametrize("nblocks,spread", [(65, 3), (500, 1), (500, 4000), (4096, 300), (20000, 40)])
def test_fast_elias_fano_equals_scalar_walk(mode, nblocks, spread):
    rnd = random.Random(nblocks * 7 + spread + mode)
    blob, mids, plens, lens = _container(rnd, nblocks, mode, spread)
    names, starts, got_plens, orig_lens, total_len, pos = V._parse(blob)
    assert got_plens == plens and orig_lens == lens and total_len == sum(lens) and pos == len(blob)
    assert names == [V.KOLR_NAMES[m] for m in mids]
    assert starts == [starts[0] + int(x) for x in np.concatenate([[0], np.cumsum(plens)[:-1]])]


def test_scalar_bit_reader_matches_big_integer_semantics():
    rnd = random.Random(3)
    buf = bytes(rnd.randrange(256) for _ in range(300))
    v, total = int.from_bytes(buf, "big"), len(buf) * 8
    br, pos = V._BitsIn(buf), 0
    while pos < total - 70:
        k = rnd.randrange(0, 67)
        want = (v >> (total - pos - k)) & ((1 << k) - 1) if k else 0
        assert br.bits(k) == want
        pos += k
        assert br.bit() == (v >> (total - 1 - pos)) & 1
        pos += 1
    br.pos = total - 3
    with pytest.raises(ValueError, match="out of data"):
        br.bits(4)


def test_truncated_toc_raises_like_the_scalar_walk():
    rnd = random.Random(9)
    blob, *_ = _container(rnd, 300, V.MODE_FIXED, 50)
    bad = bytearray(blob)
    # zero the TOC bit string: no upper-bit ones left -> the scalar loops run and fail the reference's way
    hdr = 14
    for i in range(hdr + 8, hdr + 400):
        bad[i] = 0
    with pytest.raises((ValueError, IndexError, KeyError)):
        V._parse(bytes(bad))


class _SlowBits:
    """the TOC bit writer as one growing integer (what _BitWriter.getvalue_bits of the reference amounts to): the linear-time
    writer and its array appends must produce the same bytes"""

    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, val, k):
        if k:
            self.acc = (self.acc << k) | (val & ((1 << k) - 1))
            self.n += k

    def unary(self, q):
        self.acc = (self.acc << (q + 1)) | (((1 << q) - 1) << 1)
        self.n += q + 1

    def rice(self, seq, k):
        for v in seq:
            self.unary(v >> k)
            self.put(v, k)

    def put_array(self, bits):
        for b in np.asarray(bits).tolist():
            self.put(int(b), 1)

    def put_fixed(self, vals, k):
        for v in np.asarray(vals).tolist():
            self.put(int(v), k)

    def value(self):
        nbytes = (self.n + 7) // 8
        return (self.acc << (nbytes * 8 - self.n)).to_bytes(nbytes, "big"), self.n


@pytest.mark.parametrize("mode", [V.MODE_FIXED, V.MODE_CDC])
@pytest.mark.parametrize("nblocks,spread", [(1, 1), (2, 70000), (65, 3), (700, 1), (700, 4000), (3000, 300)])
def test_linear_time_toc_writer_equals_big_integer_writer(mode, nblocks, spread, monkeypatch):
    rnd = random.Random(nblocks * 11 + spread + mode)
    fast, mids, plens, lens = _container(rnd, nblocks, mode, spread)
    rnd = random.Random(nblocks * 11 + spread + mode)
    monkeypatch.setattr(V, "_Bits", _SlowBits)
    slow, mids2, plens2, lens2 = _container(rnd, nblocks, mode, spread)
    assert (mids, plens, lens) == (mids2, plens2, lens2)
    assert fast == slow


def test_bit_writer_pieces():
    rnd = random.Random(9)
    for _ in range(200):
        a, b = _SlowBits(), V._Bits()
        b._FLUSH = rnd.choice([8, 64, 4096])
        for _ in range(rnd.randint(0, 40)):
            op = rnd.randrange(5)
            if op == 0:
                k = rnd.choice([0, 1, 7, 8, 13, 64, 300, 5000]); v = rnd.getrandbits(k + 2)
                a.put(v, k); b.put(v, k)
            elif op == 1:
                q = rnd.choice([0, 1, 9, 70, 9000]); a.unary(q); b.unary(q)
            elif op == 2:
                seq = [rnd.randrange(300) for _ in range(rnd.randrange(20))]; k = rnd.randrange(8)
                a.rice(seq, k); b.rice(seq, k)
            elif op == 3:
                bits = np.array([rnd.randrange(2) for _ in range(rnd.choice([0, 1, 5, 8, 9, 64, 1001]))], dtype=np.uint8)
                a.put_array(bits); b.put_array(bits)
            else:
                k = rnd.choice([0, 1, 5, 13, 40, 62])
                vals = np.array([rnd.getrandbits(max(k, 1)) for _ in range(rnd.randrange(30))], dtype=np.int64)
                a.put_fixed(vals, k); b.put_fixed(vals, k)
        assert a.value() == b.value()
