"""The reference's quirks, one test per item, with expectations DERIVED BY HAND from the cited reference
lines — not produced by any implementation.  Each expectation is checked three ways:
  * against the CPU oracle (always),
  * against the unmodified Python reference when /root/reference is mounted (here; absent on the GPU box),
  * against the CUDA path through the C-ABI / the drop-ins (`-m gpu`).
KF = final/kolm_final.py, V22 = final_researched/kolm_final_researched_v2-2.py."""
import pytest

from oracle import oracle as O
from oracle import ref_loader as R


class Bits:
    """MSB-first bit string builder (KF.py:411-450 BitWriter semantics: zero padded to a byte)."""

    def __init__(self):
        self.s = ""

    def put(self, v, n):
        self.s += format(v, "0%db" % n) if n else ""
        return self

    def ones(self, q):
        self.s += "1" * q
        return self

    def bytes(self):
        s = self.s + "0" * (-len(self.s) % 8)
        return bytes(int(s[i:i + 8], 2) for i in range(0, len(s), 8))


# ---- hand-derived vectors ---------------------------------------------------------------------------------------------------
# KF model 2 token stream (KF.py:636-691): 2 bits (nz<<1 | zero), 4 bits k0, 4 bits k1, then tag + code per token.
# [0]: one zero run r=1.  Rice cost k=0: (1>>0)+1+0 = 2, never below; gamma cost 2*1-1 = 1  ->  2 < 1 false: gamma.  '1'.
KF_ZERO = (bytes([0]), Bits().put(0, 2).put(0, 4).put(0, 4).put(0, 1).put(1, 1).bytes())
# [3]: non-zero x = v-1 = 2.  Rice k=0: 2+1 = 3 (k=1: 1+1+1 = 3, not strictly less): k1 = 0, c1 = 3; gamma codes x+1 = 3 (quirk 3):
# cost 2*2-1 = 3  ->  3 < 3 false: gamma('011').  Were x = 2 coded instead, the code would be '010'.
KF_THREE = (bytes([3]), Bits().put(0, 2).put(0, 4).put(0, 4).put(1, 1).put(0b011, 3).bytes())
# [0,0,0]: r = 3.  Rice k=0: 4, k=1: 1+1+1 = 3, k=2: 0+1+2 = 3 (not strictly less): k0 = 1, c0 = 3; gamma 2*2-1 = 3 -> gamma is
# used, and k0 = 1 is written all the same (quirk 2).
KF_RUN3 = (bytes(3), Bits().put(0, 2).put(1, 4).put(0, 4).put(0, 1).put(0b011, 3).bytes())
# [200]*4: x = 199.  Rice k=6: (199>>6)+1+6 = 10 per token (k=5: 6+1+5 = 12), c1 = 40; gamma of 200: 2*8-1 = 15 each = 60 -> Rice.
# Flags: nz = 1, zero = 0 -> '10' (quirk 1: nz is the HIGH bit).  No zero runs: k0 = 0.  Token: tag 1, 3 ones, 0, low 6 bits = 7.
_b = Bits().put(0b10, 2).put(0, 4).put(6, 4)
for _ in range(4):
    _b.put(1, 1).ones(3).put(0, 1).put(199 & 63, 6)
KF_RICE = (bytes([200]) * 4, _b.bytes())
# [0]*40 + [1]: r = 40: Rice k=3: 5+1+3 = 9 (k=2: 10+3 = 13, k=4: 2+1+4 = 7, k=5: 1+1+5 = 7 not less, k=6: 0+1+6 = 7): k0 = 4, c0 = 7;
# gamma(40) = 2*6-1 = 11 -> Rice for zero runs (flag bit 0).  x = 0: Rice k=0: 0+1 = 1; gamma(x+1 = 1) = 1 -> 1 < 1 false: gamma '1'.
KF_MIXED = (bytes(40) + bytes([1]), Bits().put(0b01, 2).put(4, 4).put(0, 4).put(0, 1).ones(40 >> 4).put(0, 1).put(40 & 15, 4).put(1, 1).put(1, 1).bytes())

# Re-Pair (V22.py:1841-1911).  'cdcdabab': (c,d) and (a,b) both occur twice -> the smaller tuple (97,98) first (quirk 7), then (99,100);
# afterwards Y Y X X has no pair twice.  'RP', ULEB 256, ULEB nrules, rules, ULEB len, symbols.
RP_TIE = (b"cdcdabab", b"RP" + b"\x80\x02" + b"\x02" + bytes([97, 98, 99, 100]) + b"\x04" + b"\x81\x02\x81\x02\x80\x02\x80\x02")
# 'aaa': (a,a) counted twice (overlapping), but only one non-overlapping replacement -> the rule is NOT recorded (V22.py:1880-1882)
RP_DROP = (b"aaa", b"RP" + b"\x80\x02" + b"\x00" + b"\x03" + b"aaa")

# residual coders: KF 'xor' is XOR (KF.py:559), V22 'xor' is a difference mod 256 (V22.py:2109) (quirk 4); ULEB128 of each residual
XOR_IN = bytes([5, 3, 131])
KF_XOR = bytes([5, 5 ^ 3]) + bytes([((3 ^ 131) & 0x7F) | 0x80, 1])          # 3^131 = 128 -> 80 01
V22_DELTA = bytes([5]) + bytes([((3 - 5) & 0x7F) | 0x80, 1]) + bytes([128 & 0x7F | 0x80, 1])   # 254 -> FE 01; 131-3 = 128 -> 80 01

# LZ77 distance 0 (not in Appendix B; found by the round-1 review): KF's decoder has no check, `out[-0]` is out[0] (KF.py:745-752), so
# literal 'a', literal 'b', match(len 3, dist 0) decodes to 'ab' + 'aaa'; with nothing decoded yet out[0] raises IndexError.
# V22's decoder raises ValueError("distance 0") (V22.py:1794-1795).
LZ_DIST0 = (bytes([0, 97, 0, 98, 1, 3, 0]), 5, b"abaaa")
LZ_DIST0_EMPTY = (bytes([1, 3, 0]), 3)

EMPTY_KOLM = bytes.fromhex("4b4f4c4d" "00200000" "0000000000000000" "0000")                 # quirk 10
EMPTY_KOLR = bytes.fromhex("4b4f4c520008000000000000000004000000000000")
ONE_KOLM = bytes.fromhex("4b4f4c4d002000000100000000000000010000010000000100000041")
ONE_KOLR = bytes.fromhex("4b4f4c52000800000100000001000605010101000100014841")


def lz_tokens(p):
    """-> [(len, dist)] of the match tokens of an LZ77 payload (KF.py:600-616 / V22.py:1740-1762: 00 byte | 01 ULEB len ULEB dist)."""
    out, i = [], 0

    def uleb():
        nonlocal i
        v = sh = 0
        while True:
            b = p[i]; i += 1
            v |= (b & 0x7F) << sh; sh += 7
            if b < 128:
                return v
    while i < len(p):
        t = p[i]; i += 1
        if t == 0:
            i += 1
        else:
            assert t == 1
            ln = uleb(); out.append((ln, uleb()))
    return out


KF_CASES = [KF_ZERO, KF_THREE, KF_RUN3, KF_RICE, KF_MIXED]


# ---- oracle -------------------------------------------------------------------------------------------------------------------
def test_oracle_kf_token_stream_quirks_1_2_3():
    for mtf, want in KF_CASES:
        assert O.kf_rice_pack(mtf) == want, mtf
        assert O.kf_rice_unpack(want, len(mtf)) == mtf


def test_oracle_xor_vs_delta_quirk_4():
    assert O.residual_encode(XOR_IN, 0) == KF_XOR
    assert O.residual_encode(XOR_IN, 1) == V22_DELTA


def test_oracle_bitplane_pads_to_groups_of_eight_quirk_5():
    # nine zero MTF bytes: plain Rice(k=2) = 9 * 3 bits -> 4 bytes; the bit-plane variant codes ceil(9/8)*8 = 16 symbols -> 48 bits
    assert O.v22_rice_pack(bytes(9), 0) == bytes(4)
    assert O.v22_rice_pack(bytes(9), 1) == bytes(6)


def _lz_inputs():
    import random
    rnd = random.Random(4)
    a = bytes(rnd.randrange(256) for _ in range(6000))
    return a + a, b"a" * 1000


def test_oracle_lz77_windows_quirk_6():
    far, run = _lz_inputs()
    toks = lz_tokens(O.lz77_encode(far, 4096, 0))
    assert all(d <= 4096 for _, d in toks) and sum(l for l, _ in toks) < 600      # the copy 6000 back is out of reach
    toks = lz_tokens(O.lz77_encode(run, 255, 127))
    assert toks[0] == (127, 1) and max(l for l, _ in toks) == 127                      # literal 'a', then capped overlapping matches
    assert lz_tokens(O.lz77_encode(run, 4096, 0)) == [(999, 1)]                        # V22: unbounded length


def test_oracle_repair_tie_and_dropped_rule_quirk_7():
    for d, want in (RP_TIE, RP_DROP):
        assert O.repair_compress(d) == want
        assert O.repair_decompress(want, len(d)) == d


def test_oracle_selection_ties_quirk_8():
    assert O.encode_block(O.PROFILE_KOLM, b"A")[0] == 0            # raw 1 == xor 1 -> id 0
    mid, _, sizes = O.encode_block(O.PROFILE_KOLR, b"A")
    assert mid == 0 and sizes[0] == sizes[1] == 1


def test_oracle_lz77_distance_zero():
    pay, n, want = LZ_DIST0
    assert O.lz77_decode(pay, n, 0) == want
    with pytest.raises(O.OracleError) as e:
        O.lz77_decode(LZ_DIST0_EMPTY[0], LZ_DIST0_EMPTY[1], 0)
    assert e.value.code == -5                                      # IndexError in the reference
    with pytest.raises(O.OracleError) as e:
        O.lz77_decode(pay, n, 4096)
    assert e.value.code == -2                                      # ValueError in the reference


def test_oracle_empty_and_one_byte_kolm_quirk_10():
    assert O.kf_compress(b"") == EMPTY_KOLM
    assert O.kf_compress(b"A") == ONE_KOLM


# ---- the unmodified reference (only where it is mounted) ----------------------------------------------------------------------
@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted")
def test_reference_agrees_with_the_hand_derived_vectors():
    KF, V = R.load_kf(), R.load_v22()
    for mtf, want in KF_CASES:                                     # the BBWT is a bijection: the block whose transform has this MTF
        blk = KF.bbwt_inverse(KF.mtf_decode(list(mtf)))
        assert KF.mtf_encode(KF.bbwt_forward(blk)) == list(mtf)
        assert KF.encode_model_bbwt_mtf(blk)[0] == want, mtf
        assert KF.decode_model_bbwt_mtf(want, len(mtf)) == blk
    mtf9 = bytes(9)
    blk9 = V.bbwt_inverse(V.mtf_decode(list(mtf9)))
    got = {name: fn(blk9)[0] for fn, name in V._select_encoders() if name in ("bbwt", "bbwt_bp")}
    assert got["bbwt"] == bytes(4) and got["bbwt_bp"] == bytes(6)  # quirk 5
    far, run = _lz_inputs()
    assert all(d <= 4096 for _, d in lz_tokens(V.encode_lz77(far)[0]))
    assert lz_tokens(V.encode_lz77(run)[0]) == [(999, 1)] and lz_tokens(KF.encode_model_lz77(run)[0])[0] == (127, 1)
    assert KF.encode_model_xor(XOR_IN)[0] == KF_XOR
    assert V.encode_xor(XOR_IN)[0] == V22_DELTA
    for d, want in (RP_TIE, RP_DROP):
        assert V.repair_compress(d)[0] == want
    assert KF.compress(b"") == EMPTY_KOLM and KF.compress(b"A") == ONE_KOLM
    assert V.compress_blocks_fixed(b"", 2048) == EMPTY_KOLR and V.compress_blocks_fixed(b"A", 2048) == ONE_KOLR
    assert KF.decode_model_lz77(LZ_DIST0[0], LZ_DIST0[1]) == LZ_DIST0[2]
    with pytest.raises(IndexError):
        KF.decode_model_lz77(*LZ_DIST0_EMPTY)
    with pytest.raises(ValueError):
        V.decode_lz77(LZ_DIST0[0], LZ_DIST0[1])
    assert KF.decompress(ONE_KOLM + b"junk") == b"A"               # quirk 9: KF ignores trailing bytes ...
    with pytest.raises(ValueError):
        V.decompress(ONE_KOLR + b"\0")                              # ... V22 rejects them (V22.py:2547-2549)


# ---- CUDA path ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_stage_quirks():
    import gpu_util as G
    c = G.ctx()
    blocks = [m for m, _ in KF_CASES]
    t, off = G.batch(blocks)
    pay, po = c.rice_kf_encode(t, off)
    got = pay.cpu().numpy().tobytes()
    for i, (_, want) in enumerate(KF_CASES):
        assert got[po[i]:po[i + 1]] == want, i
    assert G.unbatch(c.rice_kf_decode(pay, po, off), off) == blocks
    t, off = G.batch([XOR_IN])
    for kind, want in ((0, KF_XOR), (1, V22_DELTA)):
        p, o = c.residual_encode(t, off, kind)
        assert p.cpu().numpy().tobytes()[:o[1]] == want
    t, off = G.batch([bytes(9)])
    for flags, n in ((0, 4), (1, 6)):
        p, o, _ = c.rice_k2_encode(t, off, flags)
        assert p.cpu().numpy().tobytes()[:o[1]] == bytes(n)
    far, run = _lz_inputs()
    t, off = G.batch([far, run])
    p, o = c.lz77_encode(t, off, 4096, 0)
    b = p.cpu().numpy().tobytes()
    assert b[o[0]:o[1]] == O.lz77_encode(far, 4096, 0) and all(d <= 4096 for _, d in lz_tokens(b[o[0]:o[1]]))
    assert lz_tokens(b[o[1]:o[2]]) == [(999, 1)]
    p, o = c.lz77_encode(t, off, 255, 127)
    b = p.cpu().numpy().tobytes()
    assert lz_tokens(b[o[1]:o[2]])[0] == (127, 1) and max(l for l, _ in lz_tokens(b[o[1]:o[2]])) == 127
    import numpy as np
    import torch
    from kolmogorovlike_datacompressor_b200 import _lib
    pay = torch.from_numpy(np.frombuffer(LZ_DIST0[0] + LZ_DIST0_EMPTY[0], dtype=np.uint8).copy()).cuda()
    po = np.array([0, len(LZ_DIST0[0]), len(LZ_DIST0[0]) + len(LZ_DIST0_EMPTY[0])], dtype=np.int64)
    assert c.lz77_decode(pay, po[:2], np.array([0, 5]), 0).cpu().numpy().tobytes()[:5] == LZ_DIST0[2]
    with pytest.raises(_lib.KolmError) as e:
        c.lz77_decode(pay[po[1]:].clone(), po[1:] - po[1], np.array([0, 3]), 0)
    assert e.value.code == -7                                      # KOLM_E_INDEX -> IndexError in the drop-in
    with pytest.raises(_lib.KolmError) as e:
        c.lz77_decode(pay, po[:2], np.array([0, 5]), 4096)
    assert e.value.code == -5                                      # KOLM_E_CORRUPT -> ValueError
    t, off = G.batch([RP_TIE[0], RP_DROP[0]])
    p, o = c.repair_encode(t, off)
    b = p.cpu().numpy().tobytes()
    assert b[o[0]:o[1]] == RP_TIE[1] and b[o[1]:o[2]] == RP_DROP[1]
    assert G.unbatch(c.repair_decode(p, o, off), off) == [RP_TIE[0], RP_DROP[0]]


@pytest.mark.gpu
def test_gpu_container_quirks():
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    assert KF.compress(b"") == EMPTY_KOLM and KF.compress(b"A") == ONE_KOLM
    assert V.compress_blocks_fixed(b"", 2048) == EMPTY_KOLR and V.compress_blocks_fixed(b"A", 2048) == ONE_KOLR
    assert KF.decompress(EMPTY_KOLM) == b"" and V.decompress(EMPTY_KOLR) == b""
    assert KF.decompress(ONE_KOLM + b"junk") == b"A"               # quirk 9
    with pytest.raises(ValueError):
        V.decompress(ONE_KOLR + b"\0")
    assert KF._encode_block(b"A")[0] == 0                          # quirk 8
