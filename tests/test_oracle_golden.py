"""Pin the C++ oracle against the golden vectors produced by the unmodified Python reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib
import json
import os

import pytest

import datasets
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = json.load(open(os.path.join(GOLD, "small.json")))
CASES = datasets.small_cases()


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def same(b, rec):
    return len(b) == rec["len"] and sha(b) == rec["sha256"]


@pytest.mark.parametrize("name", sorted(SMALL))
def test_stage_vectors(name):
    d, g = CASES[name], SMALL[name]
    assert sha(d) == g["input_sha256"]
    assert O.duval(d) == g["lyndon_starts"]
    L = O.bbwt_forward(d)
    assert same(L, g["bbwt"])
    assert O.bbwt_forward_literal(d) == L
    assert O.bbwt_inverse(L) == d
    m = O.mtf_encode(L)
    assert same(m, g["mtf"])
    assert O.mtf_decode(m) == L


@pytest.mark.parametrize("name", sorted(SMALL))
def test_kf_models(name):
    d, g = CASES[name], SMALL[name]["kf"]
    for mid in range(4):
        p = O.encode_model(O.PROFILE_KOLM, mid, d)
        assert same(p, g[str(mid)]), (name, mid)
        assert O.decode_model(O.PROFILE_KOLM, mid, p, len(d)) == d
    _, prm = O.kf_rice_pack(O.mtf_encode(O.bbwt_forward(d)), with_params=True)
    assert {k: int(v) for k, v in prm.items()} == g["m2_meta"]
    mid, payload, sizes = O.encode_block(O.PROFILE_KOLM, d)
    assert mid == g["selected"]
    for tb in (512, 8192):
        blob = O.kf_compress(d, tb)
        assert same(blob, g["container_%d" % tb]), (name, tb)
        assert O.kf_decompress(blob, len(d)) == d


@pytest.mark.parametrize("name", sorted(SMALL))
def test_v22_models(name):
    d, g = CASES[name], SMALL[name]["v22"]
    for mid, nm in enumerate(O.V22_NAMES):
        p = O.encode_model(O.PROFILE_KOLR, mid, d)
        assert same(p, g[nm]), (name, nm)
        if mid == 3 and len(d) % 8:
            with pytest.raises(O.OracleError):       # reference decoder IndexError
                O.decode_model(O.PROFILE_KOLR, mid, p, len(d))
        else:
            assert O.decode_model(O.PROFILE_KOLR, mid, p, len(d)) == d, (name, nm)
    if d:
        assert g["v2_new"] == {"error": "NameError"}     # dead candidate
    mid, payload, sizes = O.encode_block(O.PROFILE_KOLR, d)
    assert sizes == g["sizes"][:10]
    live = [(s, i) for i, s in enumerate(g["sizes"]) if s is not None]
    assert mid == min(live)[1]


@pytest.mark.parametrize("name", sorted(SMALL))
def test_chunking(name):
    d, g = CASES[name], SMALL[name]
    assert [list(x) for x in O.kf_cdc(d, 256, 512, 1024)] == g["kf_cdc_512_bounds"]
    if d:
        assert [list(x) for x in O.v22_cdc(d, 128, 256, 512)] == g["v22"]["cdc_256_bounds"]


def test_gear_tables():
    import struct
    kf = O.gear("kf")
    assert kf[:3] == [2395527356, 355203201, 2773882754]
    assert hashlib.sha256(struct.pack("<256I", *kf)).hexdigest().startswith("c2572028f0243919")
    v = O.gear("v22")
    assert v[:3] == [3836725727, 2937111989, 1130492583]
    assert hashlib.sha256(struct.pack("<256I", *v)).hexdigest().startswith("9876e4fa338f4ad3")
