"""GPU parity: decode_new_pipeline (method 10, V22.py:1578-1648) through the C-ABI vs the golden vectors of the Python reference
and vs the CPU oracle at larger block sizes; drop-in decoder registry and error behaviour."""
import json
import os
import random

import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "v2new.json")))


def _decode(payloads, lens, ctx):
    import gpu_util as G
    import numpy as np
    pt, poff = G.batch(payloads)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    return G.unbatch(ctx.v2new_decode(pt, poff, off), off)


def test_v2new_decode_golden_batch():
    import gpu_util as G
    names = sorted(GOLD)
    pays = [bytes.fromhex(GOLD[k]["payload_hex"]) for k in names]
    want = [bytes.fromhex(GOLD[k]["input_hex"]) for k in names]
    got = _decode(pays, [len(w) for w in want], G.ctx())
    for k, w, g in zip(names, want, got):
        assert g == w, k


def test_v2new_decode_large_blocks_vs_oracle():
    import gpu_util as G
    rnd = random.Random(5)
    text = datasets.medium_cases()["text_big"]
    blocks = [(text * 8)[:65536], datasets.fixture("sine")[:50001], datasets.fixture("gradient")[3000:3000 + 40000],
              bytes(rnd.randrange(256) for _ in range(4097)), b"", b"x", datasets.fixture("pattern")[60000:60000 + 70000],
              bytes(20000), datasets.fixture("checker")[:30000]]
    force = [None, (2, 0), (5, 1), (4, 0), None, None, (3, 0), (1, 4), (2, 3)]
    pays = [O.v2new_encode(b, force=f) for b, f in zip(blocks, force)]
    got = _decode(pays, [len(b) for b in blocks], G.ctx())
    for i, (b, g) in enumerate(zip(blocks, got)):
        assert g == b, i
    # every model, forced, on one 1 MiB-class batch of unaligned blocks
    forced = [(1, 1), (1, 2), (1, 3), (1, 4), (1, 9), (2, 0), (2, 1), (2, 2), (2, 3), (3, 0), (4, 0), (5, 0), (5, 1), (0, 0), (6, 0)]
    blocks = [(text * 3)[7 * i:7 * i + 30011 + i] for i in range(len(forced))]
    pays = [O.v2new_encode(b, force=f) for b, f in zip(blocks, forced)]
    got = _decode(pays, [len(b) for b in blocks], G.ctx())
    for f, b, g in zip(forced, blocks, got):
        assert g == b, f


def test_v2new_decode_errors():
    import gpu_util as G
    from kolmogorovlike_datacompressor_b200._lib import KolmError
    v = GOLD["text"]
    data, pay = bytes.fromhex(v["input_hex"]), bytes.fromhex(v["payload_hex"])
    for bad in (pay[:2], pay[:len(pay) // 2], bytes([pay[0] | 7]) + pay[1:]):
        with pytest.raises(KolmError) as e:
            _decode([bad], [len(data)], G.ctx())
        assert e.value.code == -5


def test_v2new_dropin_decoder_and_container():
    """_select_decoders()[10] decodes reference payloads; a KOLR container whose TOC names method 10 decompresses."""
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    dec = V._select_decoders()[10]
    for k in ("text", "sine_1000", "forced_m4_p0_runs18_1000" if "forced_m4_p0_runs18_1000" in GOLD else "fib"):
        v = GOLD[k]
        assert dec(bytes.fromhex(v["payload_hex"]), v["orig_len"], None) == bytes.fromhex(v["input_hex"]), k
    with pytest.raises(ValueError):
        dec(bytes.fromhex(GOLD["text"]["payload_hex"])[:7], GOLD["text"]["orig_len"], None)
    # container: let the drop-in assemble with candidates restricted to v2_new payloads made by the oracle
    data = (datasets.medium_cases()["text_big"] * 2)[:5000]
    eng = V._engine()
    real = eng.encode_kolr_area
    try:
        import numpy as np

        def fake(d, bounds, names):
            pays = [O.v2new_encode(d[a:b]) for a, b in bounds]
            return np.full(len(bounds), 10), np.array([len(p) for p in pays]), np.frombuffer(b"".join(pays), dtype=np.uint8)
        eng.encode_kolr_area = fake
        blob = V.compress_blocks_fixed(data, 2048)
    finally:
        eng.encode_kolr_area = real
    assert V.decompress(blob) == data


def test_v2new_encode_golden_and_large():
    """kolm_v2new_enc == the reference's encode_new_pipeline (parallel=False) on the golden inputs, == the oracle on large blocks."""
    import gpu_util as G
    names = sorted(k for k, v in GOLD.items() if not v["forced"])
    blocks = [bytes.fromhex(GOLD[k]["input_hex"]) for k in names]
    t, off = G.batch(blocks)
    out, out_off = G.ctx().v2new_encode(t, off)
    got = out.cpu().numpy().tobytes()
    for i, k in enumerate(names):
        assert got[out_off[i]:out_off[i + 1]] == bytes.fromhex(GOLD[k]["payload_hex"]), k
    rnd = random.Random(9)
    text = datasets.medium_cases()["text_big"]
    blocks = [(text * 8)[:65536], datasets.fixture("sine")[:50001], datasets.fixture("gradient")[3000:3000 + 40000],
              bytes(rnd.randrange(256) for _ in range(4097)), b"", b"x", datasets.fixture("pattern")[60000:60000 + 70000],
              bytes(20000), datasets.fixture("checker")[:30000], bytes(rnd.randrange(4) for _ in range(12345)), b"ab" * 4097]
    t, off = G.batch(blocks)
    out, out_off = G.ctx().v2new_encode(t, off)
    got = out.cpu().numpy().tobytes()
    for i, b in enumerate(blocks):
        assert got[out_off[i]:out_off[i + 1]] == O.v2new_encode(b), i
    dec = G.unbatch(G.ctx().v2new_decode(out, out_off, off), off)
    assert dec == blocks


def test_v2new_dropin_opt_in():
    """G_ENABLE_V2_NEW: off = the shipped reference (NameError swallowed, never method 10); on = candidate 10 competes."""
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    data = bytes.fromhex(GOLD["sine_1000"]["input_hex"])
    enc = dict((n, f) for f, n in V._select_encoders())["v2_new"]
    with pytest.raises(NameError):
        enc(data)
    ref_blob = V.compress_blocks_fixed(data, 2048)
    V.G_ENABLE_V2_NEW = True
    try:
        payload, _ = enc(data)
        assert payload == bytes.fromhex(GOLD["sine_1000"]["payload_hex"])
        # full selection with candidate 10 in play: winners == first minimum over the oracle's eleven sizes; container round-trips
        big = datasets.fixture("sine")[2000:2000 + 6 * 2048] + datasets.fixture("gradient")[5000:5000 + 3 * 2048] + bytes(range(256)) * 16
        bounds = V.fixed_boundaries(big, 2048)
        mids, lens, _ = V._engine().encode_kolr_area(big, bounds, V._candidate_names())
        for (a, b), mid, ln in zip(bounds, mids, lens):
            sizes = [len(O.encode_model(2, m, big[a:b])) for m in range(10)] + [len(O.v2new_encode(big[a:b]))]
            assert (int(mid), int(ln)) == (sizes.index(min(sizes)), min(sizes)), (a, sizes)
        assert 10 in set(int(m) for m in mids)
        blob = V.compress_blocks_fixed(big, 2048)
        assert V.decompress(blob) == big
    finally:
        V.G_ENABLE_V2_NEW = False
    assert V.compress_blocks_fixed(data, 2048) == ref_blob
