"""GPU parity at the drop-in boundary: byte-identical KOLM / KOLR containers vs the golden vectors produced by the
unmodified Python reference, decode round trips, error behaviour."""
import hashlib
import json
import lzma
import os

import pytest

import datasets

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = json.load(open(os.path.join(GOLD, "small.json")))
MEDIUM = json.load(open(os.path.join(GOLD, "medium.json"))) if os.path.exists(os.path.join(GOLD, "medium.json")) else {}


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def same(b, rec):
    return len(b) == rec["len"] and sha(b) == rec["sha256"]


def test_kolm_small_containers():
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    for name, d in sorted(datasets.small_cases().items()):
        g = SMALL[name]["kf"]
        for tb in (512, 8192):
            blob = KF.compress(d, target_block=tb)
            assert same(blob, g["container_%d" % tb]), (name, tb)
            assert KF.decompress(blob) == d, (name, tb)
        if d:
            mid, payload, plen = KF._encode_block(d)
            assert mid == g["selected"], name


def test_kolr_small_containers():
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    for name, d in sorted(datasets.small_cases().items()):
        g = SMALL[name]["v22"]
        for bs in (512, 2048):
            blob = V.compress_blocks_fixed(d, bs)
            assert same(blob, g["fixed_%d" % bs]), (name, bs)
            rt = g["fixed_%d_roundtrip" % bs]
            if rt is True:
                assert V.decompress(blob) == d, (name, bs)
            else:                                   # the reference's own decoder raises (bit-plane variant, len % 8 != 0)
                with pytest.raises(Exception):
                    V.decompress(blob)
        if d:
            blob = V.compress_blocks_cdc(d, 128, 256, 512)
            assert same(blob, g["cdc_256"]), name
            assert [list(x) for x in V.cdc_fast_boundaries_strict(d, 128, 256, 512)] == g["cdc_256_bounds"], name


@pytest.mark.skipif(not MEDIUM, reason="medium goldens not generated")
def test_medium_containers():
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    for name, d in sorted(datasets.medium_cases().items()):
        g = MEDIUM[name]
        blob = KF.compress(d, target_block=2048)
        assert same(blob, g["kf_container_2048"]), name
        assert KF.decompress(blob) == d
        assert same(V.compress_blocks_fixed(d, 2048), g["v22_fixed_2048"]), name
        assert same(V.compress_blocks_cdc(d, 1024, 2048, 4096), g["v22_cdc_2048"]), name


@pytest.mark.parametrize("name", sorted(datasets.FIXTURES))
def test_fixture_containers_kolm(name):
    """BASELINE cfg 1 and north_star: byte-exact containers on all test_binary_files (default target_block 8192)."""
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    g = json.load(open(os.path.join(GOLD, "fixture_%s_kf.json" % name)))
    d = datasets.fixture(name)
    blob = KF.compress(d)
    assert [list(x) for x in _kolm_table(blob)] == g["blocks"], name
    assert same(blob, g["container"]), name
    xz = os.path.join(GOLD, "fixture_%s_kf.bin.xz" % name)
    if os.path.exists(xz):
        assert blob == lzma.open(xz).read()
    assert KF.decompress(blob) == d


@pytest.mark.parametrize("name", sorted(datasets.FIXTURES))
def test_fixture_containers_kolr(name):
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    path = os.path.join(GOLD, "fixture_%s_v22.json" % name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = json.load(open(path))
    d = datasets.fixture(name)
    blob = V.compress_blocks_fixed(d, 2048)
    assert same(blob, g["container"]), name
    try:
        assert V.decompress(blob) == d
    except IndexError:
        pass        # a bit-plane winner on a short last block is undecodable in the reference too


def _kolm_table(blob):
    import struct
    nb = struct.unpack_from("<H", blob, 16)[0]
    p, out = 18, []
    for _ in range(nb):
        m = blob[p]
        ol, pl = struct.unpack_from("<II", blob, p + 1)
        out.append((m, ol, pl))
        p += 9 + pl
    return out


def test_error_behaviour():
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF
    from kolmogorovlike_datacompressor_b200 import kolm_final_researched_v2_2 as V
    with pytest.raises(ValueError):
        KF.decompress(b"XXXX" + bytes(20))
    with pytest.raises(ValueError):
        V.decompress(b"XXXX" + bytes(20))
    d = datasets.small_cases()["text"]
    blob = KF.compress(d, 512)
    with pytest.raises(EOFError):
        KF.decompress(blob[:len(blob) - 5])
    v = V.compress_blocks_fixed(d, 512)
    with pytest.raises(ValueError):
        V.decompress(v + b"\0")                     # strict trailing-byte check (v2-2.py:2547-2549)
    assert KF.compress(b"") == bytes.fromhex("4b4f4c4d" "00200000" "0000000000000000" "0000")
    assert V.compress_blocks_fixed(b"", 2048) == bytes.fromhex("4b4f4c520008000000000000000004000000000000")
    assert V.decompress(V.compress_blocks_fixed(b"", 2048)) == b""



def test_command_lines(tmp_path, capsys):
    """The reference's two command lines (kolm_final.py:963-984, kolm_final_researched_v2-2.py:2624-2692): same flags, same
    default output names and messages, byte-identical containers (cfg 1: checker.bmp -> 1843 bytes)."""
    from kolmogorovlike_datacompressor_b200 import kolm_final as KF, kolm_final_researched_v2_2 as V
    name = sorted(datasets.FIXTURES)[0]
    d = datasets.fixture(name)
    src = tmp_path / "in.bin"
    src.write_bytes(d)
    assert KF.main([str(src)]) == 0
    blob = (tmp_path / "in.bin.kolm").read_bytes()
    g = json.load(open(os.path.join(GOLD, "fixture_%s_kf.json" % name)))
    assert same(blob, g["container"])
    assert f"Compressed {len(d)} bytes to {len(blob)} bytes" in capsys.readouterr().out
    assert KF.main(["-d", str(tmp_path / "in.bin.kolm"), "-o", str(tmp_path / "back")]) == 0
    assert (tmp_path / "back").read_bytes() == d
    assert V.main(["-i", str(src), "-b", "2048"]) == 0
    blob = (tmp_path / "in.bin.kolr").read_bytes()
    g = json.load(open(os.path.join(GOLD, "fixture_%s_v22.json" % name)))
    assert same(blob, g["container"])
    assert "[FIXED(block=2048)] Compressed" in capsys.readouterr().out
    # --only filters the candidate list and method ids index the FILTERED list (v2-2.py:2170-2176), so, as in the reference,
    # only a filter that keeps id 0 = raw in place decodes again with the full decoder table
    assert V.main(["-i", str(src), "--fastcdc", "-b", "4096", "--only", "raw", "-o", str(tmp_path / "c.kolr")]) == 0
    assert "[FastCDC(min=2048, avg=4096, max=8192)]" in capsys.readouterr().out
    assert V.main(["-d", "-i", str(tmp_path / "c.kolr")]) == 0
    assert (tmp_path / "c.out").read_bytes() == d
    assert V.main(["-i", str(src), "--fastcdc", "-b", "4096"]) == 0 and V.G_ONLY_METHOD is None
    assert V.decompress((tmp_path / "in.bin.kolr").read_bytes()) == d


def test_corpus_as_sequence_of_containers():
    """Inputs beyond one container's header limits become a sequence of containers, each the reference's bytes for its span."""
    from kolmogorovlike_datacompressor_b200 import corpus, kolm_final as K, kolm_final_researched_v2_2 as V
    data = (datasets.fixture("pattern")[:150000] + datasets.fixture("sine")[:100000] + datasets.medium_cases()["text_big"] * 3)
    parts = corpus.compress_corpus(data, 2048, "kolr", max_blocks=40)
    spans = corpus.container_spans(len(data), 2048, max_blocks=40)
    assert len(parts) == len(spans) > 3
    for (a, b), c in zip(spans, parts):
        assert c == V.compress_blocks_fixed(data[a:b], 2048)
    assert corpus.decompress_corpus(parts, "kolr") == data
    parts = corpus.compress_corpus(data, 4096, "kolm", max_bytes=100000)
    assert len(parts) >= 3 and corpus.decompress_corpus(parts, "kolm") == data
    assert parts[0] == K.compress(data[:corpus.container_spans(len(data), 2048, max_bytes=100000)[0][1]], 4096)
