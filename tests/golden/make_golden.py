#!/usr/bin/env python3
"""Generate the golden vectors by running the UNMODIFIED Python reference.

Runs only in the build container (needs /root/reference).  The outputs are committed:

    tests/golden/small.json      stage-level + per-model + tiny-container vectors (<= 4 KiB inputs)
    tests/golden/medium.json     container-level vectors for 16-24 KiB inputs
    tests/golden/fixture_<name>_<kf|v22>.json (+ .bin.xz for small containers)
                                 full test_binary_files containers (sha256, per-block table)

    python tests/golden/make_golden.py small
    python tests/golden/make_golden.py medium
    python tests/golden/make_golden.py fixtures        # ~15-25 min on 8 cores
"""
from __future__ import annotations

import hashlib
import json
import lzma
import os
import struct
import sys
import time
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader  # noqa: E402
import datasets  # noqa: E402


def sha(b: bytes) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def rec(b: bytes) -> dict:
    """sha256 + length, and the bytes themselves when short."""
    b = bytes(b)
    d = {"len": len(b), "sha256": sha(b)}
    if len(b) <= 96:
        d["hex"] = b.hex()
    return d


def parse_kolm(blob: bytes):
    """(method, orig_len, payload_len) per block of a KOLM container (kolm_final.py:883-901)."""
    p = 18
    nb = struct.unpack_from("<H", blob, 16)[0]
    out = []
    for _ in range(nb):
        m = blob[p]
        ol, pl = struct.unpack_from("<II", blob, p + 1)
        out.append([m, ol, pl])
        p += 9 + pl
    return out


def v22_block_table(V, data: bytes, boundaries):
    """Re-run the V22 selection loop to expose per-block (method, orig_len, payload_len, sizes)."""
    cands = V._select_encoders()
    table = []
    for (a, b) in boundaries:
        block = data[a:b]
        sizes = []
        for enc, _name in cands:
            try:
                payload, _ = enc(block)
                sizes.append(len(payload))
            except Exception:
                sizes.append(None)
        live = [(s, i) for i, s in enumerate(sizes) if s is not None]
        best = min(live)  # strict '<' in reference == lowest id on ties
        table.append([best[1], b - a, best[0], sizes])
    return table


def gen_small():
    K = ref_loader.load_kf()
    V = ref_loader.load_v22()
    out = {}
    for name, data in datasets.small_cases().items():
        t0 = time.time()
        e = {"n": len(data), "input_sha256": sha(data)}
        # ---- stage level (identical in KF and V22; asserted) ----
        facs = K.duval_lyndon(data)
        assert facs == V.duval_lyndon(data)
        e["lyndon_starts"] = [a for a, _ in facs]
        L = K.bbwt_forward(data)
        assert L == V.bbwt_forward(data)
        assert K.bbwt_inverse(L) == data
        e["bbwt"] = rec(L)
        mtf = K.mtf_encode(L)
        assert K.mtf_decode(mtf) == L
        e["mtf"] = rec(bytes(mtf))
        # ---- KF models ----
        kf = {}
        for mid, enc in K._ENCODERS.items():
            payload, meta = enc(data)
            kf[str(mid)] = rec(payload)
            if mid == 2:
                kf["m2_meta"] = {k: int(v) for k, v in meta.items()}
            assert K._DECODERS[mid](payload, len(data)) == data
        mid, payload, plen = K._encode_block(data)
        kf["selected"] = mid
        for tb in (512, 8192):
            blob = K.compress(data, target_block=tb)
            assert K.decompress(blob) == data
            kf["container_%d" % tb] = rec(blob)
            kf["blocks_%d" % tb] = parse_kolm(blob)
        e["kf"] = kf
        # ---- V22 candidates ----
        v = {}
        sizes = []
        for i, (enc, nm) in enumerate(V._select_encoders()):
            try:
                payload, _ = enc(data)
                v[nm] = rec(payload)
                sizes.append(len(payload))
            except Exception as ex:  # v2_new: NameError in the shipped reference
                v[nm] = {"error": type(ex).__name__}
                sizes.append(None)
        v["sizes"] = sizes
        for bs in (512, 2048):
            blob = V.compress_blocks_fixed(data, bs)
            v["fixed_%d" % bs] = rec(blob)
            try:
                ok = V.decompress(blob) == data
            except Exception as ex:
                ok = "raises:" + type(ex).__name__
            v["fixed_%d_roundtrip" % bs] = ok
        if len(data) >= 1:
            blob = V.compress_blocks_cdc(data, 128, 256, 512)
            v["cdc_256"] = rec(blob)
            v["cdc_256_bounds"] = V.cdc_fast_boundaries_strict(data, 128, 256, 512)
        e["v22"] = v
        e["kf_cdc_512_bounds"] = K.cdc_fast_boundaries(data, 256, 512, 1024)
        out[name] = e
        print(f"small {name:22s} n={len(data):5d} {time.time()-t0:6.1f}s", flush=True)
    with open(os.path.join(HERE, "small.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


def _medium_one(args):
    name, = args
    K = ref_loader.load_kf()
    V = ref_loader.load_v22()
    data = datasets.medium_cases()[name]
    e = {"n": len(data), "input_sha256": sha(data)}
    blob = K.compress(data, target_block=2048)
    assert K.decompress(blob) == data
    e["kf_container_2048"] = rec(blob)
    e["kf_blocks_2048"] = parse_kolm(blob)
    blob = V.compress_blocks_fixed(data, 2048)
    e["v22_fixed_2048"] = rec(blob)
    e["v22_fixed_2048_table"] = v22_block_table(V, data, V.fixed_boundaries(data, 2048))
    blob = V.compress_blocks_cdc(data, 1024, 2048, 4096)
    e["v22_cdc_2048"] = rec(blob)
    e["v22_cdc_2048_bounds"] = V.cdc_fast_boundaries_strict(data, 1024, 2048, 4096)
    return name, e


def gen_medium():
    names = list(datasets.medium_cases().keys())
    with Pool(min(8, len(names))) as p:
        res = dict(p.map(_medium_one, [(n,) for n in names]))
    with open(os.path.join(HERE, "medium.json"), "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)


def _fixture_one(args):
    name, which = args
    t0 = time.time()
    data = datasets.fixture(name)
    e = {"n": len(data), "input_sha256": sha(data), "fixture": datasets.FIXTURES[name]}
    if which == "kf":
        K = ref_loader.load_kf()
        blob = K.compress(data)  # target_block 8192 (reference default)
        e["call"] = "kolm_final.compress(data)"
        e["blocks"] = parse_kolm(blob)
    else:
        V = ref_loader.load_v22()
        blob = V.compress_blocks_fixed(data, 2048)
        e["call"] = "compress_blocks_fixed(data, 2048)"
    e["container"] = rec(blob)
    e["seconds"] = round(time.time() - t0, 1)
    base = os.path.join(HERE, f"fixture_{name}_{which}")
    if len(blob) <= 300_000:
        with lzma.open(base + ".bin.xz", "wb", preset=9) as f:
            f.write(blob)
    with open(base + ".json", "w") as f:
        json.dump(e, f, indent=0, sort_keys=True)
    print(f"fixture {name} {which}: {len(blob)} bytes in {e['seconds']} s", flush=True)
    return name, which


def gen_fixtures():
    jobs = [(n, w) for w in ("kf", "v22") for n in datasets.FIXTURES]
    with Pool(8) as p:
        p.map(_fixture_one, jobs)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    assert ref_loader.available(), "reference not mounted"
    {"small": gen_small, "medium": gen_medium, "fixtures": gen_fixtures}[what]()
