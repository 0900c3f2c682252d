"""benchmark_compare — the reference's only published experiment (final/benchmark_compare.py:157-256) on the GPU drop-in.

Same five hand-made data sets, same three coders (the KOLM container of `kolm_final`, the BBWT -> MTF -> run-length
baseline, the naive LZ77 baseline), same columns (ratio, comp_ms, decomp_ms, valid).  The stage functions come from the
drop-in module `kolm_final` of this package, i.e. every transform runs through libkolm_b200 on the GPU.  The reference's
`source_code` sample is the first 4 KiB of its own script; here it is the first 4 KiB of this file.  The PNG bar chart of
the reference needs matplotlib (not a dependency of this package): it is written only when matplotlib imports.

    python -m kolmogorovlike_datacompressor_b200.benchmark_compare
"""
from __future__ import annotations

import random
import time
from typing import Dict, List

from . import kolm_final
from .kolm_final import (bbwt_forward, bbwt_inverse, compress as kolm_compress, decode_model_lz77, decompress as kolm_decompress,
                         encode_model_lz77, mtf_decode, mtf_encode, uleb128_decode_stream, uleb128_encode)


def baseline_bbwt_mtf_rle_encode(block: bytes) -> bytes:
    """BBWT -> MTF -> (tag, ULEB) run-length stream (benchmark_compare.py:67-95)."""
    seq = mtf_encode(bbwt_forward(block))
    out = bytearray()
    run = 0
    for v in seq:
        if v == 0:
            run += 1
            continue
        if run > 0:
            out.append(0)
            out += uleb128_encode(run)
            run = 0
        out.append(1)
        out += uleb128_encode(v - 1)
    if run > 0:
        out.append(0)
        out += uleb128_encode(run)
    return bytes(out)


def baseline_bbwt_mtf_rle_decode(payload: bytes, orig_len: int) -> bytes:
    """Inverse of the above (benchmark_compare.py:98-130); same ValueError cases."""
    seq: List[int] = []
    i, n = 0, len(payload)
    while i < n:
        tag = payload[i]
        i += 1
        value, i = uleb128_decode_stream(payload, i)
        if tag == 0:
            seq.extend([0] * value)
        elif tag == 1:
            seq.append(value + 1)
        else:
            raise ValueError(f"unknown tag {tag} in baseline decode")
    if len(seq) != orig_len:
        raise ValueError(f"baseline decode produced {len(seq)} symbols, expected {orig_len}")
    return bbwt_inverse(mtf_decode(bytes(seq)))


def baseline_lz77_encode(block: bytes) -> bytes:
    payload, _meta = encode_model_lz77(block)
    return payload


def baseline_lz77_decode(payload: bytes, orig_len: int) -> bytes:
    return decode_model_lz77(payload, orig_len)


def data_sets() -> Dict[str, bytes]:
    """benchmark_compare.py:166-173."""
    return {
        "repetitive_text": b"A" * 2000 + b"B" * 1000 + (b"CD" * 500),
        "english_like": (b"In compression we favor short programs and transparent circuits. " * 20),
        "source_code": open(__file__, "rb").read()[:4096],
        "byte_counter": bytes([i % 256 for i in range(4096)]),
        "random_bytes": bytes(random.Random(42).getrandbits(8) for _ in range(4096)),
    }


def _timed(fn, *a):
    t0 = time.perf_counter()
    r = fn(*a)
    return r, (time.perf_counter() - t0) * 1000.0


def run_benchmarks(sets: Dict[str, bytes] | None = None) -> List[Dict[str, object]]:
    """One row per (dataset, algorithm), the reference's columns.  Returns a list of dicts (a pandas DataFrame of it is
    `pandas.DataFrame(rows)`, which is what the reference builds)."""
    rows: List[Dict[str, object]] = []
    kolm_compress(b"warm up the context")                              # context creation is not part of any row
    for name, data in (sets or data_sets()).items():
        n = len(data)
        for algo, enc, dec in (("kolm_final", lambda d: kolm_compress(d), lambda p, _n: kolm_decompress(p)),
                               ("baseline_bbwt_mtf_rle", baseline_bbwt_mtf_rle_encode, baseline_bbwt_mtf_rle_decode),
                               ("baseline_lz77", baseline_lz77_encode, baseline_lz77_decode)):
            payload, cms = _timed(enc, data)
            try:
                back, dms = _timed(dec, payload, n)
                ok = back == data
            except Exception:
                ok, dms = False, float("nan")
            rows.append({"dataset": name, "algorithm": algo, "ratio": len(payload) / n, "comp_ms": cms, "decomp_ms": dms, "valid": ok,
                         "bytes": len(payload)})
    return rows


def main() -> int:
    rows = run_benchmarks()
    print("%-16s %-22s %8s %10s %10s %5s" % ("dataset", "algorithm", "ratio", "comp_ms", "decomp_ms", "valid"))
    for r in rows:
        print("%-16s %-22s %8.4f %10.2f %10.2f %5s" % (r["dataset"], r["algorithm"], r["ratio"], r["comp_ms"], r["decomp_ms"], r["valid"]))
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        import pandas as pd
        df = pd.DataFrame(rows)
        fig, axs = plt.subplots(3, 1, figsize=(8, 10))
        for ax, metric in zip(axs, ["ratio", "comp_ms", "decomp_ms"]):
            df.pivot(index="dataset", columns="algorithm", values=metric).plot.bar(ax=ax)
            ax.set_ylabel(metric)
        plt.tight_layout()
        plt.savefig("kolm_comparison_plot.png", dpi=150)
    except ImportError:
        pass
    return 0 if all(r["valid"] for r in rows) else 1


if __name__ == "__main__":
    raise SystemExit(main())
