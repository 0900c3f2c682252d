"""Deterministic small parity inputs shared by make_golden.py and the tests.

Shapes follow the reference's own ad-hoc datasets (benchmark_compare.py:166-173 and the
self-test list in kolm_final_researched_v2-2.cpp:4726-4789) at sizes the pure-Python
reference finishes in seconds, plus slices of the four binary fixtures and edge cases.
"""
from __future__ import annotations

import lzma
import os
import random

_FIX = os.path.join(os.path.dirname(__file__), "golden", "fixtures")

FIXTURES = {
    "checker": "example_checker_640x480.bmp",
    "gradient": "example_gradient_1024x768.bmp",
    "pattern": "example_pattern_blocks.bin",
    "sine": "example_sine_44k_3s.wav",
}
_fix_cache = {}


def fixture(name: str) -> bytes:
    """One of the reference's test_binary_files (committed xz-compressed, data only)."""
    if name not in _fix_cache:
        with lzma.open(os.path.join(_FIX, FIXTURES[name] + ".xz"), "rb") as f:
            _fix_cache[name] = f.read()
    return _fix_cache[name]


_PARA = (b"In a hole in the ground there lived a hobbit. Not a nasty, dirty, wet hole, "
         b"filled with the ends of worms and an oozy smell, nor yet a dry, bare, sandy hole "
         b"with nothing in it to sit down on or to eat: it was a hobbit-hole, and that means comfort.\n")

_CODE = (b"def step(state, taps=0x96):\n    fb = 0\n    for bit in range(8):\n"
         b"        if (taps >> bit) & 1:\n            fb ^= (state >> bit) & 1\n"
         b"    return ((state << 1) & 0xFF) | fb\n\n"
         b"for (int i = 0; i < n; ++i) { out[i] = in[i] ^ prev; prev = in[i]; }\n")


def small_cases() -> dict:
    """name -> bytes; every case is <= 4 KiB so the Python reference handles it quickly."""
    rnd = random.Random(42)
    c = {}
    c["empty"] = b""
    c["one"] = b"A"
    c["aa"] = b"aa"
    c["ab"] = b"ab"
    c["ba"] = b"ba"
    c["banana"] = b"banana"
    c["zeros17"] = bytes(17)
    c["ff33"] = b"\xff" * 33
    c["abab"] = b"ab" * 301
    c["abcabc"] = b"abc" * 211
    c["desc"] = bytes(range(255, -1, -1)) * 3
    c["repetitive_text"] = b"A" * 2000 + b"B" * 1000 + (b"CD" * 500)
    c["english_like"] = b"In compression we favor short programs and transparent circuits. " * 20
    c["code_like"] = (_CODE * 12)[:3001]
    c["byte_counter"] = bytes(i % 256 for i in range(4096))
    c["random_bytes"] = bytes(rnd.getrandbits(8) for _ in range(4096))
    c["random_small_alpha"] = bytes(rnd.choice(b"abc") for _ in range(2500))
    c["random_binary"] = bytes(rnd.choice(b"\x00\x01") for _ in range(1999))
    c["text"] = (_PARA * 10)[:3350]
    c["zero_2k"] = bytes(2048)
    c["ramp_odd"] = bytes(i & 0xFF for i in range(1021))
    c["utf8_mixed"] = ("Grüße, 世界! Καλημέρα κόσμε — 数据压缩 test \U0001f600 " * 30).encode("utf-8")[:2043]
    c["runs18"] = b"".join(bytes([v & 0xFF]) * 18 for v in range(120))
    fib = [1, 1]
    while len(fib) < 1500:
        fib.append((fib[-1] + fib[-2]) & 0xFF)
    c["fib"] = bytes(fib)
    c["checker_0"] = fixture("checker")[:2048]
    c["checker_100000"] = fixture("checker")[100000:102048]
    c["gradient_200000"] = fixture("gradient")[200000:202048]
    c["pattern_393316"] = fixture("pattern")[393316:395364]
    c["sine_1000"] = fixture("sine")[1000:3048]
    c["sine_head_odd"] = fixture("sine")[:1234]
    return c


def medium_cases() -> dict:
    """Larger inputs (16-64 KiB) for container-level goldens (minutes in the Python reference)."""
    c = {}
    c["checker_16k"] = fixture("checker")[:16384]
    c["pattern_mix_24k"] = (fixture("pattern")[0:4096] + fixture("pattern")[65536 * 2:65536 * 2 + 4096]
                            + fixture("pattern")[65536 * 3:65536 * 3 + 4096] + fixture("pattern")[65536 * 4:65536 * 4 + 4096]
                            + fixture("pattern")[65536 * 5:65536 * 5 + 4096] + fixture("pattern")[65536 * 9:65536 * 9 + 4096])
    c["sine_20k"] = fixture("sine")[:20001]
    c["text_big"] = (_PARA * 80)[:20000]
    return c
