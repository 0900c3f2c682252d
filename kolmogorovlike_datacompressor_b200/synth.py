"""Seeded synthetic corpora for the BASELINE.json configurations.

S1  low-entropy text-like corpus (cfg 2): Zipfian words over a 4096-word vocabulary.
S2  mixed gradient / sine / pattern / checker segments (cfg 3).
S3  repeating 8-segment mix of S1, S2 and random bytes (cfg 4, 5).
All generators are pure numpy and deterministic; `n` is the exact output size in bytes.
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
MIB = 1 << 20


def _vocab(rng):
    lens = 2 + (np.arange(4096) % 9)
    p = 1.0 / (np.arange(len(_LETTERS)) + 1.0)
    p /= p.sum()
    table = np.zeros((4096, 10), dtype=np.uint8)
    for i in range(4096):
        table[i, :lens[i]] = _LETTERS[rng.choice(len(_LETTERS), size=lens[i], p=p)]
    return table, lens


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(16, len(os.sched_getaffinity(0)))), thread_name_prefix="kolm-synth")
    return _POOL


def s1_text(n: int, seed: int = 0xC0FFEE) -> np.ndarray:
    """Words sampled with p(i) ~ 1/(i+1), joined by ' ', '\n' after every 12th word, truncated to n bytes.
    The random stream is drawn sequentially (the corpus is a function of (n, seed) only); the table lookups and the expansion of
    word ids into bytes run on a thread pool, slice by slice (numpy releases the GIL in all of them)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    table, lens = _vocab(rng)
    ext = np.zeros((4096, 11), dtype=np.uint8)            # letters, then one separator slot
    ext[:, :10] = table
    flat = ext.reshape(-1)
    pw = 1.0 / (np.arange(4096) + 1.0)
    pw /= pw.sum()
    cdf = np.cumsum(pw)
    wl_tab = (lens + 1).astype(np.int64)
    out = np.empty(n, dtype=np.uint8)
    filled = 0
    widx = 0
    chunk = 1 << 21
    pool = _pool()
    nsl = 16
    step = chunk // nsl
    while filled < n:
        u = rng.random(chunk)
        ids = np.empty(chunk, dtype=np.int64)

        def look(k):
            ids[k * step:(k + 1) * step] = np.searchsorted(cdf, u[k * step:(k + 1) * step], side="right").clip(0, 4095)
        list(pool.map(look, range(nsl)))
        wl = wl_tab[ids]
        ends = np.cumsum(wl)
        tot = int(ends[-1])
        take = min(tot, n - filled)
        base_w, base_o = widx, filled

        def expand(k):
            a, b = k * step, (k + 1) * step
            o0 = int(ends[a] - wl[a])                       # byte offset of word a inside this chunk
            if o0 >= take:
                return
            w = wl[a:b]
            e = ends[a:b] - o0
            m = int(e[-1])
            word_of = np.repeat(np.arange(b - a, dtype=np.int64), w)
            within = np.arange(m, dtype=np.int64) - np.repeat(e - w, w)
            buf = flat[ids[a:b][word_of] * 11 + within]
            sep = np.full(b - a, 0x20, dtype=np.uint8)
            sep[(np.arange(base_w + a, base_w + b) % 12) == 11] = 0x0A
            buf[e - 1] = sep
            m = min(m, take - o0)
            out[base_o + o0:base_o + o0 + m] = buf[:m]
        list(pool.map(expand, range(nsl)))
        widx += chunk
        filled += take
    return out


def _seg_gradient(n):
    i = np.arange(n, dtype=np.int64)
    pix, ch = i // 3, i % 3
    x, y = pix % 1024, pix // 1024
    r = (255 - (x >> 2)) & 255
    g = (255 - y // 3) & 255
    b = ((x + y) >> 3) & 255
    return np.where(ch == 0, r, np.where(ch == 1, g, b)).astype(np.uint8)


def _seg_sine(n):
    t = np.arange((n + 1) // 2, dtype=np.float64)
    s = np.round(30000.0 * np.sin(2.0 * np.pi * 440.0 * t / 44100.0)).astype("<i2")
    return s.view(np.uint8)[:n].copy()


def _seg_pattern(n):
    out = np.empty(n, dtype=np.uint8)
    sub = 65536
    for k in range(0, n, sub):
        m = min(sub, n - k)
        kind = (k // sub) % 6
        if kind == 0:
            out[k:k + m] = 0
        elif kind == 1:
            out[k:k + m] = 0xFF
        elif kind == 2:
            out[k:k + m] = np.arange(m) & 0xFF
        elif kind == 3:
            f = np.empty(m, dtype=np.uint8)
            a, b = 1, 1
            for i in range(m):
                f[i] = a
                a, b = b, (a + b) & 0xFF
            out[k:k + m] = f
        elif kind == 4:
            x = np.empty(m, dtype=np.uint64)
            v = 20251018
            for i in range(m):
                v = (1664525 * v + 1013904223) & 0xFFFFFFFF
                x[i] = v >> 24
            out[k:k + m] = x.astype(np.uint8)
        else:
            out[k:k + m] = ((np.arange(m) // 18) & 0xFF).astype(np.uint8)
    return out


def _seg_checker(n):
    i = np.arange(n, dtype=np.int64)
    pix = i // 3
    x, y = pix % 640, pix // 640
    return np.where(((x // 32) + (y // 32)) % 2 == 0, 40, 200).astype(np.uint8)


_S2_CACHE = {}


def s2_segment(kind: str, n: int = MIB) -> np.ndarray:
    key = (kind, n)
    if key not in _S2_CACHE:
        _S2_CACHE[key] = {"g": _seg_gradient, "s": _seg_sine, "p": _seg_pattern, "c": _seg_checker}[kind](n)
    return _S2_CACHE[key]


def s2_mixed(n: int) -> np.ndarray:
    """1 MiB segments cycling gradient, sine, pattern, checker."""
    out = np.empty(n, dtype=np.uint8)
    for k in range(0, n, MIB):
        m = min(MIB, n - k)
        out[k:k + m] = s2_segment("gspc"[(k // MIB) % 4], MIB)[:m]
    return out


def s3_mix(n: int, seed: int = 7, text_seed: int = 0xC0FFEE) -> np.ndarray:
    """1 MiB segments in the repeating pattern [S1, S1, S2g, S2s, S2p, S2c, S1, random]; `seed` drives the random segments,
    `text_seed` the S1 text (the containers of a multi-container corpus use different seeds)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nseg = (n + MIB - 1) // MIB
    n_s1 = sum(1 for k in range(nseg) if k % 8 in (0, 1, 6))
    text = s1_text(max(1, n_s1) * MIB, seed=text_seed)
    out = np.empty(n, dtype=np.uint8)
    ti = 0
    for k in range(nseg):
        a = k * MIB
        m = min(MIB, n - a)
        r = k % 8
        if r in (0, 1, 6):
            out[a:a + m] = text[ti * MIB:ti * MIB + m]
            ti += 1
        elif r == 7:
            out[a:a + m] = rng.integers(0, 256, size=m, dtype=np.uint8)
        else:
            out[a:a + m] = s2_segment("gspc"[r - 2], MIB)[:m]
    return out
