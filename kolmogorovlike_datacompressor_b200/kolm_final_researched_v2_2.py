"""Drop-in for the reference's final_researched/kolm_final_researched_v2-2.py ('KOLR' container + TOC) with the
per-block hot path on the GPU.

Same entry points, signatures, container bytes and error behaviour:
    compress_blocks_fixed(data, block_size=8192)                       v2-2.py:2332-2445
    compress_blocks_cdc(data, min_size=4096, avg_size=8192, max_size=16384)   v2-2.py:2213-2326
    decompress(container)                                              v2-2.py:2451-2550
    _select_encoders() / _select_decoders(), G_NO_LZ77 / G_ONLY_METHOD / G_PROGRESS   v2-2.py:92-96, 2152-2207
    fixed_boundaries, cdc_fast_boundaries_strict                       v2-2.py:210-320

Candidate ids = index in the (possibly filtered) list {raw, xor(delta), bbwt, bbwt_bp, bbwt_nib, bbwt_br, bbwt_gray, lz77,
lfsr_pred, repair, v2_new}; the smallest payload wins, first wins ties.  v2_new raises NameError in the shipped reference
and is therefore never selected; the same holds here.  The TOC (RLE + canonical Huffman of method ids, Rice
run lengths, ZigZag/Rice block lengths in CDC mode, Elias-Fano payload ends) is built on the host with the same heapq
tie behaviour as the reference.
"""
from __future__ import annotations

import heapq
import struct
from collections import Counter
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from .engine import KOLR_NAMES, Engine, cdc_boundaries

G_NO_LZ77: bool = False
G_ONLY_METHOD: Optional[str] = None
G_PROGRESS: bool = False
# Not a reference global.  The shipped reference can never select method 10: its encode_new_pipeline raises NameError (missing
# imports on the default parallel=True path, v2-2.py:1037-1042) and the selection loops swallow it.  False reproduces that —
# containers are byte-identical to the reference's.  True evaluates the candidate the way the reference's own function does
# with parallel=False (v2-2.py:1030-1032), i.e. what the file produces once its imports are repaired.
G_ENABLE_V2_NEW: bool = False

MODE_FIXED = 0
MODE_CDC = 1


GPU_CDC_MIN_BYTES = 8 << 20   # below this the host scan (~1 GB/s) beats a copy + launch + list read-back


def _engine() -> Engine:
    e = Engine.shared()
    e.enable_v2_new = bool(G_ENABLE_V2_NEW)
    return e


def _print_progress(label: str, i: int, n: int, final: bool = False) -> None:
    if not G_PROGRESS:
        return
    if not final:
        print(f"[{label}] block {i}/{n} ...", end="\r", flush=True)
    else:
        print(f"[{label}] block {n}/{n} done.", flush=True)


# ---- ULEB128 / chunking ----------------------------------------------------------------------------
def uleb128_encode(n: int) -> bytes:
    if n < 0:
        raise ValueError("ULEB128 only supports unsigned integers")
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def uleb128_decode_stream(data: bytes, pos: int = 0) -> Tuple[int, int]:
    shift = result = 0
    while True:
        if pos >= len(data):
            raise ValueError("Truncated ULEB128")
        b = data[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def fixed_boundaries(data: bytes, block_size: int = 8192) -> List[Tuple[int, int]]:
    n = len(data)
    if n == 0:
        return []
    if block_size <= 0:
        raise ValueError("block_size must be positive")
    return [(i, min(n, i + block_size)) for i in range(0, n, block_size)]


class _FixedBounds:
    """fixed_boundaries(data of n bytes, block_size) as a read-only sequence without the list: what _assemble needs of it in FIXED
    mode is the length and the last block (dist.compress_kolr_fixed_corpus assembles containers of tens of thousands of blocks
    that were encoded elsewhere)."""

    def __init__(self, n: int, block_size: int):
        if block_size <= 0:
            raise ValueError("block_size must be positive")
        self.n, self.bs = int(n), int(block_size)
        self.count = (self.n + self.bs - 1) // self.bs

    def __len__(self) -> int:
        return self.count

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self.count))]
        if i < 0:
            i += self.count
        if not 0 <= i < self.count:
            raise IndexError("block index out of range")
        return (i * self.bs, min(self.n, (i + 1) * self.bs))

    def __iter__(self):
        return ((a, min(self.n, a + self.bs)) for a in range(0, self.n, self.bs))


def cdc_fast_boundaries_strict(data: bytes, min_size: int = 4096, avg_size: int = 8192, max_size: int = 16384,
                               merge_orphan_tail: bool = True) -> List[Tuple[int, int]]:
    if len(data) == 0:
        return []
    if not (min_size > 0 and min_size <= avg_size <= max_size):
        raise ValueError("Require 0 < min_size <= avg_size <= max_size")
    if avg_size < 64:
        raise ValueError("avg_size too small; use >= 64")
    if not merge_orphan_tail:
        raise NotImplementedError("merge_orphan_tail=False is not used by the reference's compressors")
    data = bytes(data)
    if len(data) >= GPU_CDC_MIN_BYTES:                       # the per-byte scan runs on the GPU, the chain over chunks on the host
        return _engine().cdc_boundaries("v22", data, min_size, avg_size, max_size)
    return cdc_boundaries("v22", data, min_size, avg_size, max_size)


# ---- candidate registry ------------------------------------------------------------------------------
def _candidate_names() -> List[str]:
    names = list(KOLR_NAMES)
    if G_NO_LZ77:
        names = [n for n in names if n != "lz77"]
    if G_ONLY_METHOD is not None:
        only = G_ONLY_METHOD.lower()
        names = [n for n in names if n.lower() == only]
        if not names:
            raise ValueError(f"--only={G_ONLY_METHOD} not found in candidates")
    return names


def _select_encoders() -> List[Tuple[Callable[[bytes], Tuple[bytes, Dict[str, Any]]], str]]:
    def mk(name):
        return lambda block: (_engine().kolr_model_payload(bytes(block), name), {})
    return [(mk(n), n) for n in _candidate_names()]


def _select_decoders() -> List[Callable[[bytes, int, Dict[str, Any]], bytes]]:
    def mk(name):
        return lambda payload, byte_length, meta=None: _engine().decode_blocks([(name, bytes(payload), byte_length)])[0]
    return [mk(n) for n in KOLR_NAMES]


# ---- TOC coders (host; O(nblocks)) ---------------------------------------------------------------------
class _Bits:
    """MSB-first bit accumulator (== _BitWriter.getvalue_bits).  Whole bytes leave the integer accumulator as soon as a few
    thousand bits are pending, so a TOC of n blocks costs O(n) (one growing big integer made it quadratic: 0.8 s of host time per
    61 440 blocks, more than the GPU spent on encoding them)."""
    _FLUSH = 4096

    def __init__(self):
        self.acc = 0                               # the pending bits (fewer than _FLUSH + the last value's width)
        self.pend = 0
        self.n = 0                                 # all bits so far
        self.done: List[bytes] = []

    def _spill(self):
        r = self.pend % 8                          # whole bytes out, fewer than 8 bits stay
        if self.pend > r:
            self.done.append((self.acc >> r).to_bytes((self.pend - r) // 8, "big"))
            self.acc &= (1 << r) - 1
            self.pend = r

    def put(self, val: int, k: int):
        if k:
            self.acc = (self.acc << k) | (val & ((1 << k) - 1))
            self.pend += k
            self.n += k
            if self.pend >= self._FLUSH:
                self._spill()

    def unary(self, q: int):
        self.acc = (self.acc << (q + 1)) | (((1 << q) - 1) << 1)
        self.pend += q + 1
        self.n += q + 1
        if self.pend >= self._FLUSH:
            self._spill()

    def rice(self, seq, k: int):
        for v in seq:
            self.unary(v >> k)
            self.put(v, k)

    def put_array(self, bits: "np.ndarray"):
        """append a 0/1 uint8 array (MSB first = array order)"""
        nb = int(bits.size)
        if not nb:
            return
        self._spill()
        r = self.pend                              # < 8 bits wait in acc: they lead the array
        if r:
            lead = np.array([(self.acc >> (r - 1 - i)) & 1 for i in range(r)], dtype=np.uint8)
            bits = np.concatenate((lead, bits.astype(np.uint8, copy=False)))
        whole = (bits.size // 8) * 8
        if whole:
            self.done.append(np.packbits(bits[:whole]).tobytes())
        tail = bits[whole:]
        self.acc = 0
        for b in tail.tolist():
            self.acc = (self.acc << 1) | int(b)
        self.pend = int(tail.size)
        self.n += nb

    def put_fixed(self, vals: "np.ndarray", k: int):
        """put(v, k) for every v of an integer array (k <= 63)"""
        if k <= 0 or not vals.size:
            return
        v = vals.astype(np.uint64, copy=False)
        sh = np.arange(k - 1, -1, -1, dtype=np.uint64)
        self.put_array(((v[:, None] >> sh[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1))

    def value(self) -> Tuple[bytes, int]:
        tail = b""
        if self.pend:
            nbytes = (self.pend + 7) // 8
            tail = (self.acc << (nbytes * 8 - self.pend)).to_bytes(nbytes, "big")
        return b"".join(self.done) + tail, self.n


class _BitsIn:
    """MSB-first bit reader over the TOC bit string (v2-2.py:1220-1262 read side).  Constant time per bit (the first version
    shifted one big integer per bit, which made the TOC walk quadratic: 24 MB/s decompress at the default 2 KiB blocks)."""

    def __init__(self, buf: bytes):
        self.buf = bytes(buf)
        self.total = len(self.buf) * 8
        self.pos = 0

    def bit(self) -> int:
        if self.pos >= self.total:
            raise ValueError("BitReader: out of data")
        b = (self.buf[self.pos >> 3] >> (7 - (self.pos & 7))) & 1
        self.pos += 1
        return b

    def bits(self, k: int) -> int:
        if k <= 0:
            return 0
        end = self.pos + k
        if end > self.total:                       # the reference consumes what is left, then raises
            self.pos = self.total
            raise ValueError("BitReader: out of data")
        v = int.from_bytes(self.buf[self.pos >> 3:(end + 7) >> 3], "big")
        v = (v >> ((-end) & 7)) & ((1 << k) - 1)
        self.pos = end
        return v

    def rice(self, k: int) -> int:
        q = 0
        while self.bit() == 1:
            q += 1
        return (q << k) | self.bits(k)

    def unpacked(self):
        """the whole bit string as a numpy 0/1 array (for the vectorised Elias-Fano walk)"""
        return np.unpackbits(np.frombuffer(self.buf, dtype=np.uint8))


def _zz_enc(x: int) -> int:
    return (x << 1) if x >= 0 else ((-x) << 1) - 1


def _zz_dec(n: int) -> int:
    return (n >> 1) if (n & 1) == 0 else -((n + 1) >> 1)


class _HuffNode:
    __slots__ = ("w", "sym", "left", "right")

    def __init__(self, w, sym=None, left=None, right=None):
        self.w, self.sym, self.left, self.right = w, sym, left, right

    def __lt__(self, other):                       # same ordering as the reference node (v2-2.py:1271-1276)
        if self.w != other.w:
            return self.w < other.w
        return (self.sym if self.sym is not None else -1) < (other.sym if other.sym is not None else -1)


def _huff_lengths(freq: Dict[int, int]) -> Dict[int, int]:
    heap = [_HuffNode(max(1, f), sym=s) for s, f in freq.items()]
    if not heap:
        return {}
    if len(heap) == 1:
        return {heap[0].sym: 1}
    heapq.heapify(heap)
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        heapq.heappush(heap, _HuffNode(a.w + b.w, left=a, right=b))
    lengths: Dict[int, int] = {}
    stack = [(heap[0], 0)]
    while stack:
        nd, d = stack.pop()
        if nd.sym is not None:
            lengths[nd.sym] = max(1, d)
        else:
            stack.append((nd.left, d + 1))
            stack.append((nd.right, d + 1))
    return lengths


def _huff_canonical(lengths: Dict[int, int]):
    enc, dec = {}, {}
    code = prev = maxlen = 0
    for sym, L in sorted(lengths.items(), key=lambda kv: (kv[1], kv[0])):
        if L != prev:
            code <<= (L - prev)
            prev = L
        enc[sym] = (code, L)
        dec[(L, code)] = sym
        maxlen = max(maxlen, L)
        code += 1
    return enc, dec, maxlen


def _best_rice_k(seq) -> int:
    best_k, best_bits = 0, 1 << 60
    for k in range(8):
        bits = sum((v >> k) + 1 + k for v in seq)
        if bits < best_bits:
            best_bits, best_k = bits, k
    return best_k


def _ef_choose_l(U: int, n: int) -> int:
    if n <= 0 or U <= 1:
        return 0
    avg = U // n
    if avg <= 1:
        return 0
    return max(0, avg.bit_length() - 1)


def _pack_mode_and_size(mode: int, size: int) -> int:
    if mode not in (MODE_FIXED, MODE_CDC):
        raise ValueError("invalid mode")
    if size < 0 or size > 0x7FFFFFFF:
        raise ValueError("size out of range (must fit in 31 bits)")
    return ((mode & 1) << 31) | (size & 0x7FFFFFFF)


def _assemble(data: bytes, boundaries, mode: int, size_field: int, encoded=None) -> bytes:
    """Container for `boundaries`.  encoded = (method ids, payload lengths, payload area) when the blocks were already encoded
    elsewhere (dist.compress_kolr_*: block ranges sharded over several GPUs); otherwise they are encoded here."""
    out = bytearray(b"KOLR")
    out += struct.pack("<I", _pack_mode_and_size(mode, size_field))
    # dist.*_corpus pass the input LENGTH when the blocks were encoded elsewhere (the assembling rank holds no input bytes)
    out += struct.pack("<I", data if isinstance(data, int) else len(data))   # struct.error beyond 4 GiB-1 / 65535 blocks, like the reference
    out += struct.pack("<H", len(boundaries))
    names = _candidate_names()
    nblocks = len(boundaries)
    label = "FIXED" if mode == MODE_FIXED else "Fast CDC"
    _print_progress(label, 0, nblocks)
    if nblocks and encoded is not None:
        mids_np, lens_np, area = np.asarray(encoded[0], dtype=np.int64), np.asarray(encoded[1], dtype=np.int64), encoded[2]
    elif nblocks:
        mids_np, lens_np, area = _engine().encode_kolr_area(data, boundaries, names)
        mids_np, lens_np = np.asarray(mids_np, dtype=np.int64), np.asarray(lens_np, dtype=np.int64)
    else:
        mids_np, lens_np, area = np.zeros(0, np.int64), np.zeros(0, np.int64), b""
    _print_progress(label + " COMPRESS", nblocks, nblocks, final=True)
    total_payload = int(lens_np.sum())
    # ---- TOC header (no per-block Python: a container holds up to 65 535 blocks)
    if nblocks:
        first = np.concatenate(([0], np.flatnonzero(np.diff(mids_np)) + 1))          # where a run of equal method ids starts
        run_syms: List[int] = mids_np[first].tolist()
        run_lens: List[int] = np.diff(np.concatenate((first, [nblocks]))).tolist()
    else:
        run_syms, run_lens = [], []
    lengths = _huff_lengths(Counter(run_syms))
    enc_tbl, _, _ = _huff_canonical(lengths)
    best_k = _best_rice_k(run_lens)
    toc_header = bytearray()
    toc_header += uleb128_encode(len(run_syms))
    toc_header += uleb128_encode(len(enc_tbl))
    for sym, L in sorted(lengths.items(), key=lambda kv: (kv[1], kv[0])):
        toc_header += uleb128_encode(sym)
        toc_header += uleb128_encode(L)
    toc_header += uleb128_encode(best_k)
    deltas: List[int] = []
    best_k2 = 0
    if mode == MODE_FIXED:
        toc_header += uleb128_encode(boundaries[-1][1] - boundaries[-1][0] if nblocks > 0 else 0)
    else:
        deltas = [_zz_enc((b - a) - size_field) for a, b in boundaries]
        best_k2 = _best_rice_k(deltas)
        toc_header += uleb128_encode(best_k2)
    # ---- TOC bitstream
    bw = _Bits()
    for s in run_syms:
        c, L = enc_tbl[s]
        bw.put(c, L)
    bw.rice(run_lens, best_k)
    if mode == MODE_CDC:
        bw.rice(deltas, best_k2)
    n = nblocks
    l = _ef_choose_l(total_payload, n)
    m = (total_payload + ((1 << l) - 1)) >> l
    total = m + n
    if n and total_payload < (1 << 62) and l <= 62:
        # Elias-Fano of the cumulative payload ends: low l bits of every end, then the unary bitmap with bit (end >> l) + i set,
        # written MSB first into `total` bits — as bit arrays (the per-block big-integer shifts were quadratic in the block count)
        P = np.cumsum(lens_np)
        bw.put_fixed(P & ((1 << l) - 1) if l else P, l)
        hi = (P >> l) + np.arange(n, dtype=np.int64)
        bitmap = np.zeros(total, dtype=np.uint8)
        bitmap[hi] = 1
        bw.put_array(bitmap)
    else:
        P, s = [], 0
        for L in lens_np.tolist():
            s += L
            P.append(s)
        for x in P:
            bw.put(x, l)
        if total:
            bitmap = 0
            for i, x in enumerate(P):
                bitmap |= 1 << (total - 1 - ((x >> l) + i))
            bw.put(bitmap, total)
    toc_bits, toc_bitlen = bw.value()
    out += uleb128_encode(len(toc_header))
    out += uleb128_encode(toc_bitlen)
    out += uleb128_encode(total_payload)
    out += toc_header
    out += toc_bits
    head = bytes(out)
    n_area = len(area)
    if n_area < (32 << 20):
        return b"".join((head, memoryview(area)))       # one copy of the payload area
    # large payload areas: the container is filled in place by several threads (one copy, page faults spread over the cores)
    from .engine import _new_bytes, _par_copy
    blob, sink = _new_bytes(len(head) + n_area)
    sink[:len(head)] = np.frombuffer(head, dtype=np.uint8)
    _par_copy(sink[len(head):], np.frombuffer(memoryview(area), dtype=np.uint8))
    return blob


def compress_blocks_fixed(data: bytes, block_size: int = 8192) -> bytes:
    data = bytes(data)
    return _assemble(data, fixed_boundaries(data, block_size), MODE_FIXED, block_size)


def compress_blocks_cdc(data: bytes, min_size: int = 4096, avg_size: int = 8192, max_size: int = 16384) -> bytes:
    data = bytes(data)
    return _assemble(data, cdc_fast_boundaries_strict(data, min_size, avg_size, max_size), MODE_CDC, avg_size)


def _parse(container: bytes):
    """Header + TOC walk of decompress (v2-2.py:2451-2530): -> (method names, payload starts, payload lengths, orig lengths,
    total_len, end position of the payload area)."""
    if len(container) < 4 or container[:4] != b"KOLR":
        raise ValueError("Invalid magic")
    pos = 4
    packed = struct.unpack_from("<I", container, pos)[0]
    pos += 4
    mode, size_field = (packed >> 31) & 1, packed & 0x7FFFFFFF
    total_len = struct.unpack_from("<I", container, pos)[0]
    pos += 4
    nblocks = struct.unpack_from("<H", container, pos)[0]
    pos += 2
    toc_hdr_len, pos = uleb128_decode_stream(container, pos)
    toc_bitlen, pos = uleb128_decode_stream(container, pos)
    total_payload, pos = uleb128_decode_stream(container, pos)
    if pos + toc_hdr_len > len(container):
        raise ValueError("Truncated TOC header")
    toc_header = container[pos:pos + toc_hdr_len]
    pos += toc_hdr_len
    toc_bit_bytes = (toc_bitlen + 7) // 8
    if pos + toc_bit_bytes > len(container):
        raise ValueError("Truncated TOC bits")
    toc_bits = container[pos:pos + toc_bit_bytes]
    pos += toc_bit_bytes
    p = 0
    n_runs, p = uleb128_decode_stream(toc_header, p)
    K, p = uleb128_decode_stream(toc_header, p)
    lengths = {}
    for _ in range(K):
        sym, p = uleb128_decode_stream(toc_header, p)
        L, p = uleb128_decode_stream(toc_header, p)
        lengths[sym] = L
    k_runs, p = uleb128_decode_stream(toc_header, p)
    if mode == MODE_FIXED:
        last_orig_len, p = uleb128_decode_stream(toc_header, p)
    else:
        k_orig, p = uleb128_decode_stream(toc_header, p)
    _, dec_tbl, maxlen = _huff_canonical(lengths)
    br = _BitsIn(toc_bits)
    run_syms = []
    for _ in range(n_runs):
        c = 0
        for L in range(1, maxlen + 1):
            c = (c << 1) | br.bit()
            if (L, c) in dec_tbl:
                run_syms.append(dec_tbl[(L, c)])
                break
        else:
            raise ValueError("Huffman decode failed")
    run_lens = [br.rice(k_runs) for _ in range(n_runs)]
    method_ids: List[int] = []
    for s, r in zip(run_syms, run_lens):
        method_ids.extend([s] * r)
    if len(method_ids) != nblocks:
        raise ValueError("Method id RLE expands to wrong size")
    if mode == MODE_FIXED:
        orig_lens = [size_field] * (nblocks - 1) + ([last_orig_len] if nblocks > 0 else [])
    else:
        orig_lens = [size_field + _zz_dec(br.rice(k_orig)) for _ in range(nblocks)]
    l = _ef_choose_l(total_payload, nblocks)
    m = (total_payload + ((1 << l) - 1)) >> l
    P = None
    P_np = None
    if nblocks > 64 and l <= 62 and br.pos + nblocks * l + 1 <= br.total:
        # well-formed TOCs: low bits as one reshape, the upper-bit unary code as one flatnonzero (same values as the loops below,
        # which stay for short or damaged TOCs so that their errors are the reference's)
        a = br.unpacked()
        p0 = br.pos
        hi0 = p0 + nblocks * l
        span = a[hi0:hi0 + m + nblocks]
        ones_np = np.flatnonzero(span)[:nblocks]
        if len(ones_np) == nblocks:
            if l:
                w = (1 << np.arange(l - 1, -1, -1, dtype=np.int64))
                lows_np = a[p0:hi0].reshape(nblocks, l).astype(np.int64) @ w
            else:
                lows_np = np.zeros(nblocks, dtype=np.int64)
            P_np = ((ones_np.astype(np.int64) - np.arange(nblocks, dtype=np.int64)) << l) | lows_np
            P = P_np.tolist()
            br.pos = hi0 + int(ones_np[-1]) + 1
    if P is None:
        lows = [br.bits(l) for _ in range(nblocks)]
        ones = []
        for idx in range(m + nblocks):
            if br.bit():
                ones.append(idx)
                if len(ones) == nblocks:
                    break
        P = [((ones[i] - i) << l) | lows[i] for i in range(nblocks)]
    if P and P[-1] != total_payload:
        raise ValueError("Payload EF sum mismatch")
    if pos + total_payload > len(container):
        raise ValueError("Truncated payload area")
    area0 = pos
    pos += total_payload
    if P_np is not None:                                             # many blocks: the per-block lists come out of numpy in one go
        mids_np = np.asarray(method_ids, dtype=np.int64)
        bad_id = np.flatnonzero((mids_np < 0) | (mids_np >= len(KOLR_NAMES)))
        if len(bad_id):
            raise ValueError(f"Unknown method_id {int(mids_np[bad_id[0]])}")
        names = np.array(KOLR_NAMES, dtype=object)[mids_np].tolist()
        starts = (area0 + np.concatenate(([0], P_np[:-1]))).tolist()
        plens_np = np.diff(P_np, prepend=0)
        if bool((plens_np < 0).any()):
            raise ValueError("Payload EF offsets not monotone")
        return names, starts, plens_np.tolist(), orig_lens, total_len, pos
    names = []
    for i in range(nblocks):
        mid = method_ids[i]
        if mid < 0 or mid >= len(KOLR_NAMES):
            raise ValueError(f"Unknown method_id {mid}")
        names.append(KOLR_NAMES[mid])
    starts = [area0] + [area0 + x for x in P[:-1]] if nblocks else []
    plens = [P[0]] + [P[i] - P[i - 1] for i in range(1, nblocks)] if nblocks else []
    if any(x < 0 for x in plens):
        raise ValueError("Payload EF offsets not monotone")
    return names, starts, plens, orig_lens, total_len, pos


def decompress(container: bytes) -> bytes:
    container = bytes(container)
    names, starts, plens, orig_lens, total_len, pos = _parse(container)
    nblocks = len(names)
    _print_progress("DECOMPRESS", 0, nblocks)
    out = _engine().decode_container(container, names, starts, plens, orig_lens) if nblocks else b""
    _print_progress("DECOMPRESS", nblocks, nblocks, final=True)
    if len(out) != total_len:
        raise ValueError(f"Length mismatch: got {len(out)}, expect {total_len}")
    if pos != len(container):
        raise ValueError(f"Extra trailing {len(container) - pos} bytes after container end")
    return out


def main(argv=None) -> int:
    """Command line of the reference (v2-2.py:2624-2692): same flags, defaults, derived FastCDC sizes and messages.  `--experiment`
    (matplotlib plot over random.Random data) is not part of the hot path and is not offered."""
    import argparse
    import os
    global G_NO_LZ77, G_ONLY_METHOD, G_PROGRESS
    ap = argparse.ArgumentParser(description="Kolmogorov researched compressor (GPU hot path)")
    ap.add_argument("-i", "--input", nargs="?", help="Input file to compress or decompress")
    ap.add_argument("-d", "--decompress", action="store_true", help="Decompress")
    ap.add_argument("-o", "--output", help="Output file")
    ap.add_argument("-b", "--block", type=int, default=2048, help="Target block size (FIXED) or avg_size (FastCDC)")
    ap.add_argument("--FastCDC", "--fastcdc", dest="fastcdc", action="store_true",
                    help="Use Fast Content-Defined Chunking (avg_size = --block). When off, use fixed-size chunking.")
    ap.add_argument("--no-lz77", action="store_true", help="Disable LZ77 candidate")
    ap.add_argument("--only", type=str, default=None, help="Only use a single model by name (e.g. lz77, raw)")
    ap.add_argument("--progress", action="store_true", help="Show per-block progress")
    a = ap.parse_args(argv)
    G_NO_LZ77, G_ONLY_METHOD, G_PROGRESS = bool(a.no_lz77), a.only, bool(a.progress)
    if not a.input:
        ap.print_help()
        return 0
    data = open(a.input, "rb").read()
    if a.decompress:
        out = decompress(data)
        name = a.output or (os.path.splitext(a.input)[0] + ".out")
        open(name, "wb").write(out)
        print(f"Decompressed {len(data)} bytes to {len(out)} bytes → {name}")
        return 0
    if a.fastcdc:
        avg = max(64, a.block)                               # v2-2.py:2672-2675
        mn = max(64, min(avg, avg // 2 if avg >= 2 else 64))
        mx = max(avg, avg * 2)
        blob = compress_blocks_cdc(data, min_size=mn, avg_size=avg, max_size=mx)
        mode = f"FastCDC(min={mn}, avg={avg}, max={mx})"
    else:
        blob = compress_blocks_fixed(data, block_size=a.block)
        mode = f"FIXED(block={a.block})"
    name = a.output or (a.input + ".kolr")
    open(name, "wb").write(blob)
    ratio = len(blob) / len(data) if len(data) else 1.0
    print(f"[{mode}] Compressed {len(data)} bytes to {len(blob)} bytes (ratio {ratio:.3f}) → {name}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
