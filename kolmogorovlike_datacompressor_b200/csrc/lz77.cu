// lz77.cu — LZ77 match search + greedy parse, exact reference semantics.
//
//   KF  encode_model_lz77   kolm_final.py:567-617   window 255, match cap 127
//   V22 encode_lz77         kolm_final_researched_v2-2.py:1686-1763   window 4096 (of this block), no cap
//   both: min match 3, overlap allowed, scan distances 1,2,... and keep a STRICTLY longer match
//         => nearest among the longest; greedy parse; tokens [0,byte] / [1,ULEB len,ULEB dist].
//
// Every match of length >= 3 shares its first three bytes with its source, so the candidate set of a
// position is exactly the earlier positions with the same trigram inside the window.  One stable radix
// sort of (trigram, position) per batch (the BBWT engine's sort) lays every trigram class out
// contiguously in position order; a position walks its class backwards = nearest first.
//   1. k_lz_match : per position, best (len, dist) with extension capped at `pcap`; a candidate is only
//                   extended if it can be strictly longer (byte at offset best_len matches).
//   2. k_lz_chunks: per 256-position chunk, for every entry e the exit of the greedy walk and the first
//                   "capped" position met (match reached pcap but not the true limit).
//   3. k_lz_parse : one CTA per block follows the chunk table; capped matches on the parse path are
//                   re-evaluated exactly by the whole CTA (all candidates, full length).
//   4. k_lz_mark / k_lz_emit: token starts, sizes, look-back offsets, byte scatter.
#include "common.cuh"

#define LZ_CHUNK 256
#define LZ_NONE 0xffffffffu
#define LZ_CAPPED 0x80000000u

struct LzArgs {
    const u8* in; const BlockInfo* binfo; const TileDesc* tiles;
    const u32* K; const u32* V;        // sorted (trigram, position)
    u32* rp;                           // position -> index in sorted order
    u32* mlen; u32* mdist;             // per position best match (len may carry LZ_CAPPED)
    u32* cexit; u32* ccap;             // per position: greedy exit of its chunk / first capped position on the way
    u32* centry;                       // per chunk: position where the parse enters (LZ_NONE: skipped)
    u8* tok;                           // per position: encoded size of the token starting here (0: none)
    u32 window, maxlen, pcap;
};

__device__ __forceinline__ u32 uleb_size(u32 v) { return v < 128u ? 1u : v < 16384u ? 2u : v < 2097152u ? 3u : v < 268435456u ? 4u : 5u; }

__global__ void __launch_bounds__(KOLM_THREADS) k_lz_trikeys(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                             const BlockInfo* __restrict__ binfo, u32* __restrict__ K, u32* __restrict__ V) {
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff;
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 pg = td.start + x, lp = pg - bi.pbase;
        u32 key = (lp + 2 < bi.len) ? (((u32)src[lp] << 16) | ((u32)src[lp + 1] << 8) | src[lp + 2]) : 0xffffffu;
        K[pg] = key; V[pg] = pg;
    }
}

__global__ void __launch_bounds__(KOLM_THREADS) k_lz_rankpos(const u32* __restrict__ V, const TileDesc* __restrict__ tiles, u32* __restrict__ rp) {
    TileDesc td = tiles[blockIdx.x];
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) rp[V[td.start + x]] = td.start + x;
}

// common prefix of src[a..] and src[b..] (a < b), starting at offset `from`, at most `upto`
__device__ __forceinline__ u32 lz_extend(const u8* __restrict__ src, u32 a, u32 b, u32 from, u32 upto) {
    u32 m = from;
    while (m < upto && src[a + m] == src[b + m]) ++m;
    return m;
}

__global__ void __launch_bounds__(KOLM_THREADS) k_lz_match(LzArgs a) {
    TileDesc td = a.tiles[blockIdx.x];
    BlockInfo bi = a.binfo[td.block];
    const u8* src = a.in + bi.ioff;
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 pg = td.start + x, lp = pg - bi.pbase;
        u32 lim = bi.len - lp;
        if (a.maxlen && lim > a.maxlen) lim = a.maxlen;
        u32 best = 0, bd = 0; bool capped = false;
        if (lim >= 3) {
            u32 cap = lim < a.pcap ? lim : a.pcap;
            u32 s = a.rp[pg];
            u32 key = a.K[s];
            u32 lo = lp > a.window ? lp - a.window : 0;
            for (u32 t = s; t > bi.pbase; ) {
                --t;
                if (a.K[t] != key) break;
                u32 j = a.V[t] - bi.pbase;
                if (j < lo) break;
                if (best && src[j + best] != src[lp + best]) continue;     // cannot be strictly longer
                u32 m = lz_extend(src, j, lp, 3, cap);
                if (m > best) { best = m; bd = lp - j; if (m == cap) { capped = cap < lim; break; } }
            }
        }
        a.mlen[pg] = best | (capped ? LZ_CAPPED : 0u);
        a.mdist[pg] = bd;
    }
}

// thread per chunk, backwards: exit[e] = first position >= chunk end reached from e; ccap[e] = first capped position met
__global__ void k_lz_chunks(LzArgs a, int nblocks_unused, u32 total_chunks, const u32* __restrict__ chunk_block, const u32* __restrict__ chunk_first) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    u32 b = chunk_block[c];
    BlockInfo bi = a.binfo[b];
    u32 lc = c - chunk_first[b];
    u32 p0 = lc * LZ_CHUNK, p1 = min(p0 + LZ_CHUNK, bi.len);
    for (u32 lp = p1; lp-- > p0; ) {
        u32 pg = bi.pbase + lp;
        u32 ml = a.mlen[pg];
        if (ml & LZ_CAPPED) { a.cexit[pg] = LZ_NONE; a.ccap[pg] = lp; continue; }
        u32 nx = lp + (ml >= 3 ? ml : 1);
        if (nx >= p1) { a.cexit[pg] = nx; a.ccap[pg] = LZ_NONE; }
        else { a.cexit[pg] = a.cexit[bi.pbase + nx]; a.ccap[pg] = a.ccap[bi.pbase + nx]; }
    }
    a.centry[c] = LZ_NONE;
}

// one CTA per block: follow the chunk table; resolve capped matches exactly
__global__ void __launch_bounds__(256) k_lz_parse(LzArgs a, const u32* __restrict__ chunk_first) {
    __shared__ u32 s_pos, s_cap;
    __shared__ unsigned long long s_best;
    const u32 b = blockIdx.x, tid = threadIdx.x;
    BlockInfo bi = a.binfo[b];
    const u8* src = a.in + bi.ioff;
    const u32 cf = chunk_first[b];
    u32 pos = 0;
    while (pos < bi.len) {
        if (tid == 0) {
            u32 c = pos / LZ_CHUNK;
            if (a.centry[cf + c] == LZ_NONE) a.centry[cf + c] = pos;
            s_cap = a.ccap[bi.pbase + pos];
            s_pos = a.cexit[bi.pbase + pos];
            s_best = 0;
        }
        __syncthreads();
        u32 cap = s_cap;
        if (cap == LZ_NONE) { pos = s_pos; __syncthreads(); continue; }
        // exact evaluation of position `cap`: every candidate in the window, full length; best = (max len, min dist)
        u32 lp = cap, pg = bi.pbase + lp;
        u32 lim = bi.len - lp;
        if (a.maxlen && lim > a.maxlen) lim = a.maxlen;
        u32 s = a.rp[pg], key = a.K[s];
        u32 lo = lp > a.window ? lp - a.window : 0;
        const u32 lane = tid & 31, w = tid >> 5;
        // warp w takes candidates q = w, w+8, ... (q-th nearest); lanes compare 32 bytes per step
        for (u32 q = w; q < s - bi.pbase; q += 8) {
            u32 t = s - 1 - q;
            if (a.K[t] != key) break;
            u32 j = a.V[t] - bi.pbase;
            if (j < lo) break;                               // classes are in position order: everything further is outside the window
            u32 m = 0;
            for (;;) {
                u32 o = m + lane;
                bool eq = o < lim && src[j + o] == src[lp + o];
                u32 neq = ~__ballot_sync(0xffffffffu, eq);
                if (neq) { m += __ffs(neq) - 1; break; }
                m += 32;
            }
            if (lane == 0 && m >= 3) atomicMax(&s_best, ((unsigned long long)m << 32) | (unsigned long long)(0xffffffffu - (lp - j)));
        }
        __syncthreads();
        u32 blen = (u32)(s_best >> 32), bdist = 0xffffffffu - (u32)s_best;
        if (tid == 0) { a.mlen[pg] = blen; a.mdist[pg] = bdist; }
        pos = lp + (blen >= 3 ? blen : 1);
        __syncthreads();
    }
}

// thread per chunk: walk from the chunk's entry, mark token starts with their encoded size
__global__ void k_lz_mark(LzArgs a, u32 total_chunks, const u32* __restrict__ chunk_block, const u32* __restrict__ chunk_first) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total_chunks) return;
    u32 b = chunk_block[c];
    BlockInfo bi = a.binfo[b];
    u32 lc = c - chunk_first[b];
    u32 p1 = min((lc + 1) * LZ_CHUNK, bi.len);
    u32 lp = a.centry[c];
    if (lp == LZ_NONE) return;
    while (lp < p1) {
        u32 pg = bi.pbase + lp;
        u32 ml = a.mlen[pg] & ~LZ_CAPPED;
        if (ml >= 3) { a.tok[pg] = (u8)(1 + uleb_size(ml) + uleb_size(a.mdist[pg])); lp += ml; }
        else { a.tok[pg] = 2; lp += 1; }
    }
}

// sizes per block (pass 0) or byte scatter (pass 1)
template <bool EMIT>
__global__ void __launch_bounds__(KOLM_THREADS) k_lz_emit(LzArgs a, u64* lb, u64* __restrict__ bacc, u8* __restrict__ out, const i64* __restrict__ cap_total, u64 cap) {
    if (EMIT && (u64)*cap_total > cap) return;              // exact total known before any byte is emitted: never write past the caller's buffer
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = EMIT ? lb_take_ticket(lb) : blockIdx.x;
    if (tile == LB_NO_TILE) return;
    const TileDesc td = a.tiles[tile];
    const BlockInfo bi = a.binfo[td.block];
    u32 sz[KOLM_IPT]; u64 sum = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { u32 r = tid * KOLM_IPT + i; sz[i] = r < td.count ? a.tok[td.start + r] : 0; sum += sz[i]; }
    u64 tot;
    u64 incl = block_scan_incl(sum, 0ull, OpAdd(), s_warp, &tot);
    if (!EMIT) {
        if (tid == 0 && tot) atomicAdd((unsigned long long*)(bacc + (size_t)td.block * 64 + 32), (unsigned long long)tot);
        return;
    }
    if (tid < 32) {
        u64 e = lb_exclusive(lb, tile, (td.flags & 1u) != 0, tot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u8* dst = out + bacc[(size_t)td.block * 64 + 33] + s_excl + (incl - sum);
    const u8* src = a.in + bi.ioff;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if (sz[i]) {
            u32 pg = td.start + tid * KOLM_IPT + i;
            if (sz[i] == 2 && (a.mlen[pg] & ~LZ_CAPPED) < 3) { dst[0] = 0; dst[1] = src[pg - bi.pbase]; }
            else {
                u32 ml = a.mlen[pg] & ~LZ_CAPPED, md = a.mdist[pg];
                u8* p = dst; *p++ = 1;
                while (ml >= 128) { *p++ = (u8)(ml | 0x80); ml >>= 7; } *p++ = (u8)ml;
                while (md >= 128) { *p++ = (u8)(md | 0x80); md >>= 7; } *p++ = (u8)md;
            }
            dst += sz[i];
        }
    }
}

// single CTA: payload offsets from the per-block sizes in bacc[b*64+32]
__global__ void k_lz_plan(u64* __restrict__ bacc, i64* __restrict__ poff, int nblocks) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        int b = base + threadIdx.x;
        u64 bytes = b < nblocks ? bacc[(size_t)b * 64 + 32] : 0;
        u64 v = bytes;
        for (int o = 1; o < 32; o <<= 1) { u64 n = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (u32)o) v += n; }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = v;
        __syncthreads();
        u64 pre = 0;
        for (u32 i = 0; i < (threadIdx.x >> 5); ++i) pre += s_w[i];
        u64 carry = s_carry;
        if (b < nblocks) { poff[b] = (i64)(carry + pre + v - bytes); bacc[(size_t)b * 64 + 33] = carry + pre + v - bytes; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) poff[nblocks] = (i64)s_carry;
}

__global__ void k_lz_chunkmap(const BlockInfo* __restrict__ binfo, const u32* __restrict__ cfirst, u32* __restrict__ cblock, int nblocks) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    u32 n = (binfo[b].len + LZ_CHUNK - 1) / LZ_CHUNK, f = cfirst[b];
    for (u32 k = 0; k < n; ++k) cblock[f + k] = b;
}

// decode (v0): one thread per block  (KF.py:723-760 / V22.py:1765-1812)
__global__ void k_lz_dec(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo, u8* __restrict__ out,
                         int* __restrict__ err, int nblocks, u32 window_check) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    BlockInfo bi = binfo[b];
    const u8* d = pay + pay_off[b];
    i64 n = pay_off[b + 1] - pay_off[b], i = 0;
    u8* dst = out + bi.ioff;
    u32 o = 0; int e = KOLM_OK;
    while (i < n && o < bi.len) {
        u8 flag = d[i++];
        if (flag == 0) { if (i >= n) { e = KOLM_E_TRUNCATED; break; } dst[o++] = d[i++]; }
        else if (flag == 1) {
            u64 v[2];
            for (int q = 0; q < 2 && !e; ++q) {
                u64 r = 0; int sh = 0;
                for (;;) { if (i >= n) { e = KOLM_E_TRUNCATED; break; } u8 x = d[i++]; if (sh < 64) r |= (u64)(x & 0x7F) << sh; if (!(x & 0x80)) break; sh += 7; }
                v[q] = r;
            }
            if (e) break;
            u64 len = v[0], dist = v[1];
            if (dist == 0) {
                // V22 rejects distance 0 (V22.py:1794-1795); KF has no such check: `out[-0]` is out[0], so the token repeats the
                // first output byte, and raises IndexError only while the output is still empty (KF.py:745-752)
                if (window_check) { e = KOLM_E_CORRUPT; break; }
                if (len && o == 0 && bi.len) { e = KOLM_E_INDEX; break; }
                for (u64 t = 0; t < len && o < bi.len; ++t) dst[o++] = dst[0];
                continue;
            }
            for (u64 t = 0; t < len && o < bi.len; ++t) {
                u32 avail = window_check ? min(o, window_check) : o;
                if (dist > avail) { e = KOLM_E_CORRUPT; break; }
                dst[o] = dst[o - dist]; ++o;
            }
            if (e) break;
        } else { e = KOLM_E_CORRUPT; break; }
    }
    if (!e && o != bi.len) e = KOLM_E_CORRUPT;
    err[b] = e;
}


// decode: one warp per block.  Tokens are parsed in order, but a match is copied by all 32 lanes at once: the copied region is
// periodic with period `dist`, so byte k of the match equals out[o - dist + (k mod dist)] — every lane reads only bytes that
// were complete before the token started, overlapping matches included.
__global__ void __launch_bounds__(128) k_lz_dec_warp(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo,
                                                     u8* __restrict__ out, int* __restrict__ err, int nblocks, u32 window_check) {
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const BlockInfo bi = binfo[b];
    const u8* d = pay + pay_off[b];
    const i64 n = pay_off[b + 1] - pay_off[b];
    u8* dst = out + bi.ioff;
    i64 i = 0; u32 o = 0; int e = KOLM_OK;
    while (i < n && o < bi.len) {
        // every lane parses the (short) token header redundantly from the same addresses (broadcast loads)
        u8 flag = d[i++];
        if (flag == 0) {
            if (i >= n) { e = KOLM_E_TRUNCATED; break; }
            if (lane == 0) dst[o] = d[i];
            ++i; ++o;
        } else if (flag == 1) {
            u64 v[2];
            for (int q = 0; q < 2 && !e; ++q) {
                u64 r = 0; int sh = 0;
                for (;;) { if (i >= n) { e = KOLM_E_TRUNCATED; break; } u8 x = d[i++]; if (sh < 64) r |= (u64)(x & 0x7F) << sh; if (!(x & 0x80)) break; sh += 7; }
                v[q] = r;
            }
            if (e) break;
            u64 len = v[0], dist = v[1];
            u32 avail = window_check ? min(o, window_check) : o;
            u64 room = (u64)(bi.len - o);
            u32 cnt = (u32)(len < room ? len : room);
            if (dist == 0) {                                                // see k_lz_dec: V22 rejects, KF repeats out[0] (IndexError on empty output)
                if (window_check) { e = KOLM_E_CORRUPT; break; }
                if (cnt && o == 0) { e = KOLM_E_INDEX; break; }
                __syncwarp();
                const u8 first = dst[0];
                for (u32 k = lane; k < cnt; k += 32) dst[o + k] = first;
                o += cnt;
                __syncwarp();
                continue;
            }
            if (cnt && dist > avail) { e = KOLM_E_CORRUPT; break; }       // reference: "distance beyond window" / "Invalid LZ77 distance"
            __syncwarp();                                                  // earlier literal / match stores are visible to all lanes
            const u32 dd = (u32)dist;
            const u8* srcp = dst + o - dd;
            if (dd >= 32) { for (u32 k = lane; k < cnt; k += 32) dst[o + k] = srcp[k % dd]; }
            else { for (u32 k = lane; k < cnt; k += 32) dst[o + k] = srcp[k % dd]; }
            o += cnt;
            __syncwarp();
        } else { e = KOLM_E_CORRUPT; break; }
    }
    if (!e && o != bi.len) e = KOLM_E_CORRUPT;
    if (lane == 0) err[b] = e;
}

// decode, v2: the same warp-per-block decoder with the token stream read through a REGISTER window — lane l holds four
// bytes of a 128-byte window of the payload, a byte is one shuffle away — instead of one dependent global load per header byte
// (the serial token chain was ~500 cycles per token: 53 ms for a 1 MiB block of short tokens, 88 % of KOLR decompress at 1 MiB
// blocks), runs of literal tokens (flag 0, byte) decoded up to sixteen at a time (their layout is fixed, so a ballot over the
// even offsets finds the run), and no modulo in the copy of a non-overlapping match.  The last 8 KiB of the output are mirrored
// in a shared-memory ring per warp: a match reads what the previous tokens just stored, and through global memory that is an
// L2 round trip per token (~600 cycles on blocks of short matches: 50 ms per MiB); matches of up to 2 KiB at distances of up to
// 6 KiB — all of them with the reference's windows of 255 and 4096 — read the ring instead.  The window path runs while at least
// 160 payload bytes lie ahead; the block's tail goes through the byte-wise code of v1, whose error behaviour is the reference's.
#define LZ_RING 8192u
#define LZ_RING_MAXLEN 2048u
__global__ void __launch_bounds__(128) k_lz_dec_warp2(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo,
                                                      u8* __restrict__ out, int* __restrict__ err, int nblocks, u32 window_check) {
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const BlockInfo bi = binfo[b];
    const u8* d = pay + pay_off[b];
    const i64 n = pay_off[b + 1] - pay_off[b];
    u8* dst = out + bi.ioff;
    i64 i = 0; u32 o = 0; int e = KOLM_OK;
    __shared__ u8 s_ring[4][LZ_RING];
    u8* ring = s_ring[threadIdx.x >> 5];
    // window: 32 aligned words starting at payload index wbase (wbase + misalignment of d is a multiple of 4)
    const u32 mis = (u32)((uintptr_t)d & 3u);
    i64 wbase = -(i64)mis - 256;                             // "no window"
    u32 wword = 0;
    while (i + 160 <= n && o < bi.len) {
        if (i - wbase > 60) {                                // refill so that 68 bytes from i on are inside the window
            wbase = ((i + mis) & ~(i64)3) - mis;
            wword = *reinterpret_cast<const u32*>(d + wbase + 4 * lane);     // aligned; at most 128 + 3 bytes past i, inside the payload
        }
#define LZ_GB(j) ((__shfl_sync(0xffffffffu, wword, (u32)((j) - wbase) >> 2) >> (8u * ((u32)((j) - wbase) & 3u))) & 0xFFu)
        // ---- all tokens that start in the next 32 payload bytes at once.  Lane l assumes a token starts at byte i + l: a literal
        //      is two bytes, a match 1 + the two ULEB lengths, which are runs of continuation bits in a 64-bit ballot.  The real
        //      starts are the lanes reachable from lane 0 by "next = l + token length" (five doubling steps); their output
        //      lengths are scanned, the literals stored in one step and the matches copied in order.  Anything unusual (bad
        //      flag, value over four bytes, distance 0 or beyond the window, output nearly full) leaves the window to the
        //      token-at-a-time code below, whose checks are the reference's.
        // (a run of eight or more literal tokens ahead is cheaper through the literal path below: one ballot instead of ~30 collectives)
        bool litrun;
        {
            const u32 fb = LZ_GB(i + 2 * (i64)(lane & 15));
            const u32 zm = __ballot_sync(0xffffffffu, lane < 16 && fb == 0);
            litrun = (u32)(__ffs(~zm) - 1) >= 8u;
        }
        if (!litrun) {
            // my nine bytes (payload i + lane .. i + lane + 8) from three words of the register window: everything about "the
            // token that would start at my byte" is then lane-local arithmetic
            u32 blo, bhi, b8;
            {
                const u32 off = (u32)(i - wbase) + lane, wi = off >> 2, sh = 8u * (off & 3u);
                const u32 w0 = __shfl_sync(0xffffffffu, wword, wi), w1 = __shfl_sync(0xffffffffu, wword, wi + 1), w2 = __shfl_sync(0xffffffffu, wword, wi + 2);
                blo = __funnelshift_r(w0, w1, sh); bhi = __funnelshift_r(w1, w2, sh); b8 = (sh ? (w2 >> sh) : w2) & 0xFFu;   // wi + 2 <= 31: off <= 60 + 31
            }
            const u32 b0 = blo & 0xFFu;
            u32 tl = 1, kind = 2, mlen = 0, mdist = 0;
            if (b0 == 0) { tl = 2; kind = 0; }
            else if (b0 == 1) {
                // bytes 1..8 as one 64-bit word; a ULEB value ends at its first byte without the continuation bit
                const u64 v8 = ((u64)(blo >> 8)) | ((u64)bhi << 24) | ((u64)b8 << 56);
                const u64 stop = ~v8 & 0x8080808080808080ull;
                const u32 l1 = stop ? (u32)(__ffsll((long long)stop) >> 3) : 9u;               // bytes of the first value
                if (l1 <= 4) {
                    const u64 v2 = v8 >> (8 * l1), stop2 = stop >> (8 * l1);
                    const u32 l2 = stop2 ? (u32)(__ffsll((long long)stop2) >> 3) : 9u;
                    if (l2 <= 4) {
                        tl = 1 + l1 + l2; kind = 1;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if ((u32)q < l1) mlen |= ((u32)(v8 >> (8 * q)) & 0x7Fu) << (7 * q);
                            if ((u32)q < l2) mdist |= ((u32)(v2 >> (8 * q)) & 0x7Fu) << (7 * q);
                        }
                    }
                }
            }
            const u32 nx1 = lane + tl;
            u32 R = 1u, j = nx1;
#pragma unroll
            for (int st = 0; st < 4; ++st) {                 // valid tokens are at least two bytes: at most sixteen start in the window, and a chain that meets a bad one stops there
                const u32 tgt = (((R >> lane) & 1u) && j < 32) ? (1u << j) : 0u;
                R |= __reduce_or_sync(0xffffffffu, tgt);
                const u32 jj = __shfl_sync(0xffffffffu, j, j & 31u);
                j = j < 32 ? jj : j;
            }
            const bool start = (R >> lane) & 1u;
            const u32 outlen = !start ? 0u : kind == 0 ? 1u : mlen;
            u32 incl = outlen;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, sh); if (lane >= (u32)sh) incl += t; }
            const u32 excl = incl - outlen, total = __shfl_sync(0xffffffffu, incl, 31);
            const u32 om = o + excl;
            const u32 avail = window_check ? min(om, window_check) : om;
            const bool odd = start && (kind == 2 || (kind == 1 && (mdist == 0 || (mlen && mdist > avail))));
            // (total <= 2 KiB: every match takes the ring path below, and no store of this window can reach a ring slot that one of
            //  its matches still reads — sources lie at most 6 KiB behind, the ring holds 8)
            if (!__any_sync(0xffffffffu, odd) && (u64)o + total <= (u64)bi.len && total <= LZ_RING_MAXLEN) {
                if (start && kind == 0) { const u8 v = (u8)(blo >> 8); dst[om] = v; ring[om & (LZ_RING - 1)] = v; }
                u32 mm = __ballot_sync(0xffffffffu, start && kind == 1 && mlen);
                while (mm) {
                    const u32 l = __ffs(mm) - 1; mm &= mm - 1;
                    const u32 cnt = __shfl_sync(0xffffffffu, mlen, l), dd = __shfl_sync(0xffffffffu, mdist, l), oo = __shfl_sync(0xffffffffu, om, l);
                    __syncwarp();
                    if (cnt <= LZ_RING_MAXLEN && dd <= LZ_RING - LZ_RING_MAXLEN) {
                        const u32 sb = oo - dd;
                        if (dd >= cnt) { for (u32 k = lane; k < cnt; k += 32) { const u8 v = ring[(sb + k) & (LZ_RING - 1)]; dst[oo + k] = v; ring[(oo + k) & (LZ_RING - 1)] = v; } }
                        else { for (u32 k = lane; k < cnt; k += 32) { const u8 v = ring[(sb + k % dd) & (LZ_RING - 1)]; dst[oo + k] = v; ring[(oo + k) & (LZ_RING - 1)] = v; } }
                    } else {
                        const u8* srcp = dst + oo - dd;
                        if (dd >= cnt) { for (u32 k = lane; k < cnt; k += 32) { const u8 v = srcp[k]; dst[oo + k] = v; ring[(oo + k) & (LZ_RING - 1)] = v; } }
                        else { for (u32 k = lane; k < cnt; k += 32) { const u8 v = srcp[k % dd]; dst[oo + k] = v; ring[(oo + k) & (LZ_RING - 1)] = v; } }
                    }
                }
                __syncwarp();
                const u32 last = 31u - __clz(R);             // the last token that starts in the window
                i += (i64)__shfl_sync(0xffffffffu, nx1, last);
                o += total;
                continue;
            }
        }
        const u32 flag = LZ_GB(i);
        if (flag == 0) {
            // literal tokens: (0, byte) pairs; lanes 0..15 look at the even offsets
            const u32 fb = LZ_GB(i + 2 * (i64)(lane & 15));
            const u32 vb = LZ_GB(i + 2 * (i64)(lane & 15) + 1);
            const u32 zm = __ballot_sync(0xffffffffu, lane < 16 && fb == 0);
            u32 t = __ffs(~zm) - 1;                          // consecutive literal tokens from i (>= 1, <= 16)
            const u32 room = bi.len - o;
            if (t > room) t = room;
            if (lane < t) { dst[o + lane] = (u8)vb; ring[(o + lane) & (LZ_RING - 1)] = (u8)vb; }
            i += 2 * (i64)t; o += t;
        } else if (flag == 1) {
            u64 v[2]; i64 p = i + 1;
            bool longv = false;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                u64 r = 0; int sh = 0;
                for (;;) {
                    if (p - i >= 36) { longv = true; break; }               // absurdly padded value: let v1's code deal with it
                    const u32 x = LZ_GB(p); ++p;
                    if (sh < 64) r |= (u64)(x & 0x7F) << sh;
                    if (!(x & 0x80)) break;
                    sh += 7;
                }
                v[q] = r;
                if (longv) break;
            }
            if (longv) break;                                // (i unchanged)
            i = p;
            const u64 len = v[0], dist = v[1];
            const u32 avail = window_check ? min(o, window_check) : o;
            const u64 room = (u64)(bi.len - o);
            const u32 cnt = (u32)(len < room ? len : room);
            if (dist == 0) {                                                // see k_lz_dec: V22 rejects, KF repeats out[0] (IndexError on empty output)
                if (window_check) { e = KOLM_E_CORRUPT; break; }
                if (cnt && o == 0) { e = KOLM_E_INDEX; break; }
                __syncwarp();
                const u8 first = dst[0];
                for (u32 k = lane; k < cnt; k += 32) { dst[o + k] = first; ring[(o + k) & (LZ_RING - 1)] = first; }
                o += cnt;
                __syncwarp();
                continue;
            }
            if (cnt && dist > avail) { e = KOLM_E_CORRUPT; break; }
            __syncwarp();                                                  // earlier literal / match stores are visible to all lanes
            const u32 dd = (u32)dist;
            if (cnt <= LZ_RING_MAXLEN && dd <= LZ_RING - LZ_RING_MAXLEN) {
                // source and destination both inside the ring and disjoint there (dd + cnt <= LZ_RING): shared-memory latency only
                const u32 sb = o - dd;
                if (dd >= cnt) { for (u32 k = lane; k < cnt; k += 32) { const u8 v = ring[(sb + k) & (LZ_RING - 1)]; dst[o + k] = v; ring[(o + k) & (LZ_RING - 1)] = v; } }
                else { for (u32 k = lane; k < cnt; k += 32) { const u8 v = ring[(sb + k % dd) & (LZ_RING - 1)]; dst[o + k] = v; ring[(o + k) & (LZ_RING - 1)] = v; } }
            } else {
                const u8* srcp = dst + o - dd;
                if (dd >= cnt) { for (u32 k = lane; k < cnt; k += 32) { const u8 v = srcp[k]; dst[o + k] = v; ring[(o + k) & (LZ_RING - 1)] = v; } }
                else { for (u32 k = lane; k < cnt; k += 32) { const u8 v = srcp[k % dd]; dst[o + k] = v; ring[(o + k) & (LZ_RING - 1)] = v; } }
            }
            o += cnt;
            __syncwarp();
        } else { e = KOLM_E_CORRUPT; break; }
#undef LZ_GB
    }
    __syncwarp();
    // ---- tail (and anything unusual): v1's byte-wise token loop
    while (!e && i < n && o < bi.len) {
        u8 flag = d[i++];
        if (flag == 0) {
            if (i >= n) { e = KOLM_E_TRUNCATED; break; }
            if (lane == 0) dst[o] = d[i];
            ++i; ++o;
        } else if (flag == 1) {
            u64 v[2];
            for (int q = 0; q < 2 && !e; ++q) {
                u64 r = 0; int sh = 0;
                for (;;) { if (i >= n) { e = KOLM_E_TRUNCATED; break; } u8 x = d[i++]; if (sh < 64) r |= (u64)(x & 0x7F) << sh; if (!(x & 0x80)) break; sh += 7; }
                v[q] = r;
            }
            if (e) break;
            u64 len = v[0], dist = v[1];
            u32 avail = window_check ? min(o, window_check) : o;
            u64 room = (u64)(bi.len - o);
            u32 cnt = (u32)(len < room ? len : room);
            if (dist == 0) {
                if (window_check) { e = KOLM_E_CORRUPT; break; }
                if (cnt && o == 0) { e = KOLM_E_INDEX; break; }
                __syncwarp();
                const u8 first = dst[0];
                for (u32 k = lane; k < cnt; k += 32) dst[o + k] = first;
                o += cnt;
                __syncwarp();
                continue;
            }
            if (cnt && dist > avail) { e = KOLM_E_CORRUPT; break; }
            __syncwarp();
            const u32 dd = (u32)dist;
            const u8* srcp = dst + o - dd;
            for (u32 k = lane; k < cnt; k += 32) dst[o + k] = srcp[k % dd];
            o += cnt;
            __syncwarp();
        } else { e = KOLM_E_CORRUPT; break; }
    }
    if (!e && o != bi.len) e = KOLM_E_CORRUPT;
    if (lane == 0) err[b] = e;
}

// out_off == nullptr: device mode (kolm_encode_blocks) — the offsets stay in c->d_poff, nothing comes home, no synchronisation
int kolm_lz77_enc_impl(kolm_ctx* c, const u8* in, u32 window, u32 maxlen, u8* out, size_t out_cap, i64* out_off, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    if (!nb) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    if (window == 0) return KOLM_E_ARG;
    const i64 N = c->total_bytes;
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * 64 * 8, s));
    if (nt) {
        // chunk tables: first chunk of each block (host), owner block of each chunk (device)
        u32 total_chunks = 0;
        u32* cfirst = c->h_u32 + 2 * (size_t)nb;
        for (int b = 0; b < nb; ++b) { cfirst[b] = total_chunks; total_chunks += (c->h_binfo[b].len + LZ_CHUNK - 1) / LZ_CHUNK; }
        if ((size_t)total_chunks > (size_t)c->max_tiles * 128) return KOLM_E_CAPACITY;
        u32* d_cfirst = c->d_atile0;                               // [nb]   (free outside the sort rounds)
        u32* d_centry = c->d_thist;                                // [total_chunks]
        u32* d_cblock = c->d_thist + (size_t)c->max_tiles * 128;   // [total_chunks] second half of the histogram scratch
        CUDA_TRY(cudaMemcpyAsync(d_cfirst, cfirst, (size_t)nb * 4, cudaMemcpyHostToDevice, s));
        KL(c, KC_MISC, N * 9, s, k_lz_trikeys<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_k0, c->d_v0));
        u32 *K, *V;
        KOLM_TRY(radix_sort(c, c->d_tiles, nt, N, c->d_btile0, c->d_btilen, 24, c->d_k0, c->d_v0, c->d_k1, c->d_v1, &K, &V, s));
        u32* Kfree = (K == c->d_k0) ? c->d_k1 : c->d_k0;
        u32* Vfree = (V == c->d_v0) ? c->d_v1 : c->d_v0;
        LzArgs a;
        a.in = in; a.binfo = c->d_binfo; a.tiles = c->d_tiles; a.K = K; a.V = V; a.rp = c->d_rank; a.mlen = c->d_sa; a.mdist = c->d_nr;
        a.cexit = Kfree; a.ccap = Vfree; a.centry = d_centry; a.tok = c->d_tmp8a; a.window = window; a.maxlen = maxlen;
        a.pcap = maxlen ? maxlen : 256u;
        KL(c, KC_MISC, N * 8, s, k_lz_rankpos<<<nt, KOLM_THREADS, 0, s>>>(V, c->d_tiles, c->d_rank));
        KL(c, KC_MISC, N * 20, s, k_lz_match<<<nt, KOLM_THREADS, 0, s>>>(a));
        KL(c, KC_MISC, (i64)total_chunks * 4, s, k_lz_chunkmap<<<(nb + 127) / 128, 128, 0, s>>>(c->d_binfo, d_cfirst, d_cblock, nb));
        CUDA_TRY(cudaMemsetAsync(c->d_tmp8a, 0, c->total_elems, s));
        KL(c, KC_MISC, N * 16, s, k_lz_chunks<<<(total_chunks + 127) / 128, 128, 0, s>>>(a, nb, total_chunks, d_cblock, d_cfirst));
        KL(c, KC_MISC, N, s, k_lz_parse<<<nb, 256, 0, s>>>(a, d_cfirst));
        KL(c, KC_MISC, N * 9, s, k_lz_mark<<<(total_chunks + 127) / 128, 128, 0, s>>>(a, total_chunks, d_cblock, d_cfirst));
        KL(c, KC_MISC, N, s, k_lz_emit<false><<<nt, KOLM_THREADS, 0, s>>>(a, c->d_lb, c->d_bacc, out, c->d_poff + nb, (u64)out_cap));
        KL(c, KC_RICE_PLAN, (i64)nb * 16, s, k_lz_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_poff, nb));
        int lgrid = nt;
        KOLM_TRY(kolm_lb_reset_mode(c, false, nt, &lgrid, 1, s));
        KL(c, KC_MISC, N * 10, s, k_lz_emit<true><<<lgrid, KOLM_THREADS, 0, s>>>(a, c->d_lb, c->d_bacc, out, c->d_poff + nb, (u64)out_cap));
    } else {
        KL(c, KC_RICE_PLAN, (i64)nb * 16, s, k_lz_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_poff, nb));
    }
    CUDA_TRY(cudaGetLastError());
    if (!out_off) return KOLM_OK;
    CUDA_TRY(cudaMemcpyAsync(c->h_poff, c->d_poff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_off, c->h_poff, (size_t)(nb + 1) * 8);
    if ((size_t)out_off[nb] > out_cap) return KOLM_E_CAPACITY;    // callers size `out` at 2*n + 16: the worst case is all literals
    return KOLM_OK;
}

int kolm_lz77_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, u32 window_check, u8* out, cudaStream_t s) {
    const int nb = c->nblocks;
    if (!nb) return KOLM_OK;
    memcpy(c->h_poff, pay_off, (size_t)(nb + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    static int warpdec = -1;
    if (warpdec < 0) { const char* e = getenv("KOLM_LZ_DEC_WARP"); warpdec = e ? atoi(e) : 2; }
    if (warpdec >= 2) KL(c, KC_MISC, c->total_bytes * 2, s, k_lz_dec_warp2<<<(nb + 3) / 4, 128, 0, s>>>(pay, c->d_poff, c->d_binfo, out, c->d_err, nb, window_check));
    else if (warpdec) KL(c, KC_MISC, c->total_bytes * 2, s, k_lz_dec_warp<<<(nb + 3) / 4, 128, 0, s>>>(pay, c->d_poff, c->d_binfo, out, c->d_err, nb, window_check));
    else KL(c, KC_MISC, c->total_bytes * 2, s, k_lz_dec<<<(nb + 63) / 64, 64, 0, s>>>(pay, c->d_poff, c->d_binfo, out, c->d_err, nb, window_check));
    CUDA_TRY(cudaGetLastError());
    return rice_dec_finish(c, s);
}
