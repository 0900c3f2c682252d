// repair.cu — Re-Pair grammar candidate, exact reference semantics.
//
//   repair_compress   kolm_final_researched_v2-2.py:1841-1911  (_count_pairs :1817, _replace_non_overlapping :1824)
//   repair_decompress kolm_final_researched_v2-2.py:1916-1978
//
// Per round: count ALL adjacent (overlapping) pairs, take the most frequent (count >= 2, ties -> the
// lexicographically smallest pair), replace its occurrences left to right without overlap by a new
// symbol 256+r; stop when fewer than 2 replacements happened (that rule is not recorded).
//
// One CTA per block, the whole symbol sequence in shared memory (blocks of up to REPAIR_MAX symbols):
//   pair histogram  = open-addressing hash table in shared memory (atomicCAS insert, atomicAdd count)
//   argmax          = 64-bit atomicMax of (count << 32 | ~pair)   -> max count, smallest pair
//   replacement     = inside a maximal run of overlapping matches (only when a == b) every other one
//                     from the run start is taken; then a block-wide compaction.
// Larger blocks are passed over here and taken by the incremental kernel of repair_big.cu (same output, global-memory slabs).
#include "common.cuh"

#define REPAIR_MAX 8192
#define REPAIR_THREADS 1024
#define REPAIR_HASH 16384          // slots; load factor <= 0.5
#define REPAIR_EMPTY 0xffffffffu

// Three shapes of the same kernel, eight symbols per thread, 6 bytes of table per slot + 2 per symbol, 32 registers:
//   up to 16384 symbols  1024 threads, 224 KB   one CTA per SM (sixteen symbols per thread)
//   up to 8192 symbols   1024 threads, 115 KB   two CTAs per SM
//   up to 4096 symbols    512 threads,  58 KB   three
//   up to 2048 symbols    256 threads,  29 KB   seven   (the reference's default block size)
// The rounds are bound by their block barriers and the table sweep, which shrink with the CTA, and more blocks are in flight
// (61 440 blocks of 2 KiB: 131 ms per 32 MiB on the one-CTA shape, 30 on the first small shape, 7 with the early stop).
#define REPAIR_SMALL 2048
#define REPAIR_MID 4096
#define REPAIR_XL 16384             // the largest shape: 16 symbols per thread, 32 768 slots, 224 KB — the CDC default's longest block
#define REPAIR_XL_HASH 32768
template <int MAXLEN, int THREADS, int HASH>
struct RepairSmemT {
    u16 seq[MAXLEN];                 // compacted in place (a thread's symbols wait in registers across the scan's barriers)
    u32 hkey[HASH];
    u32 hcnt2[HASH / 2];             // 16-bit counts (<= MAXLEN - 1), two per word: 6 bytes per slot let two 8 KiB CTAs share an SM
    u32 scan[THREADS / 32];
    u32 wlast[THREADS / 32];
    u32 wt[THREADS / 32];
    unsigned long long best;
    u32 nhigh, m, hdr, wsum;
};
typedef RepairSmemT<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH> RepairSmem;
static_assert(sizeof(RepairSmemT<16384, 1024, 32768>) <= 227 * 1024, "the 16 KiB shape must fit one CTA's shared memory");

template <int HASH>
__device__ __forceinline__ u32 rp_hash(u32 k) { k *= 2654435761u; return (k >> 15) & (HASH - 1); }

// block-wide inclusive scan (sum or max) of one u32 per thread
template <bool MAXOP, int THREADS>
__device__ __forceinline__ u32 rp_scan(u32 v, u32* s_w, u32* total) {
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v = MAXOP ? max(v, n) : v + n; }
    if (lane == 31) s_w[w] = v;
    __syncthreads();
    u32 pre = 0, tot = 0;
    for (int i = 0; i < THREADS / 32; ++i) { u32 x = s_w[i]; if ((u32)i < w) pre = MAXOP ? max(pre, x) : pre + x; tot = MAXOP ? max(tot, x) : tot + x; }
    __syncthreads();
    if (total) *total = tot;
    return MAXOP ? max(pre, v) : pre + v;
}

__device__ __forceinline__ u32 rp_uleb_size(u32 v) { return v < 128u ? 1u : v < 16384u ? 2u : 3u; }
__device__ __forceinline__ u8* rp_put_uleb(u8* p, u32 v) { while (v >= 128) { *p++ = (u8)(v | 0x80); v >>= 7; } *p++ = (u8)v; return p; }

// out_tmp: per block a staging region of 4*len+64 bytes at tmp + 4*pbase; sizes[b] = payload bytes
template <int MAXLEN, int THREADS, int HASH, int MINLEN>
__global__ void __launch_bounds__(THREADS, MAXLEN > REPAIR_MAX ? 1 : 2048 / THREADS) k_repair_enc(const u8* __restrict__ in, const BlockInfo* __restrict__ binfo, u8* __restrict__ tmp,
                                                        u32* __restrict__ rules_scratch, u64* __restrict__ bacc, int* __restrict__ err,
                                                        const i64* __restrict__ limit) {
    extern __shared__ __align__(16) u8 smem_raw[];
    typedef RepairSmemT<MAXLEN, THREADS, HASH> SM;
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    const u32 tid = threadIdx.x, b = blockIdx.x;
    const BlockInfo bi = binfo[b];
    if (bi.len <= MINLEN && MINLEN) return;                 // the small shape of this kernel takes these
    if (bi.len > MAXLEN) { if (MAXLEN == REPAIR_XL && tid == 0) { err[b] = KOLM_E_UNSUPPORTED; bacc[(size_t)b * 64 + 32] = 0; } return; }   // the incremental kernel's
    const u8* src = in + bi.ioff;
    u32* rules = rules_scratch + bi.pbase;                  // up to len/2 rules, (a<<16|b)
    u32 high = 0;                                            // bytes >= 128: two ULEB bytes each
    for (u32 i = tid; i < bi.len; i += THREADS) { const u8 v = src[i]; S.seq[i] = v; high += v >> 7; }
    if (tid == 0) { S.m = bi.len; S.nhigh = 0; }
    u32 nrules = 0;
    // Early stop (kolm_encode_blocks, limit != nullptr): limit[b] = the smallest size among the block's other candidates, all of
    // which precede Re-Pair in the list, so Re-Pair is selected only if its payload is SMALLER.  Whatever the remaining rounds do,
    //   payload >= 6 + (bytes of the rules made so far) + sum over the DISTINCT adjacent pairs (x, y) of the current sequence of
    //              uleb(x) + 1:
    // in the final grammar every adjacency of the current sequence sits at one boundary — between two neighbours of the final
    // sequence or between the two sides of a later rule — and a boundary determines its pair (last symbol of the left expansion,
    // first of the right), so distinct pairs have distinct boundaries; the left symbol of a boundary is x itself or a later
    // nonterminal (a larger id: at least as many ULEB bytes), each such symbol is written once per boundary, the last symbol of the
    // final sequence costs a byte more, and 'RP', ULEB(256) and the two counts take at least 6.  Once the bound reaches limit[b]
    // the candidate cannot win: the block is marked (bacc slot 34) and its size reported as "not evaluated".  (The code adds
    // (m - 1 - D) / (f - 1) for the rules that are still needed when only D of the m - 1 adjacencies are distinct: see the test.)
    // A second bound from the byte count itself.  With n1 one-byte and n2 two-byte symbols in the sequence (every nonterminal of a
    // block of <= 16 KiB has a two-byte id) the payload is 6 + rules + n1 + 2 n2 + what the counts need, and a round that replaces
    // r occurrences of (a, b) changes it by -(r (ca + cb - 2) - (ca + cb)): nothing is gained on (1,1) pairs, r - 3 on mixed ones,
    // 2r - 4 on (2,2) pairs.  The largest pair count f never rises (new adjacencies contain the new symbol, which occurs r <= f
    // times), so every later round has r <= f; one-byte symbols are never created (x replacements of (1,1) pairs and y of mixed
    // pairs use 2x + y <= n1 of them) and z replacements of (2,2) pairs need z <= n2 + x two-byte symbols; each round pays for its
    // rule.  Maximising y (1 - 3/f) + z (2 - 4/f) - x 2/f under these constraints gives at most n1 (1 - 3/f) + n2 (2 - 4/f), i.e.
    //   payload >= 6 + rules + (3 n1 + 4 n2) / f   for f >= 3   (f = 2: 6 + rules + n1 + 2 n2)
    // which ends blocks of high bytes (sine waves, noise) long before the pair bound does.
    const i64 lim = limit ? limit[b] : (i64)0x7fffffffffffffffll;
    u32 rule_bytes = 0;
    bool stopped = false;
    constexpr u32 IPT = MAXLEN / THREADS;        // 8 (16 on the largest shape) consecutive positions per thread
    // The pair table is emptied by the sweep that reads it (one pass and one barrier less per round than clearing it up front).
    for (u32 i = tid; i < HASH; i += THREADS) { S.hkey[i] = REPAIR_EMPTY; if (i < HASH / 2) S.hcnt2[i] = 0; }
    __syncthreads();
    high = __reduce_add_sync(0xffffffffu, high);
    if ((tid & 31) == 0 && high) atomicAdd(&S.nhigh, high);
    __syncthreads();
    u32 n2 = S.nhigh, n1 = bi.len - n2;                      // one- and two-byte symbols of the sequence (kept by every thread)
    u16* const q = S.seq;
    for (;;) {
        const u32 m = S.m;
        if (m < 2) break;
        // ---- pair histogram (every reader of last round's S.best is past a barrier; the barrier below publishes the reset)
        if (tid == 0) { S.best = 0; S.wsum = 0; }
        for (u32 i = tid; i + 1 < m; i += THREADS) {
            u32 key = ((u32)q[i] << 16) | q[i + 1];
            u32 h = rp_hash<HASH>(key);
            for (;;) {
                u32 old = atomicCAS(&S.hkey[h], REPAIR_EMPTY, key);
                if (old == REPAIR_EMPTY || old == key) { atomicAdd(&S.hcnt2[h >> 1], (h & 1u) ? 0x10000u : 1u); break; }
                h = (h + 1) & (HASH - 1);
            }
        }
        __syncthreads();
        // ---- best pair = max of (count << 32 | ~pair): per thread, per warp, one shared atomic per warp (one per qualifying slot
        //      serialised thousands of 64-bit atomics on one address per round); the slots are emptied on the way
        u32 wloc = 0;                                        // ULEB bytes of the left symbols of my distinct pairs
        unsigned long long lbest = 0;
        for (u32 j = tid; j < HASH / 2; j += THREADS) {      // slots 2j and 2j + 1: one 8-byte key load, one count word
            const uint2 kk = reinterpret_cast<const uint2*>(S.hkey)[j];
            if ((kk.x & kk.y) == REPAIR_EMPTY) continue;
            const u32 cc = S.hcnt2[j];
            reinterpret_cast<uint2*>(S.hkey)[j] = make_uint2(REPAIR_EMPTY, REPAIR_EMPTY); S.hcnt2[j] = 0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const u32 k = hh ? kk.y : kk.x, cnt = hh ? cc >> 16 : cc & 0xffffu;
                if (k == REPAIR_EMPTY) continue;
                wloc += 0x10001u + ((k >> 16) >= 128u ? 1u : 0u);      // distinct pairs in the upper half, their left symbols' bytes below
                const unsigned long long v = ((unsigned long long)cnt << 32) | (unsigned long long)(~k);
                if (cnt >= 2 && v > lbest) lbest = v;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, lbest, o); if (t > lbest) lbest = t; }
        if ((tid & 31) == 0 && lbest) atomicMax(&S.best, lbest);
        if (limit) {
            wloc = __reduce_add_sync(0xffffffffu, wloc);
            if ((tid & 31) == 0 && wloc) atomicAdd(&S.wsum, wloc);
        }
        __syncthreads();
        const unsigned long long best = S.best;
        if ((u32)(best >> 32) < 2) break;                    // V22.py:1875-1876
        if (limit) {
            const i64 f = (i64)(best >> 32), gain = (i64)n1 * (f > 3 ? f - 3 : 0) + (i64)n2 * (2 * f - 4);
            // pair bound, sharpened by the rules still needed: D distinct pairs need D boundaries, K later rules and a final
            // sequence of F symbols have K + F - 1, the rules shorten the sequence by at most f each (F >= m - f K), every boundary
            // costs at least a byte on its left and a rule one more on its right: payload >= 7 + rules + W + (m - 1 - D) / (f - 1)
            const u32 W = S.wsum & 0xffffu, D = S.wsum >> 16;
            const i64 more = (i64)(m - 1u - D + (u32)f - 2u) / (f - 1);
            if ((i64)(7u + rule_bytes + W) + more >= lim || f * (i64)(6u + rule_bytes + n1 + 2u * n2) - gain >= f * lim) { stopped = true; break; }
        }
        const u32 bkey = ~(u32)best;
        const u32 newsym = 256 + nrules;
        // ---- which occurrences are replaced: position i is "flagged" if pair(i) == best
        //      taken(i) = flagged(i) and (i - start of its maximal flagged run) is even
        u32 flags = 0, lastun = 0;                           // lastun: 1 + index of the last unflagged position among mine
        u32 mine[IPT + 1];                                   // my IPT symbols and the next thread's first: the sequence is rewritten in place below
#pragma unroll
        for (u32 k = 0; k <= IPT; ++k) { const u32 i = tid * IPT + k; mine[k] = i < m ? (u32)q[i] : 0xffffu; }
#pragma unroll
        for (u32 k = 0; k < IPT; ++k) {
            u32 i = tid * IPT + k;
            bool f = (i + 1 < m) && (((mine[k] << 16) | mine[k + 1]) == bkey);
            if (f) flags |= 1u << k; else lastun = i + 1;
        }
        u32 taken = 0, ntaken = 0;
        if ((bkey >> 16) != (bkey & 0xffffu)) {              // a != b: two occurrences cannot overlap, every one is taken (no run bookkeeping, four barriers less)
            taken = flags; ntaken = (u32)__popc(flags);
        } else {
            u32 incl = rp_scan<true, THREADS>(lastun, S.scan, nullptr);
            u32 prevun = __shfl_up_sync(0xffffffffu, incl, 1);   // exclusive: last unflagged before my first item
            if ((tid & 31) == 0) prevun = 0;
            // cross-warp exclusive: recompute from the scan array is gone; do it with a second tiny scan on warp leaders
            u32* s_wlast = S.wlast;
            if ((tid & 31) == 31) s_wlast[tid >> 5] = incl;
            __syncthreads();
            if ((tid & 31) == 0 && tid) prevun = s_wlast[(tid >> 5) - 1];
            __syncthreads();
            u32 run0 = prevun;                               // run0 = first index of the current flagged run
#pragma unroll
            for (u32 k = 0; k < IPT; ++k) {
                u32 i = tid * IPT + k;
                if ((flags >> k) & 1u) { if (((i - run0) & 1u) == 0) { taken |= 1u << k; ++ntaken; } }
                else run0 = i + 1;
            }
        }
        // removed(i) = taken(i-1): the second symbol of a replaced pair disappears
        u32 tprev = __shfl_up_sync(0xffffffffu, taken >> (IPT - 1), 1) & 1u;
        u32* s_wt = S.wt;
        if ((tid & 31) == 31) s_wt[tid >> 5] = (taken >> (IPT - 1)) & 1u;
        __syncthreads();
        if ((tid & 31) == 0) tprev = tid ? s_wt[(tid >> 5) - 1] : 0;
        u32 removed = ((taken << 1) | tprev) & ((1u << IPT) - 1);
        u32 keep = 0;
#pragma unroll
        for (u32 k = 0; k < IPT; ++k) { u32 i = tid * IPT + k; if (i < m && !((removed >> k) & 1u)) ++keep; }
        u32 tot;                                             // one scan for both counts: taken in the upper half, kept below (<= MAXLEN < 2^16)
        const u32 pincl = rp_scan<false, THREADS>((ntaken << 16) | keep, S.scan, &tot);
        if ((tot >> 16) < 2) break;                          // V22.py:1880-1882: rule not recorded, sequence unchanged
        const u32 newm = tot & 0xffffu;
        u32 o = (pincl & 0xffffu) - keep;
#pragma unroll
        for (u32 k = 0; k < IPT; ++k) {                      // in place: every thread read its symbols before the scan's barriers
            u32 i = tid * IPT + k;
            if (i < m && !((removed >> k) & 1u)) q[o++] = ((taken >> k) & 1u) ? (u16)newsym : (u16)mine[k];
        }
        if (tid == 0) { rules[nrules] = bkey; S.m = newm; }
        {
            const u32 ca = rp_uleb_size(bkey >> 16), cb = rp_uleb_size(bkey & 0xffff), r = tot >> 16;
            rule_bytes += ca + cb;
            n1 -= r * ((ca == 1u) + (cb == 1u));
            n2 = n2 + r - r * ((ca == 2u) + (cb == 2u));
        }
        ++nrules;
        __syncthreads();
    }
    __syncthreads();
    if (stopped) {
        if (tid == 0) { bacc[(size_t)b * 64 + 32] = 0; bacc[(size_t)b * 64 + 34] = 1; err[b] = KOLM_OK; }
        return;
    }
    // ---- serialise: 'R','P', ULEB 256, ULEB nrules, rules, ULEB len, symbols   (V22.py:1889-1903)
    const u32 m = S.m;
    u8* dst = tmp + (size_t)bi.pbase * 4;
    u32& s_hdr = S.hdr;
    u32 rbytes = 0;
    for (u32 r = tid; r < nrules; r += THREADS) { u32 k = rules[r]; rbytes += rp_uleb_size(k >> 16) + rp_uleb_size(k & 0xffff); }
    u32 rtot;
    rp_scan<false, THREADS>(rbytes, S.scan, &rtot);
    // rules are few thousand at most: thread 0 writes header + rules sequentially, everyone writes symbols
    if (tid == 0) {
        u8* p = dst; *p++ = 'R'; *p++ = 'P'; p = rp_put_uleb(p, 256); p = rp_put_uleb(p, nrules);
        for (u32 r = 0; r < nrules; ++r) { u32 k = rules[r]; p = rp_put_uleb(p, k >> 16); p = rp_put_uleb(p, k & 0xffff); }
        p = rp_put_uleb(p, m);
        s_hdr = (u32)(p - dst);
    }
    u32 sb = 0;
#pragma unroll
    for (u32 k = 0; k < IPT; ++k) { u32 i = tid * IPT + k; if (i < m) sb += rp_uleb_size(q[i]); }
    u32 stot;
    u32 sincl = rp_scan<false, THREADS>(sb, S.scan, &stot);
    __syncthreads();
    u8* p = dst + s_hdr + (sincl - sb);
#pragma unroll
    for (u32 k = 0; k < IPT; ++k) { u32 i = tid * IPT + k; if (i < m) p = rp_put_uleb(p, q[i]); }
    if (tid == 0) { bacc[(size_t)b * 64 + 32] = (u64)s_hdr + stot; err[b] = KOLM_OK; }
}

__global__ void k_repair_gather(const u8* __restrict__ tmp, const BlockInfo* __restrict__ binfo, const u64* __restrict__ bacc, u8* __restrict__ out,
                                const i64* __restrict__ cap_total, u64 cap) {
    if ((u64)*cap_total > cap) return;                      // exact total known before any byte is emitted: never write past the caller's buffer
    const u32 b = blockIdx.x;
    const u8* src = tmp + (size_t)binfo[b].pbase * 4;
    u8* dst = out + bacc[(size_t)b * 64 + 33];
    u64 n = bacc[(size_t)b * 64 + 32];
    for (u64 i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

int kolm_repair_big_impl(kolm_ctx* c, const u8* in, u8* tmp, cudaStream_t s);

// decode (v0): one thread per block; iterative expansion with an explicit stack in global scratch
__global__ void k_repair_dec(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo, u8* __restrict__ out,
                             u32* __restrict__ scratch_rules, u32* __restrict__ scratch_stack, int* __restrict__ err, int nblocks, const int* __restrict__ todo) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    if (todo && !todo[b]) return;                            // the parallel decoder (k_repair_dec2) already produced this block
    BlockInfo bi = binfo[b];
    const u8* d = pay + pay_off[b];
    i64 n = pay_off[b + 1] - pay_off[b], p = 2;
    u32* rules = scratch_rules + bi.pbase;       // capacity len (+pad) words = len/2 rules of two 32-bit symbols (what the encoder can emit)
    u32* stack = scratch_stack + bi.pbase;
    u32 cap = ((bi.len + KOLM_PAD - 1) / KOLM_PAD) * KOLM_PAD;
    if (cap < KOLM_PAD) cap = KOLM_PAD;
    u8* dst = out + bi.ioff;
    int e = KOLM_OK;
    auto get = [&](u64& v) -> bool { v = 0; int sh = 0; for (;;) { if (p >= n) return false; u8 x = d[p++]; if (sh < 64) v |= (u64)(x & 0x7F) << sh; if (!(x & 0x80)) return true; sh += 7; } };
    u64 term = 0, nr = 0, sl = 0;
    if (n < 2 || d[0] != 'R' || d[1] != 'P') e = KOLM_E_CORRUPT;
    if (!e && !get(term)) e = KOLM_E_TRUNCATED;
    if (!e && term != 256) e = KOLM_E_CORRUPT;
    if (!e && !get(nr)) e = KOLM_E_TRUNCATED;
    if (!e && 2 * nr > cap) e = KOLM_E_CORRUPT;               // more rules than any block of this length needs (the encoder emits <= len/2): cannot expand to len bytes with every rule used
    for (u64 r = 0; r < nr && !e; ++r) {
        u64 x, y;
        if (!get(x) || !get(y)) { e = KOLM_E_TRUNCATED; break; }
        if (x >= 256 + r || y >= 256 + r) { e = KOLM_E_CORRUPT; break; }
        rules[2 * r] = (u32)x; rules[2 * r + 1] = (u32)y;
    }
    if (!e && !get(sl)) e = KOLM_E_TRUNCATED;
    u32 o = 0;
    for (u64 t = 0; t < sl && !e; ++t) {
        u64 s;
        if (!get(s)) { e = KOLM_E_TRUNCATED; break; }
        if (s >= 256 + nr) { e = KOLM_E_CORRUPT; break; }
        u32 sp = 0; stack[sp++] = (u32)s;
        while (sp) {
            u32 x = stack[--sp];
            if (x < 256) { if (o >= bi.len) { e = KOLM_E_CORRUPT; break; } dst[o++] = (u8)x; }
            else { if (sp + 2 > cap) { e = KOLM_E_CORRUPT; break; } stack[sp++] = rules[2 * (x - 256) + 1]; stack[sp++] = rules[2 * (x - 256)]; }
        }
    }
    if (!e && o != bi.len) e = KOLM_E_CORRUPT;
    err[b] = e;
}

// decode, parallel (one CTA per block).  The payload after 'RP' is one stream of ULEB128 values (256, nrules, 2*nrules rule
// symbols, the sequence length, the sequence), and a byte ends a value iff its high bit is clear — so the values are found with a
// prefix count over the bytes, the rules' expansion lengths by relaxation in creation order (a rule only names earlier rules), the
// sequence symbols' output offsets by a scan, and the output is produced in chunks of 128 bytes: binary search for the sequence
// symbol that covers the chunk start, a length-guided descent to the exact byte, then an in-order walk with a small stack.
// Anything unusual (malformed stream, grammar deeper than the stack, length mismatch) leaves the block to the serial decoder
// above, which also owns the error codes: todo[b] = 1.
#define RD2_THREADS 256
#define RD2_STACK 96
__global__ void __launch_bounds__(RD2_THREADS) k_repair_dec2(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo,
                                                            u8* __restrict__ out, u32* __restrict__ g_ra, u32* __restrict__ g_rb, u32* __restrict__ g_rl,
                                                            u32* __restrict__ g_sq, u32* __restrict__ g_so, int* __restrict__ err, int* __restrict__ todo) {
    __shared__ u32 s_scan[RD2_THREADS / 32];
    __shared__ u32 s_R, s_start, s_seqlen, s_bad;
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5, b = blockIdx.x;
    const BlockInfo bi = binfo[b];
    const u32 L = bi.len;
    const u8* d = pay + pay_off[b];
    const u32 n = (u32)(pay_off[b + 1] - pay_off[b]);
    const u32 cap = max(((L + KOLM_PAD - 1) / KOLM_PAD) * KOLM_PAD, (u32)KOLM_PAD);
    u32* ra = g_ra + bi.pbase; u32* rb = g_rb + bi.pbase; u32* rl = g_rl + bi.pbase; u32* sq = g_sq + bi.pbase; u32* so = g_so + bi.pbase;
    auto exscan = [&](u32 v, u32* total) -> u32 {
        u32 x = v;
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += t; }
        if (lane == 31) s_scan[w] = x;
        __syncthreads();
        u32 pre = 0, tot = 0;
        for (int i = 0; i < RD2_THREADS / 32; ++i) { u32 t = s_scan[i]; if ((u32)i < w) pre += t; tot += t; }
        __syncthreads();
        *total = tot;
        return pre + x - v;
    };
    if (tid == 0) {
        s_bad = 0; s_R = 0; s_start = 0; s_seqlen = 0xffffffffu;
        u32 p = 2; u64 v0 = 0, v1 = 0; bool ok = L > 0 && n >= 4 && d[0] == 'R' && d[1] == 'P';
        auto get = [&](u64& v) { v = 0; int sh = 0; for (;;) { if (p >= n) return false; u8 x = d[p++]; if (sh < 35) v |= (u64)(x & 0x7F) << sh; else if (x & 0x7F) return false; if (!(x & 0x80)) return true; sh += 7; } };
        ok = ok && get(v0) && v0 == 256 && get(v1) && 2 * v1 <= cap;
        if (!ok) s_bad = 1; else { s_R = (u32)v1; s_start = p; }
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) todo[b] = 1; return; }
    const u32 R = s_R, start = s_start;
    // ---- values
    u32 nvals = 0;
    for (u32 base = start; base < n; base += RD2_THREADS * 16) {
        const u32 i0 = base + tid * 16;
        u32 mask = 0;
#pragma unroll
        for (u32 k = 0; k < 16; ++k) if (i0 + k < n && !(d[i0 + k] & 0x80)) mask |= 1u << k;
        u32 tot;
        u32 j = nvals + exscan(__popc(mask), &tot);
        while (mask) {
            const u32 k = __ffs(mask) - 1; mask &= mask - 1;
            const u32 e = i0 + k;                              // last byte of the value; its first byte follows the previous terminator
            u32 f = e;
            while (f > start && (d[f - 1] & 0x80) && e - f < 5) --f;
            u64 v = 0;
            if (e - f >= 5) v = ~0ull;                         // longer than any 32-bit value
            else for (u32 x = f; x <= e; ++x) v |= (u64)(d[x] & 0x7F) << (7 * (x - f));
            if (j < 2 * R) { const u32 r = j >> 1; if (v >= 256ull + r) s_bad = 1; else if (j & 1) rb[r] = (u32)v; else ra[r] = (u32)v; }
            else if (j == 2 * R) { if (v > L) s_bad = 1; else s_seqlen = (u32)v; }
            else { const u32 q = j - 2 * R - 1; if (q < cap) { if (v >= 256ull + R) { if (q < s_seqlen || s_seqlen == 0xffffffffu) sq[q] = 0xffffffffu; } else sq[q] = (u32)v; } }
            ++j;
        }
        nvals += tot;
    }
    __syncthreads();
    const u32 seqlen = s_seqlen;
    if (s_bad || seqlen == 0xffffffffu || nvals < 2 * R + 1 + seqlen) { if (tid == 0) todo[b] = 1; return; }
    // ---- expansion length of every rule (0 = not known yet): chunks in creation order, relaxation inside a chunk
    for (u32 base = 0; base < R; base += RD2_THREADS) {
        const u32 r = base + tid;
        if (r < R) rl[r] = 0;
        __syncthreads();
        for (u32 it = 0; it <= RD2_THREADS; ++it) {
            int pending = 0;
            if (r < R && rl[r] == 0) {
                const u32 a = ra[r], bb = rb[r];
                const u32 la = a < 256 ? 1u : rl[a - 256], lb = bb < 256 ? 1u : rl[bb - 256];
                if (la && lb) { const u64 t = (u64)la + lb; if (t > L) s_bad = 1; rl[r] = t > L ? 1u : (u32)t; } else pending = 1;
            }
            if (!__syncthreads_or(pending)) break;
        }
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) todo[b] = 1; return; }
    // ---- output offset of every sequence symbol
    u32 total = 0;
    for (u32 base = 0; base < seqlen; base += RD2_THREADS) {
        const u32 i = base + tid;
        u32 len = 0;
        if (i < seqlen) { const u32 v = sq[i]; if (v == 0xffffffffu) s_bad = 1; else len = v < 256 ? 1u : rl[v - 256]; }
        u32 tot; const u32 ex = exscan(len, &tot);
        if (i < seqlen) so[i] = total + ex;
        if ((u64)total + tot > L) s_bad = 1;
        total += tot;
    }
    __syncthreads();
    if (s_bad || total != L) { if (tid == 0) todo[b] = 1; return; }
    // ---- expand, 128 output bytes per thread and step
    u8* dst = out + bi.ioff;
    for (u32 c0 = tid * 128; c0 < L; c0 += RD2_THREADS * 128) {
        const u32 cend = min(L, c0 + 128);
        u32 lo = 0, hi = seqlen;                               // last symbol whose offset is <= c0
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (so[mid] <= c0) lo = mid; else hi = mid; }
        u32 i = lo, k = c0 - so[lo], o = c0, sp = 0;
        u32 stack[RD2_STACK];
        u32 cur = sq[i];
        bool over = false;
        while (cur >= 256) {                                    // guided descent to byte k of the symbol
            const u32 a = ra[cur - 256], bb = rb[cur - 256];
            const u32 la = a < 256 ? 1u : rl[a - 256];
            if (k < la) { if (sp >= RD2_STACK) { over = true; break; } stack[sp++] = bb; cur = a; } else { k -= la; cur = bb; }
        }
        while (!over) {
            dst[o++] = (u8)cur;
            if (o >= cend) break;
            if (sp) cur = stack[--sp]; else cur = sq[++i];
            while (cur >= 256) { if (sp >= RD2_STACK) { over = true; break; } stack[sp++] = rb[cur - 256]; cur = ra[cur - 256]; }
        }
        if (over) s_bad = 1;
    }
    __syncthreads();
    if (tid == 0) { if (s_bad) todo[b] = 1; else err[b] = KOLM_OK; }
}

// out_off == nullptr: device mode (kolm_encode_blocks) — the offsets stay in c->d_poff, nothing comes home
// limit (device, one entry per block, or nullptr): see k_repair_enc — blocks whose lower bound reaches limit[b] stop early, get
// size 0 here and a 1 in accumulator slot 34; only the shared-memory kernel looks at it (the incremental one always finishes).
int kolm_repair_enc_impl(kolm_ctx* c, const u8* in, u8* out, size_t out_cap, i64* out_off, cudaStream_t s, const i64* limit = nullptr) {
    const int nb = c->nblocks;
    if (!nb) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    static long long big_max = -1;                            // KOLM_REPAIR_BIG_MAX: largest block (bytes) the incremental kernel takes (0: none)
    if (big_max < 0) { const char* e = getenv("KOLM_REPAIR_BIG_MAX"); big_max = e ? atoll(e) : (1ll << 30); }
    if (c->max_len > REPAIR_XL && (long long)c->max_len > big_max) return KOLM_E_UNSUPPORTED;
    typedef RepairSmemT<REPAIR_SMALL, 256, 4096> RepairSmemSmall;
    typedef RepairSmemT<REPAIR_MID, 512, 8192> RepairSmemMid;
    typedef RepairSmemT<REPAIR_XL, REPAIR_THREADS, REPAIR_XL_HASH> RepairSmemXL;
    static bool attr_set[64];
    if (c->device < 64 && !attr_set[c->device]) {
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RepairSmem)));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH, REPAIR_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RepairSmem)));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_MID, 512, 8192, REPAIR_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RepairSmemMid)));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_SMALL, 256, 4096, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RepairSmemSmall)));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_XL, REPAIR_THREADS, REPAIR_XL_HASH, REPAIR_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RepairSmemXL)));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_XL, REPAIR_THREADS, REPAIR_XL_HASH, REPAIR_MAX>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        // all of the SM's shared memory for these kernels: two CTAs of the 115 KB shape only fit the largest carve-out
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH, REPAIR_MID>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_MID, 512, 8192, REPAIR_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(k_repair_enc<REPAIR_SMALL, 256, 4096, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set[c->device] = true;
    }
    static int small_on = -1;
    if (small_on < 0) { const char* e = getenv("KOLM_REPAIR_SMALL"); small_on = e ? atoi(e) : 1; }
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * 64 * 8, s));
    u8* tmp = (u8*)c->d_k0;                                   // 4 bytes per padded element
    if (small_on) {
        // blocks of up to 2 KiB on the small shape (five CTAs per SM), up to 4 KiB on the middle one (two), the others (if any) on the 160 KB shape
        KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_enc<REPAIR_SMALL, 256, 4096, 0><<<nb, 256, sizeof(RepairSmemSmall), s>>>(in, c->d_binfo, tmp, c->d_v0, c->d_bacc, c->d_err, limit));
        if (c->max_len > REPAIR_SMALL) KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_enc<REPAIR_MID, 512, 8192, REPAIR_SMALL><<<nb, 512, sizeof(RepairSmemMid), s>>>(in, c->d_binfo, tmp, c->d_v0, c->d_bacc, c->d_err, limit));
        if (c->max_len > REPAIR_MID) KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_enc<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH, REPAIR_MID><<<nb, REPAIR_THREADS, sizeof(RepairSmem), s>>>(in, c->d_binfo, tmp, c->d_v0, c->d_bacc, c->d_err, limit));
    } else
    KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_enc<REPAIR_MAX, REPAIR_THREADS, REPAIR_HASH, 0><<<nb, REPAIR_THREADS, sizeof(RepairSmem), s>>>(in, c->d_binfo, tmp, c->d_v0, c->d_bacc, c->d_err, limit));
    if (c->max_len > REPAIR_MAX) KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_enc<REPAIR_XL, REPAIR_THREADS, REPAIR_XL_HASH, REPAIR_MAX><<<nb, REPAIR_THREADS, sizeof(RepairSmemXL), s>>>(in, c->d_binfo, tmp, c->d_v0, c->d_bacc, c->d_err, limit));
    if (c->max_len > REPAIR_XL) KOLM_TRY(kolm_repair_big_impl(c, in, tmp, s));    // blocks the shared-memory kernel passed over
    KL(c, KC_RICE_PLAN, (i64)nb * 16, s, k_lz_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_poff, nb));
    KL(c, KC_MISC, c->total_bytes, s, k_repair_gather<<<nb, 256, 0, s>>>(tmp, c->d_binfo, c->d_bacc, out, c->d_poff + nb, (u64)out_cap));
    CUDA_TRY(cudaGetLastError());
    if (!out_off) return KOLM_OK;
    CUDA_TRY(cudaMemcpyAsync(c->h_poff, c->d_poff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_off, c->h_poff, (size_t)(nb + 1) * 8);
    if ((size_t)out_off[nb] > out_cap) return KOLM_E_CAPACITY;
    return KOLM_OK;
}

int kolm_repair_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, u8* out, cudaStream_t s) {
    const int nb = c->nblocks;
    if (!nb) return KOLM_OK;
    memcpy(c->h_poff, pay_off, (size_t)(nb + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    int* todo = (int*)c->d_done;                             // blocks the parallel decoder leaves to the serial one
    CUDA_TRY(cudaMemsetAsync(todo, 0, (size_t)nb * 4, s));
    KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_dec2<<<nb, RD2_THREADS, 0, s>>>(pay, c->d_poff, c->d_binfo, out, c->d_k0, c->d_k1, c->d_v0, c->d_v1, c->d_sa, c->d_err, todo));
    KL(c, KC_MISC, c->total_bytes * 2, s, k_repair_dec<<<(nb + 63) / 64, 64, 0, s>>>(pay, c->d_poff, c->d_binfo, out, c->d_k0, c->d_v0, c->d_err, nb, todo));
    CUDA_TRY(cudaGetLastError());
    return rice_dec_finish(c, s);
}
