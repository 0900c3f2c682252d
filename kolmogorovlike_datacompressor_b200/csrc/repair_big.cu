// repair_big.cu — Re-Pair grammar candidate for blocks that do not fit the shared-memory kernel.
//
//   repair_compress   kolm_final_researched_v2-2.py:1841-1911  (_count_pairs :1817, _replace_non_overlapping :1824)
//
// Same semantics as repair.cu (most frequent adjacent pair over ALL overlapping occurrences, ties -> the smallest pair, at
// least two occurrences and at least two replacements, left-to-right non-overlapping replacement), but incremental: the
// reference recounts every pair in every round — O(rounds * n), days for a 1 MiB block in Python — while the outcome only
// depends on the counts, which change in O(1) places per replacement.
//
// One CTA per block, everything in a per-CTA slab of global memory:
//   sequence      doubly linked list over the original positions (sym / prv / nxt); positions never move, so the index order
//                 is the sequence order;
//   pair table    open addressing, key = a << 32 | b, exact count, base of an occurrence ARRAY (the positions where the pair
//                 started when it was created: every occurrence of a pair is created in ONE round — the one that creates its
//                 newer symbol, or the initial count — so the array is sized from that round's count and filled behind a
//                 cursor; pairs created with a single occurrence can never be chosen and get none; stale entries are skipped
//                 when the array is read — by all threads at once —, and a pair is processed at most once because a replaced
//                 pair can never re-form);
//   priority      a list M sorted by (count desc, pair asc) walked by a cursor + an unsorted pending list P of the pairs whose
//                 count changed since the last merge (only decrements of old pairs and new pairs containing the newest symbol,
//                 so counts never rise after a pair's creation round and an entry is valid iff its count is still current);
//                 the round's pair = the better of M's first valid entry and P's best valid entry; P is sorted and merged into
//                 the unread rest of M when it reaches RPB_PMAX entries.
//   a round       occurrence list -> valid occurrences; for a == b the runs are walked from their starts (every other
//                 occurrence is taken); neighbours' counts are decremented from the OLD links, the list is relinked, the new
//                 pairs are counted from the NEW links (each adjacent pair exactly once: left pair always, right pair unless
//                 the right neighbour is itself a replaced occurrence), touched pairs go to P.
// Latency bound pointer chasing: no roofline target.  Throughput is set by the number of rounds (~10^5 for 1 MiB of
// text) times a few microseconds of dependent global-memory steps.
#include "common.cuh"

#ifndef RPB_CTAS_PER_SM
#define RPB_CTAS_PER_SM 4
#endif
#ifndef RPB_THREADS
#define RPB_THREADS 128
#endif
#define RPB_PMAX 8192u
#ifndef RPB_V
#define RPB_V 30            // experiment mask: 1 parallel M skip, 2 warp pbest update, 4 unrolled rescan, 8 small hash, 16 register-resident small rounds
#endif
#define RPB_NIL 0xffffffffu
#define RPB_EMPTY 0xffffffffffffffffull

struct RpbSlab {
    u32 *sym, *prv, *nxt, *stamp;                       // [n]
    u64* hkey; u32 *hcnt, *hhead, *hstamp; u32 hmask;   // [H]
    u32* opos; u32 ocap;                                // occurrence arrays: [base] = entries filled, [base+1 ...] = positions
    u64 *mpair, *tpair; u32 *mcnt, *tcnt, *mslot, *tslot; u32 mcap;   // sorted list + merge target
    u64* ppair; u32 *pcnt, *pslot; u32 pcap;            // pending
    u32 *occ, *tk, *touched; u32 tcap;                  // per-round lists
    u64* rules;                                          // [n/2 + 1]
};

// Capacities.  Pairs whose count is >= 2 are distinct and share the sequence's < n adjacent positions, so fewer than n/2 exist at
// any time: that bounds the valid entries of the sorted list, of a merge, and what one round can append to the pending list
// (appended entries are distinct pairs with a current count >= 2); the pending list is merged once it holds RPB_PMAX entries.
__host__ __device__ inline u32 rpb_pow2_at_least(u64 v) { u32 p = 1024; while ((u64)p < v) p <<= 1; return p; }
__host__ __device__ inline size_t rpb_align(size_t x) { return (x + 255) & ~(size_t)255; }
// slab bytes for a block of n bytes (n >= 2)
__host__ __device__ inline size_t rpb_need(u64 n) {
    const u64 H = (RPB_V & 8) ? rpb_pow2_at_least(3 * n + 65536) : rpb_pow2_at_least(4 * n + 131072), oc = 9 * n / 2 + 64, mc = n / 2 + 64, pc = n / 2 + RPB_PMAX + 64, tc = 2 * n + 64;
    size_t b = 0;
    b += 4 * rpb_align(n * 4);
    b += rpb_align(H * 8) + 3 * rpb_align(H * 4);
    b += rpb_align(oc * 4);
    b += 2 * rpb_align(mc * 8) + 4 * rpb_align(mc * 4);
    b += rpb_align(pc * 8) + 2 * rpb_align(pc * 4);
    b += 2 * rpb_align(n * 4) + rpb_align(tc * 4);
    b += rpb_align((n / 2 + 2) * 8);
    return b + 1024;
}
__device__ inline void rpb_carve(u8* base, u64 n, RpbSlab& S) {
    const u64 H = (RPB_V & 8) ? rpb_pow2_at_least(3 * n + 65536) : rpb_pow2_at_least(4 * n + 131072), oc = 9 * n / 2 + 64, mc = n / 2 + 64, pc = n / 2 + RPB_PMAX + 64, tc = 2 * n + 64;
    u8* p = base;
    auto take = [&](size_t bytes) { u8* r = p; p += rpb_align(bytes); return r; };
    S.sym = (u32*)take(n * 4); S.prv = (u32*)take(n * 4); S.nxt = (u32*)take(n * 4); S.stamp = (u32*)take(n * 4);
    S.hkey = (u64*)take(H * 8); S.hcnt = (u32*)take(H * 4); S.hhead = (u32*)take(H * 4); S.hstamp = (u32*)take(H * 4); S.hmask = (u32)H - 1;
    S.opos = (u32*)take(oc * 4); S.ocap = (u32)oc;
    S.mpair = (u64*)take(mc * 8); S.tpair = (u64*)take(mc * 8);
    S.mcnt = (u32*)take(mc * 4); S.tcnt = (u32*)take(mc * 4); S.mslot = (u32*)take(mc * 4); S.tslot = (u32*)take(mc * 4); S.mcap = (u32)mc;
    S.ppair = (u64*)take(pc * 8); S.pcnt = (u32*)take(pc * 4); S.pslot = (u32*)take(pc * 4); S.pcap = (u32)pc;
    S.occ = (u32*)take(n * 4); S.tk = (u32*)take(n * 4); S.touched = (u32*)take(tc * 4); S.tcap = (u32)tc;
    S.rules = (u64*)take((n / 2 + 2) * 8);
}

__device__ __forceinline__ u32 rpb_hash(u64 k) { k ^= k >> 29; k *= 0x9E3779B97F4A7C15ull; k ^= k >> 32; return (u32)k; }
// slot of `key`, inserting it (count 0, empty occurrence list) if absent; safe under concurrent callers
__device__ __forceinline__ u32 rpb_slot(const RpbSlab& S, u64 key) {
    u32 h = rpb_hash(key) & S.hmask;
    for (;;) {
        u64 cur = S.hkey[h];
        if (cur == key) return h;
        if (cur == RPB_EMPTY) {
            u64 old = atomicCAS((unsigned long long*)&S.hkey[h], (unsigned long long)RPB_EMPTY, (unsigned long long)key);
            if (old == RPB_EMPTY || old == key) return h;
        }
        h = (h + 1) & S.hmask;
    }
}
__device__ __forceinline__ bool rpb_better(u32 ca, u64 pa, u32 cb, u64 pb) { return ca > cb || (ca == cb && pa < pb); }

struct RpbShared {
    u32 scan[RPB_THREADS / 32];
    u64 red_pair[RPB_THREADS / 32]; u32 red_cnt[RPB_THREADS / 32]; u32 red_idx[RPB_THREADS / 32];
    u32 n_occ_nodes, n_occ, n_tk, n_touched, n_p, mpos, mlen, n_alloc;
    u32 pbest; int pbest_ok;                             // index of the best pending entry (valid when pbest_ok)
    u64 cur_pair; u32 cur_cnt, cur_slot, cur_base; int stop;
    u32 total;
};

// block-wide exclusive sum of one u32 per thread
__device__ __forceinline__ u32 rpb_exscan(u32 v, RpbShared& sh, u32* total) {
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 x = v;
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (u32)o) x += t; }
    if (lane == 31) sh.scan[w] = x;
    __syncthreads();
    u32 pre = 0, tot = 0;
    for (int i = 0; i < RPB_THREADS / 32; ++i) { u32 t = sh.scan[i]; if ((u32)i < w) pre += t; tot += t; }
    __syncthreads();
    *total = tot;
    return pre + x - v;
}

// stable filter of the entries [lo, hi) of (pair, cnt, slot) that are still current into dst (from index 0); returns the count
__device__ u32 rpb_compact(const RpbSlab& S, RpbShared& sh, const u64* sp, const u32* sc, const u32* ss, u32 lo, u32 hi, u64* dp, u32* dc, u32* ds) {
    u32 out = 0;
    for (u32 base = lo; base < hi; base += RPB_THREADS) {
        const u32 i = base + threadIdx.x;
        u32 ok = 0; u64 p = 0; u32 c = 0, s = 0;
        if (i < hi) { p = sp[i]; c = sc[i]; s = ss[i]; ok = (c >= 2 && S.hcnt[s] == c) ? 1u : 0u; }
        u32 tot;
        const u32 ex = rpb_exscan(ok, sh, &tot);
        if (ok) { dp[out + ex] = p; dc[out + ex] = c; ds[out + ex] = s; }
        out += tot;
    }
    __syncthreads();
    return out;
}

// bitonic sort of (pair, cnt, slot)[0, n) by (cnt desc, pair asc) for any n: the mirrored formulation, in which every
// compare-exchange puts the better entry at the lower index, so indices >= n act as (never stored) worst-possible sentinels
__device__ void rpb_sort(u64* P, u32* C, u32* L, u32 n, RpbShared& sh) {
    (void)sh;
    u32 N = 1; while (N < n) N <<= 1;
    auto cx = [&](u32 i, u32 l) {
        const u64 pi = P[i], pl = P[l]; const u32 ci = C[i], cl = C[l];
        if (rpb_better(cl, pl, ci, pi)) { P[i] = pl; P[l] = pi; C[i] = cl; C[l] = ci; const u32 t = L[i]; L[i] = L[l]; L[l] = t; }
    };
    for (u32 k = 2; k <= N; k <<= 1) {
        for (u32 i = threadIdx.x; i < n; i += RPB_THREADS) { const u32 l = i ^ (k - 1); if (l > i && l < n) cx(i, l); }
        __syncthreads();
        for (u32 j = k >> 2; j > 0; j >>= 1) {
            for (u32 i = threadIdx.x; i < n; i += RPB_THREADS) { const u32 l = i ^ j; if (l > i && l < n) cx(i, l); }
            __syncthreads();
        }
    }
    __syncthreads();
}

__device__ __forceinline__ u32 rpb_uleb_size(u32 v) { return v < 128u ? 1u : v < 16384u ? 2u : v < 2097152u ? 3u : v < 268435456u ? 4u : 5u; }
__device__ __forceinline__ u8* rpb_put_uleb(u8* p, u32 v) { while (v >= 128) { *p++ = (u8)(v | 0x80); v >>= 7; } *p++ = (u8)v; return p; }

// blist: the blocks this launch handles; slabs: one slab of slab_bytes per CTA; next: work counter
__global__ void __launch_bounds__(RPB_THREADS) k_repair_big(const u8* __restrict__ in, const BlockInfo* __restrict__ binfo, const u32* __restrict__ blist, int nlist,
                                                          u8* __restrict__ slabs, size_t slab_bytes, u32* __restrict__ next, u8* __restrict__ tmp,
                                                          u64* __restrict__ bacc, int* __restrict__ err) {
    __shared__ RpbShared sh;
    __shared__ u32 s_block;
    const u32 tid = threadIdx.x;
    for (;;) {
        if (tid == 0) s_block = atomicAdd(next, 1u);
        __syncthreads();
        const u32 li = s_block;
        __syncthreads();
        if (li >= (u32)nlist) return;
        const u32 b = blist[li];
        const BlockInfo bi = binfo[b];
        const u32 n = bi.len;
        const u8* src = in + bi.ioff;
        RpbSlab S;
        rpb_carve(slabs + (size_t)blockIdx.x * slab_bytes, n, S);
        const u32 H = S.hmask + 1;
        // ---- init
        for (u32 i = tid; i < n; i += RPB_THREADS) { S.sym[i] = src[i]; S.prv[i] = i ? i - 1 : RPB_NIL; S.nxt[i] = i + 1 < n ? i + 1 : RPB_NIL; S.stamp[i] = 0; }
        for (u32 i = tid; i < H; i += RPB_THREADS) { S.hkey[i] = RPB_EMPTY; S.hcnt[i] = 0; S.hhead[i] = RPB_NIL; S.hstamp[i] = 0; }
        if (tid == 0) { sh.n_occ_nodes = 0; sh.n_p = 0; sh.mpos = 0; sh.mlen = 0; sh.pbest_ok = 0; sh.stop = 0; }
        __syncthreads();
        for (u32 i = tid; i + 1 < n; i += RPB_THREADS) {
            const u64 key = ((u64)src[i] << 32) | src[i + 1];
            const u32 s = rpb_slot(S, key);
            atomicAdd(&S.hcnt[s], 1u);
            S.tk[i] = s;
        }
        __syncthreads();
        for (u32 i = tid; i < H; i += RPB_THREADS) {
            const u32 c = S.hcnt[i];
            if (c >= 2) {
                const u32 o = atomicAdd(&sh.mlen, 1u); S.mpair[o] = S.hkey[i]; S.mcnt[o] = c; S.mslot[o] = i;
                const u32 base = atomicAdd(&sh.n_occ_nodes, c + 1u); S.hhead[i] = base; S.opos[base] = 0;
            }
        }
        __syncthreads();
        for (u32 i = tid; i + 1 < n; i += RPB_THREADS) {
            const u32 base = S.hhead[S.tk[i]];
            if (base != RPB_NIL) S.opos[base + 1u + atomicAdd(&S.opos[base], 1u)] = i;
        }
        __syncthreads();
        rpb_sort(S.mpair, S.mcnt, S.mslot, sh.mlen, sh);
        u32 nrules = 0, round = 0;
        // ---- rounds
        for (;;) {
            ++round;
            // pending list: (re)find its best valid entry when the cached one is gone
            if (sh.n_p && !sh.pbest_ok) {
                u32 bc = 0, bidx = RPB_NIL; u64 bp = RPB_EMPTY;
                const u32 npend = sh.n_p;
                if (RPB_V & 4) {
#pragma unroll 4
                    for (u32 i = tid; i < npend; i += RPB_THREADS) {
                        const u32 c = S.pcnt[i], sl = S.pslot[i]; const u64 p = S.ppair[i];
                        const u32 hc = S.hcnt[sl];
                        if (c >= 2 && hc == c && rpb_better(c, p, bc, bp)) { bc = c; bp = p; bidx = i; }
                    }
                } else {
                    for (u32 i = tid; i < npend; i += RPB_THREADS) {
                        const u32 c = S.pcnt[i];
                        if (c >= 2 && S.hcnt[S.pslot[i]] == c && rpb_better(c, S.ppair[i], bc, bp)) { bc = c; bp = S.ppair[i]; bidx = i; }
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const u32 oc = __shfl_xor_sync(0xffffffffu, bc, o), oi = __shfl_xor_sync(0xffffffffu, bidx, o); const u64 op = __shfl_xor_sync(0xffffffffu, bp, o);
                    if (rpb_better(oc, op, bc, bp)) { bc = oc; bp = op; bidx = oi; }
                }
                if ((tid & 31) == 0) { sh.red_cnt[tid >> 5] = bc; sh.red_pair[tid >> 5] = bp; sh.red_idx[tid >> 5] = bidx; }
                __syncthreads();
                if (tid == 0) {
                    for (int i = 1; i < RPB_THREADS / 32; ++i) if (rpb_better(sh.red_cnt[i], sh.red_pair[i], bc, bp)) { bc = sh.red_cnt[i]; bp = sh.red_pair[i]; bidx = sh.red_idx[i]; }
                    sh.pbest = bidx; sh.pbest_ok = 1;         // RPB_NIL: nothing valid is pending
                }
                __syncthreads();
            }
            if (!(RPB_V & 1)) {
                if (tid == 0) { u32 mp = sh.mpos; while (mp < sh.mlen && !(S.mcnt[mp] >= 2 && S.hcnt[S.mslot[mp]] == S.mcnt[mp])) ++mp; sh.mpos = mp; }
            } else if (tid < 32) {
                u32 mp = sh.mpos; const u32 ml = sh.mlen;          // warp 0 skips the entries whose count moved on, 32 at a time
                __syncwarp();
                for (;;) {
                    const u32 i = mp + tid; bool ok = false;
                    if (i < ml) { const u32 c = S.mcnt[i]; ok = c >= 2 && S.hcnt[S.mslot[i]] == c; }
                    const u32 m = __ballot_sync(0xffffffffu, ok);
                    if (m) { mp += (u32)__ffs((int)m) - 1u; break; }
                    if (mp + 32u >= ml) { mp = ml; break; }
                    mp += 32u;
                }
                if (tid == 0) sh.mpos = mp;
                __syncwarp();
            }
            if (tid == 0) {
                const u32 mp = sh.mpos;
                u32 c = 0, s = 0; u64 p = RPB_EMPTY; int from_p = 0;
                const bool hasm = mp < sh.mlen, hasp = sh.n_p && sh.pbest != RPB_NIL;
                if (hasm) { c = S.mcnt[mp]; p = S.mpair[mp]; s = S.mslot[mp]; }
                u32 pc = 0, ps = 0; u64 pp = RPB_EMPTY;
                if (hasp) { const u32 i = sh.pbest; pc = S.pcnt[i]; pp = S.ppair[i]; ps = S.pslot[i]; }
                // both candidates' occurrence-array bases are fetched now (two loads in flight) so the round does not start with one
                const u32 hbm = hasm ? S.hhead[s] : RPB_NIL, hbp = hasp ? S.hhead[ps] : RPB_NIL;
                u32 hb = hbm;
                if (hasp && rpb_better(pc, pp, c, p)) { c = pc; p = pp; s = ps; hb = hbp; from_p = 1; }
                sh.cur_cnt = c; sh.cur_pair = p; sh.cur_slot = s; sh.cur_base = hb;
                if (c < 2) sh.stop = 1;
                else if (from_p) { S.pcnt[sh.pbest] = 0; sh.pbest_ok = 0; }   // consumed
                else sh.mpos = mp + 1;
                sh.n_occ = 0; sh.n_tk = 0; sh.n_touched = 0;
            }
            __syncthreads();
            if (sh.stop) break;
            const u64 cp = sh.cur_pair; const u32 A = (u32)(cp >> 32), B = (u32)cp, cslot = sh.cur_slot;
            const u32 newsym = 256 + nrules;
            const u32 st_seen = round * 4 + 1, st_take = round * 4 + 2, st_part = round * 4 + 3;
            const u32 p0 = sh.n_p;                               // pending entries before this round's (appended after the next barriers)
            // touched pairs are only listed here (a pair can be listed up to four times per replaced occurrence: 4 * ntk <= 2n
            // entries); the pass over the list below claims each pair once — no atomic with a return value inside this chain
            auto touch = [&](u32 s) { const u32 o = atomicAdd(&sh.n_touched, 1u); if (o < S.tcap) S.touched[o] = s; };
            const u32 obase = sh.cur_base;
            const u32 nh = obase == RPB_NIL ? 0u : S.opos[obase];
            const u32 pspec = (obase != RPB_NIL && tid < 2) ? S.opos[obase + 1u + tid] : 0u;   // every array holds >= 2 hints: no need to wait for nh
            const bool fast = (RPB_V & 16) && A != B && nh <= RPB_THREADS;
            u32 ntk;
            u32 f_p = RPB_NIL, f_xpos = RPB_NIL, f_sl = RPB_NIL, f_sr = RPB_NIL;      // fast rounds: my occurrence lives in registers
            u32* const lslot = S.occ; u32* const rslot = S.occ + (n >> 1);
            if (fast) {
                // ---- small round (one position hint per thread, a != b so occurrences cannot overlap): the occurrence, its
                //      neighbours and their symbols are read once, under the OLD links, and kept in registers
                u32 q = RPB_NIL, x = RPB_NIL, y = RPB_NIL, sx = 0, sy = 0; bool valid = false;
                if (tid < nh) {
                    const u32 p = tid < 2 ? pspec : S.opos[obase + 1u + tid];
                    const bool fresh = atomicExch(&S.stamp[p], st_seen) != st_seen;   // the same position can be hinted twice
                    const u32 sp = S.sym[p]; q = S.nxt[p]; x = S.prv[p];
                    if (sp == A && q != RPB_NIL) {
                        const u32 sq = S.sym[q]; y = S.nxt[q];
                        if (x != RPB_NIL) sx = S.sym[x];
                        if (sq == B && fresh) { valid = true; f_p = p; if (y != RPB_NIL) sy = S.sym[y]; }
                    }
                }
                ntk = (u32)__syncthreads_count(valid);
                if (ntk < 2) break;                              // V22.py:1880-1882: the rule is not recorded, the sequence stays
                if (valid) { S.stamp[f_p] = st_take; S.stamp[q] = st_part; }
                __syncthreads();
                if (valid) {
                    const u32 stx = x != RPB_NIL ? S.stamp[x] : 0u, sty = y != RPB_NIL ? S.stamp[y] : 0u;
                    const bool right = y != RPB_NIL && sty != st_take;
                    // old neighbours lose an occurrence; a != b: neither (sx, a) nor (b, sy) can be the round's pair
                    if (x != RPB_NIL) { const u32 s = rpb_slot(S, ((u64)sx << 32) | A); atomicSub(&S.hcnt[s], 1u); touch(s); }
                    if (right) { const u32 s = rpb_slot(S, ((u64)B << 32) | sy); atomicSub(&S.hcnt[s], 1u); touch(s); }
                    // new neighbours: a left neighbour that is the second half of a replaced occurrence stands for that occurrence
                    if (x != RPB_NIL) {
                        const bool part = stx == st_part;
                        f_xpos = part ? S.prv[x] : x;
                        f_sl = rpb_slot(S, ((u64)(part ? newsym : sx) << 32) | newsym); atomicAdd(&S.hcnt[f_sl], 1u); touch(f_sl);
                    }
                    if (right) { f_sr = rpb_slot(S, ((u64)newsym << 32) | sy); atomicAdd(&S.hcnt[f_sr], 1u); touch(f_sr); }
                    // relink: nobody reads these words in this phase (prv[x] above is only read for second halves, which no one writes)
                    S.sym[f_p] = newsym; S.sym[q] = RPB_NIL; S.nxt[f_p] = y;
                    if (y != RPB_NIL) S.prv[y] = f_p;
                }
            } else {
            // ---- valid occurrences: all threads read the pair's occurrence array (position hints), check the hints against the
            //      live sequence and claim each position once
            {
                for (u32 i = tid; i < nh; i += RPB_THREADS) {
                    const u32 p = S.opos[obase + 1u + i];
                    if (S.sym[p] != A) continue;
                    const u32 q = S.nxt[p];
                    if (q == RPB_NIL || S.sym[q] != B) continue;
                    if (atomicExch(&S.stamp[p], st_seen) == st_seen) continue;       // the same position can be hinted twice
                    S.occ[atomicAdd(&sh.n_occ, 1u)] = p;
                }
            }
            __syncthreads();
            const u32 nocc = sh.n_occ;
            // ---- which occurrences are replaced
            if (A != B) {
                for (u32 i = tid; i < nocc; i += RPB_THREADS) { const u32 p = S.occ[i]; S.tk[i] = p; }
                if (tid == 0) sh.n_tk = nocc;
            } else {
                for (u32 i = tid; i < nocc; i += RPB_THREADS) {      // run starts walk their run: every other occurrence from the start
                    u32 p = S.occ[i];
                    const u32 x = S.prv[p];
                    if (x != RPB_NIL && S.sym[x] == A) continue;
                    for (;;) {
                        const u32 q = S.nxt[p];
                        if (q == RPB_NIL || S.sym[q] != A) break;
                        S.tk[atomicAdd(&sh.n_tk, 1u)] = p;
                        p = S.nxt[q];
                        if (p == RPB_NIL || S.sym[p] != A) break;
                    }
                }
            }
            __syncthreads();
            ntk = sh.n_tk;
            if (ntk < 2) break;                                  // V22.py:1880-1882: the rule is not recorded, the sequence stays
            for (u32 i = tid; i < ntk; i += RPB_THREADS) { const u32 p = S.tk[i]; S.stamp[p] = st_take; S.stamp[S.nxt[p]] = st_part; }
            __syncthreads();
            // ---- old neighbours lose an occurrence (old links)
            for (u32 i = tid; i < ntk; i += RPB_THREADS) {
                const u32 p = S.tk[i], q = S.nxt[p], x = S.prv[p], y = S.nxt[q];
                if (x != RPB_NIL) { const u64 k = ((u64)S.sym[x] << 32) | A; if (k != cp) { const u32 s = rpb_slot(S, k); atomicSub(&S.hcnt[s], 1u); touch(s); } }
                if (y != RPB_NIL && S.stamp[y] != st_take) { const u64 k = ((u64)B << 32) | S.sym[y]; if (k != cp) { const u32 s = rpb_slot(S, k); atomicSub(&S.hcnt[s], 1u); touch(s); } }
            }
            __syncthreads();
            // ---- relink
            for (u32 i = tid; i < ntk; i += RPB_THREADS) {
                const u32 p = S.tk[i], q = S.nxt[p], y = S.nxt[q];
                S.sym[p] = newsym; S.sym[q] = RPB_NIL; S.nxt[p] = y;
                if (y != RPB_NIL) S.prv[y] = p;
            }
            __syncthreads();
            // ---- new neighbours (new links): left pair always, right pair unless the right neighbour was replaced too.  Counted
            //      now (slots remembered: occ is free again and ntk <= n/2), their occurrence arrays are filled once sized.
            for (u32 i = tid; i < ntk; i += RPB_THREADS) {
                const u32 p = S.tk[i], x = S.prv[p], y = S.nxt[p];
                u32 sl = RPB_NIL, sr = RPB_NIL;
                if (x != RPB_NIL) { sl = rpb_slot(S, ((u64)S.sym[x] << 32) | newsym); atomicAdd(&S.hcnt[sl], 1u); touch(sl); }
                if (y != RPB_NIL && S.stamp[y] != st_take) { sr = rpb_slot(S, ((u64)newsym << 32) | S.sym[y]); atomicAdd(&S.hcnt[sr], 1u); touch(sr); }
                lslot[i] = sl; rslot[i] = sr;
            }
            }
            if (tid == 0) { S.hcnt[cslot] = 0; S.rules[nrules] = cp; sh.n_alloc = 0; }
            ++nrules;
            __syncthreads();
            // ---- touched pairs whose count is still >= 2 become pending
            const u32 ntouch = min(sh.n_touched, S.tcap);
            for (u32 i = tid; i < ntouch; i += RPB_THREADS) {
                const u32 s = S.touched[i];
                const u32 seen = atomicExch(&S.hstamp[s], round);       // claim: a pair listed several times is handled once
                const u32 c = S.hcnt[s]; const u64 key = S.hkey[s];
                if (seen != round && c >= 2) {
                    const u32 o = atomicAdd(&sh.n_p, 1u); if (o < S.pcap) { S.ppair[o] = key; S.pcnt[o] = c; S.pslot[o] = s; }
                    if ((u32)(key >> 32) == newsym || (u32)key == newsym) {       // a pair born this round: all its occurrences exist now
                        const u32 base = atomicAdd(&sh.n_occ_nodes, c + 1u);
                        if (base + c + 1u <= S.ocap) { S.hhead[s] = base; S.opos[base] = 0; sh.n_alloc = 1; }
                    }
                }
            }
            __syncthreads();
            if (sh.n_alloc) {                                    // some pair born this round occurs twice or more: fill its array
                if (fast) {
                    if (f_sl != RPB_NIL) { const u32 base = S.hhead[f_sl]; if (base != RPB_NIL) S.opos[base + 1u + atomicAdd(&S.opos[base], 1u)] = f_xpos; }
                    if (f_sr != RPB_NIL) { const u32 base = S.hhead[f_sr]; if (base != RPB_NIL) S.opos[base + 1u + atomicAdd(&S.opos[base], 1u)] = f_p; }
                } else {
                    for (u32 i = tid; i < ntk; i += RPB_THREADS) {
                        const u32 sl = lslot[i], sr = rslot[i];
                        if (sl != RPB_NIL) { const u32 base = S.hhead[sl]; if (base != RPB_NIL) S.opos[base + 1u + atomicAdd(&S.opos[base], 1u)] = S.prv[S.tk[i]]; }
                        if (sr != RPB_NIL) { const u32 base = S.hhead[sr]; if (base != RPB_NIL) S.opos[base + 1u + atomicAdd(&S.opos[base], 1u)] = S.tk[i]; }
                    }
                }
            }
            if (!(RPB_V & 2)) {
                if (tid == 0) {
                    const u32 np = min(sh.n_p, S.pcap);
                    sh.n_p = np;
                    if (sh.pbest_ok) {
                        u32 bi2 = sh.pbest;
                        for (u32 i = p0; i < np; ++i) if (bi2 == RPB_NIL || rpb_better(S.pcnt[i], S.ppair[i], S.pcnt[bi2], S.ppair[bi2])) bi2 = i;
                        if (bi2 != RPB_NIL && !(S.pcnt[bi2] >= 2 && S.hcnt[S.pslot[bi2]] == S.pcnt[bi2])) sh.pbest_ok = 0; else sh.pbest = bi2;
                    }
                }
            } else if (tid < 32) {                               // warp 0: the new pending entries against the cached best
                const u32 np = min(sh.n_p, S.pcap);
                const int ok = sh.pbest_ok; const u32 cur = sh.pbest;
                __syncwarp();
                if (ok) {                                        // new entries can only improve the cached best
                    u32 bc = 0, bidx = RPB_NIL; u64 bp = RPB_EMPTY;
                    for (u32 i = p0 + tid; i < np; i += 32) { const u32 c = S.pcnt[i]; const u64 p = S.ppair[i]; if (rpb_better(c, p, bc, bp)) { bc = c; bp = p; bidx = i; } }
                    for (int o = 16; o > 0; o >>= 1) {
                        const u32 oc = __shfl_xor_sync(0xffffffffu, bc, o), oi = __shfl_xor_sync(0xffffffffu, bidx, o); const u64 op = __shfl_xor_sync(0xffffffffu, bp, o);
                        if (rpb_better(oc, op, bc, bp)) { bc = oc; bp = op; bidx = oi; }
                    }
                    if (tid == 0) {
                        if (cur != RPB_NIL) { const u32 cc = S.pcnt[cur]; const u64 cq = S.ppair[cur]; if (bidx == RPB_NIL || !rpb_better(bc, bp, cc, cq)) { bc = cc; bp = cq; bidx = cur; } }
                        // the cached best may have been decremented this round: then it is stale and a rescan is due
                        if (bidx != RPB_NIL && !(bc >= 2 && S.hcnt[S.pslot[bidx]] == bc)) sh.pbest_ok = 0; else sh.pbest = bidx;
                    }
                }
                if (tid == 0) sh.n_p = np;
            }
            __syncthreads();
            // ---- merge the pending list into the unread rest of M
            if (sh.n_p >= RPB_PMAX) {
                const u32 np = rpb_compact(S, sh, S.ppair, S.pcnt, S.pslot, 0, sh.n_p, S.tpair, S.tcnt, S.tslot);      // pending -> T (valid only)
                rpb_sort(S.tpair, S.tcnt, S.tslot, np, sh);
                for (u32 i = tid; i < np; i += RPB_THREADS) { S.ppair[i] = S.tpair[i]; S.pcnt[i] = S.tcnt[i]; S.pslot[i] = S.tslot[i]; }
                __syncthreads();
                const u32 nm = rpb_compact(S, sh, S.mpair, S.mcnt, S.mslot, sh.mpos, sh.mlen, S.tpair, S.tcnt, S.tslot);  // rest of M -> T
                // merge by rank: no two entries compare equal (a pair has one current count)
                for (u32 i = tid; i < nm; i += RPB_THREADS) {
                    const u32 c = S.tcnt[i]; const u64 p = S.tpair[i];
                    u32 lo = 0, hi = np;
                    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (rpb_better(S.pcnt[mid], S.ppair[mid], c, p)) lo = mid + 1; else hi = mid; }
                    S.mpair[i + lo] = p; S.mcnt[i + lo] = c; S.mslot[i + lo] = S.tslot[i];
                }
                for (u32 i = tid; i < np; i += RPB_THREADS) {
                    const u32 c = S.pcnt[i]; const u64 p = S.ppair[i];
                    u32 lo = 0, hi = nm;
                    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (rpb_better(S.tcnt[mid], S.tpair[mid], c, p)) lo = mid + 1; else hi = mid; }
                    S.mpair[i + lo] = p; S.mcnt[i + lo] = c; S.mslot[i + lo] = S.pslot[i];
                }
                __syncthreads();
                if (tid == 0) { sh.mpos = 0; sh.mlen = nm + np; sh.n_p = 0; sh.pbest_ok = 0; sh.pbest = RPB_NIL; }
                __syncthreads();
            }
        }
        __syncthreads();
        // ---- serialise: 'R','P', ULEB 256, ULEB nrules, rules, ULEB len, symbols   (V22.py:1889-1903)
        u8* dst = tmp + (size_t)bi.pbase * 4;
        // rule bytes: offsets by a scan over the rules
        u32 hdr = 0;
        if (tid == 0) { u8* p = dst; *p++ = 'R'; *p++ = 'P'; p = rpb_put_uleb(p, 256); p = rpb_put_uleb(p, nrules); sh.total = (u32)(p - dst); }
        __syncthreads();
        hdr = sh.total;
        u32 roff = hdr;
        for (u32 base = 0; base < nrules; base += RPB_THREADS) {
            const u32 r = base + tid;
            u32 sz = 0; u64 k = 0;
            if (r < nrules) { k = S.rules[r]; sz = rpb_uleb_size((u32)(k >> 32)) + rpb_uleb_size((u32)k); }
            u32 tot; const u32 ex = rpb_exscan(sz, sh, &tot);
            if (r < nrules) { u8* p = dst + roff + ex; p = rpb_put_uleb(p, (u32)(k >> 32)); rpb_put_uleb(p, (u32)k); }
            roff += tot;
        }
        // live symbols in index order
        u32 m = 0;
        for (u32 base = 0; base < n; base += RPB_THREADS) {
            const u32 i = base + tid;
            const u32 live = (i < n && S.sym[i] != RPB_NIL) ? 1u : 0u;
            u32 tot; rpb_exscan(live, sh, &tot);
            m += tot;
        }
        if (tid == 0) { u8* p = rpb_put_uleb(dst + roff, m); sh.total = (u32)(p - dst); }
        __syncthreads();
        u32 soff = sh.total;
        for (u32 base = 0; base < n; base += RPB_THREADS) {
            const u32 i = base + tid;
            u32 sz = 0, v = 0;
            if (i < n) { v = S.sym[i]; if (v != RPB_NIL) sz = rpb_uleb_size(v); }
            u32 tot; const u32 ex = rpb_exscan(sz, sh, &tot);
            if (sz) rpb_put_uleb(dst + soff + ex, v);
            soff += tot;
        }
        if (tid == 0) { bacc[(size_t)b * 64 + 32] = (u64)soff; err[b] = KOLM_OK; }
        __syncthreads();
    }
}

// Runs k_repair_big over the blocks longer than the shared-memory kernel's limit.  The slabs live in a pool that is allocated
// on first use and grows on demand (c->d_rpb); as many CTAs as slabs fit (at most one per SM) share the blocks through a counter.
int kolm_repair_big_impl(kolm_ctx* c, const u8* in, u8* tmp, cudaStream_t s) {
    const int nb = c->nblocks;
    u32* list = c->h_u32;
    int nbig = 0; u64 maxn = 0;
    for (int b = 0; b < nb; ++b) if (c->h_binfo[b].len > (u32)REPAIR_XL) { ++nbig; if (c->h_binfo[b].len > maxn) maxn = c->h_binfo[b].len; }
    if (!nbig) return KOLM_OK;
    CUDA_TRY(cudaStreamSynchronize(s));                      // `list` is pinned staging that earlier copies of this call may still read
    nbig = 0;
    for (int b = 0; b < nb; ++b) if (c->h_binfo[b].len > (u32)REPAIR_XL) list[nbig++] = (u32)b;
    const size_t need = rpb_need(maxn);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = free_b + c->rpb_bytes;                   // what the pool may use: what is free now plus what it already holds
    static long long pool_gib = -1, pool_pct = -1;           // KOLM_REPAIR_POOL_GIB / KOLM_REPAIR_POOL_PCT: cap of the slab pool (absolute, share of free memory)
    if (pool_gib < 0) { const char* e = getenv("KOLM_REPAIR_POOL_GIB"); pool_gib = e ? atoll(e) : 100; if (pool_gib < 1) pool_gib = 1; }
    if (pool_pct < 0) { const char* e = getenv("KOLM_REPAIR_POOL_PCT"); pool_pct = e ? atoll(e) : 60; if (pool_pct < 5 || pool_pct > 95) pool_pct = 60; }
    budget = budget / 100 * (size_t)pool_pct < ((size_t)pool_gib << 30) ? budget / 100 * (size_t)pool_pct : ((size_t)pool_gib << 30);
    static int per_sm = -1;                                  // KOLM_REPAIR_CTAS_PER_SM: the CTAs mostly wait on dependent loads, several per SM overlap
    if (per_sm < 0) { const char* e = getenv("KOLM_REPAIR_CTAS_PER_SM"); per_sm = e ? atoi(e) : RPB_CTAS_PER_SM; if (per_sm < 1) per_sm = 1; }
    int grid = nbig < per_sm * c->sm_count ? nbig : per_sm * c->sm_count;
    if ((size_t)grid * need > budget) grid = (int)(budget / need);
    if (grid < 1) return KOLM_E_CAPACITY;
    if ((size_t)grid * need > c->rpb_bytes) {
        if (c->d_rpb) { CUDA_TRY(cudaStreamSynchronize(s)); CUDA_TRY(cudaFree(c->d_rpb)); c->d_rpb = nullptr; c->rpb_bytes = 0; }
        CUDA_TRY(cudaMalloc(&c->d_rpb, (size_t)grid * need));
        c->rpb_bytes = (size_t)grid * need;
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_active, list, (size_t)nbig * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(c->d_stats + 13, 0, 4, s));
    KL(c, KC_MISC, c->total_bytes * 16, s, k_repair_big<<<grid, RPB_THREADS, 0, s>>>(in, c->d_binfo, c->d_active, nbig, (u8*)c->d_rpb, need, c->d_stats + 13, tmp, c->d_bacc, c->d_err));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}
