// bbwt_fwd.cu — Lyndon factorisation + bijective-BWT rotation sort.
//
// Replaces  duval_lyndon  (kolm_final.py:200-225 == kolm_final_researched_v2-2.py:326-349) and
//           bbwt_forward  (kolm_final.py:227-325 == kolm_final_researched_v2-2.py:351-423).
//
// One engine, two successor functions:
//   plain  : succ_h(i) = i+h (end of block = smallest)      -> suffix ranks (ISA) -> Lyndon starts
//            = strict prefix minima of the ISA, found by a look-back min-scan.
//   cyclic : succ_h(i) = rotate by h inside i's Lyndon factor -> ω-order of all rotations -> BBWT.
//
// Prefix doubling in the Manber–Myers form with Larsson–Sadakane pruning:
//   round(h): stream the current order `sa`; for every j emit pred_h(sa[j]) if its rank is not yet
//   final ("active").  The emitted list is already ordered by the *second* key, so one stable LSD
//   radix sort by the *first* key (the element's current group start, <= ceil(log2 n) bits) yields
//   the 2h-order of every non-singleton group.  Only active records are sorted; settled suffixes
//   are never touched again.  All blocks of the batch advance in the same launches.
//
// Radix passes: per-tile digit histogram -> per-block scan -> scatter.  The scatter kernel stages its
// tile of keys/values in shared memory with 1-D TMA (cp.async.bulk + mbarrier), ranks with
// __match_any_sync into warp-private histograms, reorders inside shared memory and writes digit runs
// back coalesced.
#include <stdlib.h>
#include "common.cuh"

#define NWARPS (KOLM_THREADS / 32)
#ifndef KOLM_BALLOT_MATCH
#define KOLM_BALLOT_MATCH 0     // measured: 8 ballots per item are slower than MATCH.ANY on sm_100a (scatter 13.9 vs 11.9 ms)
#endif
#define FULL 0xffffffffu
#define SIDX(x) ((x) + ((x) >> 4))                    // smem padding: 1 word per 16 (IPT) -> conflict-free blocked access
#define SPAD (KOLM_TILE + 2 + ((KOLM_TILE + 2) >> 4) + 1)

// ------------------------------------------------------------------------------------------------
// tile maps
// ------------------------------------------------------------------------------------------------
__global__ void k_build_tiles(const BlockInfo* __restrict__ binfo, const u32* __restrict__ tile0, const u32* __restrict__ tilen,
                              const u32* __restrict__ lens /*null -> binfo.len*/, TileDesc* __restrict__ tiles, int nblocks) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        u32 nt = tilen[b];
        if (!nt) continue;
        u32 t0 = tile0[b], len = lens ? lens[b] : binfo[b].len, pb = binfo[b].pbase;
        for (u32 k = threadIdx.x; k < nt; k += blockDim.x) {
            TileDesc td;
            td.start = pb + k * KOLM_TILE;
            u32 rem = len - k * KOLM_TILE;
            td.count = rem < KOLM_TILE ? rem : KOLM_TILE;
            td.block = b;
            td.flags = (k == 0 ? 1u : 0u) | (k == nt - 1 ? 2u : 0u);
            tiles[t0 + k] = td;
        }
    }
}

// single CTA: active tile ranges per block + totals.  stats[0]=active tiles, stats[1]=active records (saturating)
__global__ void k_plan_active(const u32* __restrict__ active, const u32* __restrict__ done, u32* __restrict__ atile0,
                              u32* __restrict__ atilen, u32* __restrict__ stats, int nblocks) {
    __shared__ u32 s_w[32];
    __shared__ u32 s_carry, s_rec, s_max;
    if (threadIdx.x == 0) { s_carry = 0; s_rec = 0; s_max = 0; }
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        int b = base + threadIdx.x;
        u32 a = (b < nblocks && !done[b]) ? active[b] : 0;
        u32 nt = (a + KOLM_TILE - 1) / KOLM_TILE;
        u32 v = nt;
        for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(FULL, v, o); if (lane_id() >= (u32)o) v += n; }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = v;
        u32 ra = a;
        for (int o = 16; o > 0; o >>= 1) ra += __shfl_xor_sync(FULL, ra, o);
        __syncthreads();
        u32 pre = 0;
        for (u32 i = 0; i < (threadIdx.x >> 5); ++i) pre += s_w[i];
        u32 carry = s_carry;
        if (b < nblocks) { atile0[b] = carry + pre + v - nt; atilen[b] = nt; }
        __syncthreads();
        if (lane_id() == 0) atomicAdd(&s_rec, ra);
        if (nt) atomicMax(&s_max, nt);
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[0] = s_carry; stats[1] = s_rec; stats[2] = s_max; }
}

// ------------------------------------------------------------------------------------------------
// Lyndon factor lookup (cyclic successor / predecessor)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void find_factor(const u32* __restrict__ fst, u32 nfac, u32 len, u32 lp, u32& fs, u32& fl) {
    u32 lo = 0, hi = nfac;
    if (lp >= fst[nfac - 1]) lo = nfac - 1;
    else while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (__ldg(fst + mid) <= lp) lo = mid; else hi = mid; }
    fs = __ldg(fst + lo);
    u32 fe = (lo + 1 < nfac) ? __ldg(fst + lo + 1) : len;
    fl = fe - fs;
}

// ------------------------------------------------------------------------------------------------
// bootstrap keys
//   plain : 3 bytes (zero padded) + 3-bit length code min(len-lp,4)   -> 27-bit key, h0 = 3
//           (every suffix of length <= 3 gets a unique key, so successor-less elements are settled)
//   cyclic: 4 bytes of the rotation (infinite power of the factor)     -> 32-bit key, h0 = 4
// ------------------------------------------------------------------------------------------------
template <bool CYCLIC>
__global__ void __launch_bounds__(KOLM_THREADS) k_boot_keys(const u8* __restrict__ in, const BlockInfo* __restrict__ binfo,
                                                            const TileDesc* __restrict__ tiles, const u32* __restrict__ fstart,
                                                            const u32* __restrict__ nfac, u32* __restrict__ K, u32* __restrict__ V) {
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff;
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 pg = td.start + x, lp = pg - bi.pbase, key;
        if (CYCLIC) {
            u32 fs, fl; find_factor(fstart + bi.pbase, nfac[td.block], bi.len, lp, fs, fl);
            u32 o = lp - fs; key = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) { key = (key << 8) | src[fs + o]; if (++o == fl) o = 0; }
        } else {
            u32 rem = bi.len - lp;
            u32 b0 = src[lp], b1 = rem > 1 ? src[lp + 1] : 0, b2 = rem > 2 ? src[lp + 2] : 0;
            key = (((b0 << 16) | (b1 << 8) | b2) << 3) | (rem < 4 ? rem : 4);
        }
        K[pg] = key; V[pg] = pg;
    }
}


// ------------------------------------------------------------------------------------------------
// Alphabet compression for the cyclic bootstrap: a block that uses sigma distinct byte values needs ceil(log2 sigma)
// bits per symbol, so a 32-bit key holds h0 = floor(32 / bits) >= 4 symbols of the rotation instead of 4 bytes (text:
// sigma ~ 28 -> 5 bits -> h0 = 6).  Dense ranks preserve byte order, so the keys stay order preserving; h0 is the minimum
// over the batch because every round uses one h for all blocks.
//   per block in bacc: u32 mask[8] at u64 slot 0..3, u8 rmap[256] at u64 slot 8..39, u32 bits at u64 slot 40
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KOLM_THREADS) k_alpha_mask(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                             const BlockInfo* __restrict__ binfo, u64* __restrict__ bacc) {
    // presence flags, one byte per symbol: every occurrence is a plain store of 1 (all writers agree), three instructions per
    // input byte instead of the eight compare-and-select pairs of a register mask (0.39 -> 0.1 ms per 256 MiB)
    __shared__ u8 present[256];
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    present[threadIdx.x] = 0;
    __syncthreads();
    const u8* src = in + bi.ioff + (td.start - bi.pbase);
    if (((uintptr_t)src & 15) == 0) {                          // 16 bytes per load
        const u32 nv = td.count >> 4;
        for (u32 x = threadIdx.x; x < nv; x += KOLM_THREADS) {
            const uint4 q = reinterpret_cast<const uint4*>(src)[x];
            const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) present[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu] = 1;
        }
        for (u32 x = (nv << 4) + threadIdx.x; x < td.count; x += KOLM_THREADS) present[src[x]] = 1;
    } else {
        for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) present[src[x]] = 1;
    }
    __syncthreads();
    const u32 v = __ballot_sync(FULL, present[threadIdx.x] != 0);   // warp k holds symbols 32k .. 32k+31
    if ((threadIdx.x & 31) == 0 && v) atomicOr(reinterpret_cast<u32*>(bacc + (size_t)td.block * 64) + (threadIdx.x >> 5), v);
}


__global__ void __launch_bounds__(256) k_alpha_map(u64* __restrict__ bacc, const BlockInfo* __restrict__ binfo, u32* __restrict__ min_syms, int nblocks) {
    const int b = blockIdx.x;
    if (b >= nblocks) return;
    u32* mask = reinterpret_cast<u32*>(bacc + (size_t)b * 64);
    u8* rmap = reinterpret_cast<u8*>(bacc + (size_t)b * 64 + 8);
    const u32 c = threadIdx.x;
    u32 below = 0;
    for (u32 k = 0; k < (c >> 5); ++k) below += __popc(mask[k]);
    below += __popc(mask[c >> 5] & ((1u << (c & 31)) - 1u));
    rmap[c] = (u8)below;
    if (c == 255) {
        u32 sigma = below + ((mask[7] >> 31) & 1u);
        u32 bits = 1; while ((1u << bits) < sigma) ++bits;
        reinterpret_cast<u32*>(bacc + (size_t)b * 64 + 40)[0] = bits;
        if (binfo[b].len) atomicMin(min_syms, 32u / bits);
    }
}

// The tile's symbol codes are staged once in shared memory; a thread owns 16 consecutive positions and slides a window over
// them (one code in, one out per key).  Only positions whose h0-symbol window crosses the end of their Lyndon factor (the
// rotation wraps) take the per-position path.
__global__ void __launch_bounds__(KOLM_THREADS) k_boot_keys_alpha(const u8* __restrict__ in, const BlockInfo* __restrict__ binfo,
                                                                  const TileDesc* __restrict__ tiles, const u32* __restrict__ fstart,
                                                                  const u32* __restrict__ nfac, const u64* __restrict__ bacc, u32 h0,
                                                                  u32* __restrict__ K, u32* __restrict__ V) {
    __shared__ u8 rmap[256];
    __shared__ __align__(16) u8 sc[KOLM_TILE + 32];
    const u32 tid = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    rmap[tid] = reinterpret_cast<const u8*>(bacc + (size_t)td.block * 64 + 8)[tid];
    const u32 bits = reinterpret_cast<const u32*>(bacc + (size_t)td.block * 64 + 40)[0];
    __syncthreads();
    const u8* src = in + bi.ioff;
    const u32 t0 = td.start - bi.pbase;
    {   // stage codes of block positions [t0, t0 + count + 32) (zero beyond the block)
        const u32 want = min(td.count + 32u, bi.len - t0);
        const u8* g = src + t0;
        if (((uintptr_t)g & 15) == 0) {
            for (u32 x = tid * 16; x < KOLM_TILE + 32; x += KOLM_THREADS * 16) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (x + 16 <= want) v = *reinterpret_cast<const uint4*>(g + x);
                else if (x < want) { u32 w[4] = {0, 0, 0, 0}; for (u32 i = 0; x + i < want; ++i) w[i >> 2] |= (u32)g[x + i] << (8 * (i & 3)); v = make_uint4(w[0], w[1], w[2], w[3]); }
                u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    w[i] = (u32)rmap[w[i] & 0xFF] | ((u32)rmap[(w[i] >> 8) & 0xFF] << 8) | ((u32)rmap[(w[i] >> 16) & 0xFF] << 16) | ((u32)rmap[w[i] >> 24] << 24);
                *reinterpret_cast<uint4*>(sc + x) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            for (u32 x = tid; x < KOLM_TILE + 32; x += KOLM_THREADS) sc[x] = x < want ? rmap[g[x]] : 0;
        }
    }
    __syncthreads();
    const u32 r0 = tid * KOLM_IPT;
    if (r0 >= td.count) return;
    const u32 nmine = min((u32)KOLM_IPT, td.count - r0);
    const u32* fst = fstart + bi.pbase;
    const u32 nf = nfac[td.block];
    u32 fs, fl; find_factor(fst, nf, bi.len, t0 + r0, fs, fl);
    const u32 kmask = bits * h0 >= 32 ? 0xffffffffu : ((1u << (bits * h0)) - 1u);
    u32 keys[KOLM_IPT];
    if (t0 + r0 + nmine - 1 + h0 <= fs + fl) {               // every window of mine stays inside one factor: slide
        u32 key = 0;
        for (u32 t = 0; t + 1 < h0; ++t) key = (key << bits) | sc[r0 + t];
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) { key = ((key << bits) | sc[r0 + i + h0 - 1]) & kmask; keys[i] = key; }
    } else {
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {
            keys[i] = 0;
            if ((u32)i < nmine) {
                const u32 lp = t0 + r0 + i;
                if (lp >= fs + fl) find_factor(fst, nf, bi.len, lp, fs, fl);
                u32 o = lp - fs, key = 0;
                for (u32 t = 0; t < h0; ++t) { key = (key << bits) | rmap[src[fs + o]]; if (++o == fl) o = 0; }
                keys[i] = key;
            }
        }
    }
    const u32 pg = td.start + r0;
    if (nmine == KOLM_IPT) {                                 // td.start is a multiple of 32 elements: 16-byte aligned stores
#pragma unroll
        for (int i = 0; i < KOLM_IPT; i += 4) {
            *reinterpret_cast<uint4*>(K + pg + i) = make_uint4(keys[i], keys[i + 1], keys[i + 2], keys[i + 3]);
            *reinterpret_cast<uint4*>(V + pg + i) = make_uint4(pg + i, pg + i + 1, pg + i + 2, pg + i + 3);
        }
    } else {
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) if ((u32)i < nmine) { K[pg + i] = keys[i]; V[pg + i] = pg + i; }
    }
}

// ------------------------------------------------------------------------------------------------
// Deep bootstrap (cyclic sorts of low-entropy batches): the order by 2*h0 symbols comes from two plain LSD sorts instead of
// the bootstrap sort plus a doubling round — KH[p] = key of the h0 symbols at p (by position);
//   k_boot_lo   : K0[p] = KH[succ_h0(p)] = the key of symbols h0..2*h0-1, first sort;
//   k_boot_rekey: K[j] = KH[V[j]], second (stable) sort; k_rerank<2> then compares (K, KH[succ_h0(V)]).
// A doubling round over ~all records costs a gather, three passes, a rerank and a rank scatter (measured 14.5 ms on the
// 256 MiB text batch); four more plain passes and the re-key cost about 8.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KOLM_THREADS) k_boot_lo(const u32* __restrict__ KH, const BlockInfo* __restrict__ binfo, const TileDesc* __restrict__ tiles,
                                                          const u32* __restrict__ fstart, const u32* __restrict__ nfac, u32 h0, u32* __restrict__ K0,
                                                          u32* __restrict__ LO) {
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    const u32 r0 = threadIdx.x * KOLM_IPT;
    if (r0 >= td.count) return;
    const u32 nmine = min((u32)KOLM_IPT, td.count - r0);
    const u32* fst = fstart + bi.pbase;
    const u32 nf = nfac[td.block];
    const u32 lp0 = td.start - bi.pbase + r0;
    u32 fs, fl; find_factor(fst, nf, bi.len, lp0, fs, fl);
    const u32* kh = KH + bi.pbase;
    u32 keys[KOLM_IPT];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        keys[i] = 0;
        if ((u32)i < nmine) {
            const u32 lp = lp0 + i;
            if (lp >= fs + fl) find_factor(fst, nf, bi.len, lp, fs, fl);
            u32 o = lp - fs + h0;
            if (o >= fl) o %= fl;
            keys[i] = kh[fs + o];
        }
    }
    // the same keys twice: K0 is the first sort's key column (scrambled by the passes), LO stays by position for k_rerank<2>
    u32* dst = K0 + td.start + r0;
    u32* dlo = LO + td.start + r0;
    if (nmine == KOLM_IPT) {
#pragma unroll
        for (int i = 0; i < KOLM_IPT; i += 4) {
            const uint4 q = make_uint4(keys[i], keys[i + 1], keys[i + 2], keys[i + 3]);
            *reinterpret_cast<uint4*>(dst + i) = q; *reinterpret_cast<uint4*>(dlo + i) = q;
        }
    } else {
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) if ((u32)i < nmine) { dst[i] = keys[i]; dlo[i] = keys[i]; }
    }
}

__global__ void __launch_bounds__(KOLM_THREADS) k_boot_rekey(const u32* __restrict__ KH, const u32* __restrict__ V, const TileDesc* __restrict__ tiles,
                                                             u32* __restrict__ K) {
    const TileDesc td = tiles[blockIdx.x];
    u32 v[KOLM_IPT];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { u32 x = i * KOLM_THREADS + threadIdx.x; v[i] = x < td.count ? V[td.start + x] : 0xffffffffu; }
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) if (v[i] != 0xffffffffu) v[i] = KH[v[i]];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { u32 x = i * KOLM_THREADS + threadIdx.x; if (x < td.count) K[td.start + x] = v[i]; }
}

// ------------------------------------------------------------------------------------------------
// radix pass 1/3: per-tile digit histogram
// ------------------------------------------------------------------------------------------------
// Slot of digit d inside a 256-entry shared-memory table.  Digits that share their low five bits share a bank, and keys packed
// from 5-bit symbol codes make exactly those digits frequent together (same last symbol, different neighbour): ncu showed 60 %
// of the scatter's shared-memory wavefronts as bank conflicts on such passes.  Adding the high bits into the bank index spreads
// them over eight banks; the map is a bijection inside every group of 32, so a warp that walks 32 consecutive digits stays
// conflict free.
__device__ __forceinline__ u32 dswz(u32 d) { return (d & ~31u) | ((d + (d >> 5)) & 31u); }

__global__ void __launch_bounds__(KOLM_THREADS) k_radix_hist(const u32* __restrict__ K, const TileDesc* __restrict__ tiles,
                                                             u32* __restrict__ thist, int shift, u32 mask) {
    __shared__ u32 wh[NWARPS][256];                        // warp-private histograms: conflicts only inside a warp
    TileDesc td = tiles[blockIdx.x];
    const u32 tid = threadIdx.x, w = tid >> 5;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) wh[i][tid] = 0;
    __syncthreads();
    const u32* k = K + td.start;                            // 128-byte aligned (padded index space)
    const u32 full4 = td.count >> 2;
    for (u32 x = tid; x < full4; x += KOLM_THREADS) {
        uint4 q = reinterpret_cast<const uint4*>(k)[x];
        atomicAdd(&wh[w][dswz((q.x >> shift) & mask)], 1u); atomicAdd(&wh[w][dswz((q.y >> shift) & mask)], 1u);
        atomicAdd(&wh[w][dswz((q.z >> shift) & mask)], 1u); atomicAdd(&wh[w][dswz((q.w >> shift) & mask)], 1u);
    }
    for (u32 x = (full4 << 2) + tid; x < td.count; x += KOLM_THREADS) atomicAdd(&wh[w][dswz((k[x] >> shift) & mask)], 1u);
    __syncthreads();
    u32 tot = 0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) tot += wh[i][dswz(tid)];
    thist[(size_t)blockIdx.x * 256 + tid] = tot;
}

// radix pass 2/3: per block, turn tile histograms into block-relative scatter bases (digit-major, tile-minor).
// 1024 threads = four row groups x 256 digits: every group owns a quarter of the block's tiles, sums its column first (loads
// batched eight deep: the serial one-thread-per-digit walk over 256 tiles was pure latency, ~50 us per launch at 1 MiB blocks),
// then rewrites it as running bases on top of (digit start + the earlier groups' column sums).
#define RSCAN_GROUPS 4
__global__ void __launch_bounds__(256 * RSCAN_GROUPS) k_radix_scan(u32* __restrict__ thist, const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks) {
    __shared__ u32 part[RSCAN_GROUPS][256];
    __shared__ u32 excl[256];
    const u32 d = threadIdx.x & 255u, q = threadIdx.x >> 8;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const u32 nt = tilen[b];
        if (!nt) continue;
        const u32 per = (nt + RSCAN_GROUPS - 1) / RSCAN_GROUPS;
        const u32 ta = min(nt, q * per), tb = min(nt, ta + per);
        u32* base = thist + (size_t)tile0[b] * 256 + d;
        u32 sum = 0;
        {
            u32 t = ta;
            for (; t + 8 <= tb; t += 8) {
                u32 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = base[(size_t)(t + i) * 256];
#pragma unroll
                for (int i = 0; i < 8; ++i) sum += v[i];
            }
            for (; t < tb; ++t) sum += base[(size_t)t * 256];
        }
        part[q][d] = sum;
        __syncthreads();
        if (threadIdx.x < 32) {
            u32 s = 0, loc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                loc[i] = s;
#pragma unroll
                for (int g = 0; g < RSCAN_GROUPS; ++g) s += part[g][threadIdx.x * 8 + i];
            }
            u32 v = s;
            for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(FULL, v, o); if (threadIdx.x >= (u32)o) v += n; }
            u32 pre = v - s;
#pragma unroll
            for (int i = 0; i < 8; ++i) excl[threadIdx.x * 8 + i] = pre + loc[i];
        }
        __syncthreads();
        u32 run = excl[d];
        for (u32 g = 0; g < q; ++g) run += part[g][d];
        {
            u32 t = ta;
            for (; t + 8 <= tb; t += 8) {
                u32 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = base[(size_t)(t + i) * 256];
#pragma unroll
                for (int i = 0; i < 8; ++i) { base[(size_t)(t + i) * 256] = run; run += v[i]; }
            }
            for (; t < tb; ++t) { const u32 v = base[(size_t)t * 256]; base[(size_t)t * 256] = run; run += v; }
        }
        __syncthreads();
    }
}

// radix pass 3/3: stable scatter of one tile.  Keys/values staged with 1-D TMA into shared memory.
// The kernel is bound by shared-memory wavefronts (ncu: ~1 per element), so the reorder step keeps them low: the per-warp
// digit offsets are folded with the digit starts (one table lookup per element instead of two) and the global base is
// pre-reduced by the digit start.  (Moving key and value as one 64-bit word costs 30 more registers and a CTA per SM.)
// A warp whose first 32 records hold more than this many distinct digits ranks the rest of its records with eight ballots
// instead of MATCH.ANY (measured on the 256 MiB bench workload: threshold 18 -> scatter 11.66 -> 11.09 ms, 10 -> no gain).
#ifndef KOLM_BALLOT_THRESH
#define KOLM_BALLOT_THRESH 18
#endif
__global__ void __launch_bounds__(KOLM_THREADS) k_radix_scatter(const u32* __restrict__ Kin, const u32* __restrict__ Vin,
                                                                u32* __restrict__ Kout, u32* __restrict__ Vout,
                                                                const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                                const u32* __restrict__ thist, int shift, u32 mask, const u32* __restrict__ rekey) {
    __shared__ __align__(128) u32 sk[KOLM_TILE];
    __shared__ __align__(128) u32 sv[KOLM_TILE];
    __shared__ u32 whist[NWARPS][256];
    __shared__ u32 dstart[256];
    __shared__ u32 gofs[256];
    __shared__ __align__(8) u64 bar;
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    TileDesc td = tiles[blockIdx.x];
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (u32 i = tid; i < NWARPS * 256; i += KOLM_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    if (tid == 0) {
        u32 bytes = ((td.count + 3u) & ~3u) * 4u;       // 16-byte granules; block bases are padded so the over-read stays in scratch
        mbar_expect_tx(&bar, 2 * bytes);
        tma_load_1d(sk, Kin + td.start, bytes, &bar);
        tma_load_1d(sv, Vin + td.start, bytes, &bar);
    }
    const u32 gb = thist[(size_t)blockIdx.x * 256 + tid] + binfo[td.block].pbase;
    mbar_wait(&bar, 0);
    // ---- rank: warp w owns elements [w*IPT*32, (w+1)*IPT*32), iteration k covers 32 consecutive ones
    u32 key[KOLM_IPT], val[KOLM_IPT]; u16 off[KOLM_IPT];
    bool use_ballot = false;
#pragma unroll
    for (int k = 0; k < KOLM_IPT; ++k) {
        u32 idx = w * (KOLM_IPT * 32) + k * 32 + lane;
        bool valid = idx < td.count;
        key[k] = valid ? sk[idx] : 0; val[k] = valid ? sv[idx] : 0;
        u32 d = valid ? ((key[k] >> shift) & mask) : 0xffffffffu;
#if KOLM_BALLOT_MATCH
        u32 peers = __ballot_sync(FULL, valid);
        if (!valid) peers = ~peers;
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) { u32 m = __ballot_sync(FULL, (d >> bb) & 1u); peers &= ((d >> bb) & 1u) ? m : ~m; }
#else
        u32 peers;
        if (use_ballot) {                                   // many distinct digits in this warp: eight ballots cost the same for any mix
            peers = __ballot_sync(FULL, valid);
            if (!valid) peers = ~peers;
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) { u32 m = __ballot_sync(FULL, (d >> bb) & 1u); peers &= ((d >> bb) & 1u) ? m : ~m; }
        } else peers = __match_any_sync(FULL, d);           // MATCH.ANY: cost grows with the number of distinct values
#endif
        u32 lt = __popc(peers & lanemask_lt());
#if !KOLM_BALLOT_MATCH && defined(KOLM_BALLOT_THRESH)
        if (k == 0) use_ballot = __popc(__ballot_sync(FULL, lt == 0)) > KOLM_BALLOT_THRESH;   // distinct digits among the first 32 records
#endif
        u32 old = 0;
        if (valid && lt == 0) { const u32 ds = dswz(d); old = whist[w][ds]; whist[w][ds] = old + __popc(peers); }
        old = __shfl_sync(FULL, old, __ffs(peers) - 1);
        off[k] = (u16)(old + lt);
        __syncwarp();
    }
    __syncthreads();                                        // every thread holds its records in registers: the staging area is free
    // ---- per digit: exclusive scan over warps, then over digits
    {
        u32 run = 0;
#pragma unroll
        for (int i = 0; i < NWARPS; ++i) { u32 t = whist[i][dswz(tid)]; whist[i][dswz(tid)] = run; run += t; }
        dstart[tid] = run;     // total for digit tid (scanned below)
    }
    __syncthreads();
    if (tid < 32) {
        u32 s = 0, loc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { loc[i] = s; s += dstart[tid * 8 + i]; }
        u32 v = s;
        for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(FULL, v, o); if (tid >= (u32)o) v += n; }
        u32 pre = v - s;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) dstart[tid * 8 + i] = pre + loc[i];
    }
    __syncthreads();
    {   // whist[i][d] = slot of warp i's first element of digit d inside the tile; gofs[d] = global base - digit start
        const u32 ds = dstart[tid];
#pragma unroll
        for (int i = 0; i < NWARPS; ++i) whist[i][dswz(tid)] += ds;
        gofs[tid] = gb - ds;
    }
    __syncthreads();
    // ---- reorder inside shared memory (everything is in registers now)
#pragma unroll
    for (int k = 0; k < KOLM_IPT; ++k) {
        u32 idx = w * (KOLM_IPT * 32) + k * 32 + lane;
        if (idx < td.count) {
            u32 d = (key[k] >> shift) & mask;
            const u32 pos = whist[w][dswz(d)] + off[k];
            sk[pos] = key[k]; sv[pos] = val[k];
        }
    }
    __syncthreads();
    // ---- coalesced digit runs to global
    if (rekey) {                                            // last pass of the deep bootstrap's first sort: the key column leaves as rekey[value]
        for (u32 s0 = tid; s0 < td.count; s0 += 4 * KOLM_THREADS) {
            u32 g[4], v[4], nk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const u32 s = s0 + u * KOLM_THREADS; g[u] = 0xffffffffu; if (s < td.count) { g[u] = gofs[(sk[s] >> shift) & mask] + s; v[u] = sv[s]; } }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (g[u] != 0xffffffffu) nk[u] = rekey[v[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (g[u] != 0xffffffffu) { Kout[g[u]] = nk[u]; Vout[g[u]] = v[u]; }
        }
        return;
    }
    for (u32 s = tid; s < td.count; s += KOLM_THREADS) {
        const u32 kk = sk[s];
        const u32 g = gofs[(kk >> shift) & mask] + s;
        Kout[g] = kk; Vout[g] = sv[s];
    }
}

// ------------------------------------------------------------------------------------------------
// rerank: group heads, new ranks, new order.  BOOT 1: records are (key32, pos) of whole blocks.
// BOOT 2 (deep bootstrap): records are sorted by the 64-bit key (key32 of the rotation, key32 of the rotation h symbols
// later); the high half travels in K, the low half is looked up as KH[succ_h(pos)] (KH = a.nr, the keys by position).
// round : records are (group start, pos) of active elements; second key = rank[succ_h(pos)].
// ------------------------------------------------------------------------------------------------
struct RerankArgs {
    const u32* K; const u32* V; const TileDesc* tiles; const BlockInfo* binfo; const u32* active;
    const u32* fstart; const u32* nfac; u64* lb;
    u32* sa; u32* rank; u32* nr; u32* single; u32* newcls; u32 h;
    u32* survivors;                     // [1] records that remain unsettled after this round (non-BOOT only)
    const u32* lo;                      // BOOT 2: low key half by position (k_boot_lo), so the rerank needs no factor search
    u32* grp;                           // group start by order index (local refinement rounds), may be null
};

template <int BOOT, bool CYCLIC>
__global__ void __launch_bounds__(KOLM_THREADS, 4) k_rerank(RerankArgs a) {
    __shared__ u32 sk[SPAD];
    __shared__ u32 sk2[SPAD];
    __shared__ u64 s_warp[NWARPS];
    __shared__ u64 s_excl;
    __shared__ u32 s_cnt, s_surv;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(a.lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = a.tiles[tile];
    const BlockInfo bi = a.binfo[td.block];
    const u32 nrec = BOOT ? bi.len : a.active[td.block];
    const u32 t0 = td.start - bi.pbase;
    if (tid == 0) { s_cnt = 0; s_surv = 0; }
    u32 nsurv = 0;
    // stage (key, key2) of records t0-1 .. t0+count into smem slots 0 .. count+1.  Loads are issued in
    // phases (all K/V, then all rank gathers) so that every thread keeps IPT+1 requests in flight.
    {
        constexpr int NS = KOLM_IPT + 1;                   // slot x = i*THREADS + tid, i < NS covers count+2 <= TILE+2
        u32 kk[NS], vv[NS]; bool ok[NS];
        const u32 nf = (BOOT == 0 && CYCLIC) ? a.nfac[td.block] : 0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            u32 x = i * KOLM_THREADS + tid;
            i64 tl = (i64)t0 + x - 1;
            ok[i] = x < td.count + 2 && tl >= 0 && tl < (i64)nrec;
            u32 g = bi.pbase + (u32)tl;
            kk[i] = ok[i] ? a.K[g] : 0xffffffffu;
            vv[i] = (ok[i] && BOOT != 1) ? a.V[g] : 0;
        }
        u32 k2[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            k2[i] = 0xffffffffu;
            if (BOOT == 1) { if (ok[i]) { k2[i] = kk[i]; kk[i] = 0; } }
            else if (BOOT == 2) { /* second key = lo[position]: no successor arithmetic */ }
            else if (ok[i]) {
                u32 lp = vv[i] - bi.pbase, sp;
                if (CYCLIC) { u32 fs, fl; find_factor(a.fstart + bi.pbase, nf, bi.len, lp, fs, fl); { u32 o = lp - fs + a.h % fl; sp = fs + (o >= fl ? o - fl : o); } }
                else sp = lp + a.h;                        // active plain elements always have a successor (see boot keys)
                vv[i] = bi.pbase + sp;
            }
        }
        if (BOOT != 1) {
            const u32* second = BOOT == 2 ? a.lo : a.rank;
#pragma unroll
            for (int i = 0; i < NS; ++i) if (ok[i]) k2[i] = second[vv[i]];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            u32 x = i * KOLM_THREADS + tid;
            if (x < td.count + 2) { sk[SIDX(x)] = kk[i]; sk2[SIDX(x)] = k2[i]; }
        }
    }
    __syncthreads();
    // per-thread blocked scan: 1-based local record index of the last group-first / last head
    u32 lf = 0, lh = 0, nh = 0;
    u32 firstbits = 0, headbits = 0, nextbits = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        if (r < td.count) {
            u32 x = r + 1;
            u32 kp = sk[SIDX(x - 1)], kc = sk[SIDX(x)], kn = sk[SIDX(x + 1)];
            u32 qp = sk2[SIDX(x - 1)], qc = sk2[SIDX(x)], qn = sk2[SIDX(x + 1)];
            const bool kd = kc != kp;
            const bool first = BOOT == 2 ? (t0 + r == 0) : kd;     // deep bootstrap: the whole block is one group, K is part of the key
            const bool head = first || kd || qc != qp, nxt = (kn != kc) || (qn != qc) || (BOOT == 2 && t0 + r + 1 == nrec);
            if (first) { lf = t0 + r + 1; firstbits |= 1u << i; }
            if (head) { lh = t0 + r + 1; headbits |= 1u << i; nh += first ? (BOOT ? 1u : 0u) : 1u; }
            if (nxt) nextbits |= 1u << i;
        }
    }
    // records of my items: issued before the scans so the loads overlap the block scan and the look-back
    const u32* Vt = a.V + td.start;
    const u32* Kt = a.K + td.start;
    u32 val2[KOLM_IPT], key2r[KOLM_IPT];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        val2[i] = r < td.count ? Vt[r] : 0u;
        key2r[i] = (!BOOT && r < td.count) ? Kt[r] : 0u;
    }
    u64 agg = ((u64)lf << 31) | lh, tot;
    u64 incl = block_scan_incl(agg, 0ull, OpMax2(), s_warp, &tot);
    // exclusive for this thread = scan of preceding threads
    u64 prev = __shfl_up_sync(FULL, incl, 1);
    __shared__ u64 s_last[NWARPS];
    if ((tid & 31) == 31) s_last[tid >> 5] = incl;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : 0ull;
    if (tid < 32) {
        u64 e = lb_exclusive(a.lb, tile, (td.flags & 1u) != 0, tot, 0ull, OpMax2());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u64 ex = OpMax2()(s_excl, prev);
    u32 cf = (u32)(ex >> 31), ch = (u32)(ex & 0x7fffffffu);
    u32 pos_[KOLM_IPT], nrk_[KOLM_IPT]; u32 singles = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 tl1 = t0 + tid * KOLM_IPT + i + 1;
        if ((firstbits >> i) & 1u) cf = tl1;
        if ((headbits >> i) & 1u) ch = tl1;
        pos_[i] = key2r[i] + (tl1 - cf);
        nrk_[i] = key2r[i] + (ch - cf);
        if (((headbits >> i) & 1u) && ((nextbits >> i) & 1u)) singles |= 1u << i;
    }
    // bootstrap reranks place record r of the tile at order index t0 + r: 16 consecutive words per thread, stored as four uint4
    const bool vec = BOOT != 0 && tid * KOLM_IPT + KOLM_IPT <= td.count;
    if (vec) {
        uint4* d = reinterpret_cast<uint4*>(a.sa + td.start + tid * KOLM_IPT);
#pragma unroll
        for (int i = 0; i < KOLM_IPT; i += 4) d[i >> 2] = make_uint4(val2[i], val2[i + 1], val2[i + 2], val2[i + 3]);
        if (a.grp) {
            uint4* g = reinterpret_cast<uint4*>(a.grp + td.start + tid * KOLM_IPT);
#pragma unroll
            for (int i = 0; i < KOLM_IPT; i += 4) g[i >> 2] = make_uint4(nrk_[i], nrk_[i + 1], nrk_[i + 2], nrk_[i + 3]);
        }
    }
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        if (r < td.count) {
            const bool single = (singles >> i) & 1u;
            if (!vec) {
                a.sa[bi.pbase + pos_[i]] = val2[i];
                if (a.grp) a.grp[bi.pbase + pos_[i]] = nrk_[i];
            }
            if (BOOT) {
                a.rank[val2[i]] = nrk_[i];
                if (single && a.single) atomicOr(a.single + (val2[i] >> 5), 1u << (val2[i] & 31));
            } else { a.nr[td.start + r] = nrk_[i] | (single ? 0x80000000u : 0u); nsurv += single ? 0u : 1u; }
        }
    }
    // classes created in this tile (+ records left unsettled, packed in the high half)
    for (int o = 16; o > 0; o >>= 1) { nh += __shfl_xor_sync(FULL, nh, o); nsurv += __shfl_xor_sync(FULL, nsurv, o); }
    if ((tid & 31) == 0 && nh) atomicAdd(&s_cnt, nh);
    if ((tid & 31) == 0 && nsurv) atomicAdd(&s_surv, nsurv);
    __syncthreads();
    if (tid == 0 && s_cnt) atomicAdd(a.newcls + td.block, s_cnt);
    if (!BOOT && tid == 0 && s_surv) atomicAdd(a.survivors, s_surv);
}

__global__ void __launch_bounds__(KOLM_THREADS) k_apply(const u32* __restrict__ V, const u32* __restrict__ nr, const TileDesc* __restrict__ tiles,
                                                        u32* __restrict__ rank, u32* __restrict__ single) {
    TileDesc td = tiles[blockIdx.x];
    u32 v[KOLM_IPT], r[KOLM_IPT];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 x = i * KOLM_THREADS + threadIdx.x;
        bool ok = x < td.count;
        v[i] = ok ? V[td.start + x] : 0xffffffffu;
        r[i] = ok ? nr[td.start + x] : 0;
    }
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if (v[i] != 0xffffffffu) {
            rank[v[i]] = r[i] & 0x7fffffffu;
            if (r[i] >> 31) atomicOr(single + (v[i] >> 5), 1u << (v[i] & 31));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gather (Manber–Myers step): stream sa, emit active predecessors in order, compacted per block
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
    const u32* sa; const u32* rank; const u32* single; const TileDesc* tiles; const BlockInfo* binfo;
    const u32* fstart; const u32* nfac; const u32* done; u64* lb; u32* K; u32* V; u32* active; u32 h;
};

template <bool CYCLIC>
__global__ void __launch_bounds__(KOLM_THREADS, 5) k_gather(GatherArgs a) {
    __shared__ u64 s_warp[NWARPS];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(a.lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = a.tiles[tile];
    const BlockInfo bi = a.binfo[td.block];
    const bool done = a.done[td.block] != 0;
    u32 pk[KOLM_IPT], pv[KOLM_IPT];
    u32 amask = 0;
    if (!done) {
        u32 nf = CYCLIC ? a.nfac[td.block] : 0;
        const u32* fst = a.fstart + bi.pbase;
        u32 sv[KOLM_IPT], sw[KOLM_IPT];
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {               // phase 1: the order itself (coalesced)
            u32 r = tid * KOLM_IPT + i;
            sv[i] = r < td.count ? a.sa[td.start + r] : 0xffffffffu;
        }
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {               // phase 2: predecessor positions
            pv[i] = 0xffffffffu;
            if (sv[i] != 0xffffffffu) {
                u32 lp = sv[i] - bi.pbase;
                if (CYCLIC) { u32 fs, fl; find_factor(fst, nf, bi.len, lp, fs, fl); u32 hm = a.h % fl; u32 o = lp - fs; pv[i] = bi.pbase + fs + (o >= hm ? o - hm : o + fl - hm); }
                else if (lp >= a.h) pv[i] = bi.pbase + lp - a.h;
            }
        }
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) sw[i] = pv[i] != 0xffffffffu ? a.single[pv[i] >> 5] : 0xffffffffu;   // phase 3: settled bits
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) if (pv[i] != 0xffffffffu && !((sw[i] >> (pv[i] & 31)) & 1u)) amask |= 1u << i;
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) pk[i] = ((amask >> i) & 1u) ? a.rank[pv[i]] : 0;                     // phase 4: first keys
    }
    const u32 cnt = __popc(amask);
    u64 tot;
    u64 incl = block_scan_incl((u64)cnt, 0ull, OpAdd(), s_warp, &tot);
    if (tid < 32) {
        u64 e = lb_exclusive(a.lb, tile, (td.flags & 1u) != 0, tot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u32 o = bi.pbase + (u32)s_excl + (u32)(incl - cnt);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) if ((amask >> i) & 1u) { u32 d = o + __popc(amask & ((1u << i) - 1u)); a.K[d] = pk[i]; a.V[d] = pv[i]; }
    if (tid == 0 && (td.flags & 2u)) a.active[td.block] = (u32)(s_excl + tot);
}


// ------------------------------------------------------------------------------------------------
// Larsson–Sadakane rounds for small active sets.  The Manber–Myers gather streams the whole order every round; once few
// records are left (< 1/8 of the batch) it is cheaper to start from the survivors of the previous round:
//   k_ls_build: survivors (v) -> (second key rank[succ_h(v)], v), compacted per block; sort by the second key;
//   k_ls_key1 : replace the key by the first key rank[v]; stable sort by it -> same record order the gather path produces.
// ------------------------------------------------------------------------------------------------
struct LsArgs {
    const u32* V; const u32* nr; const TileDesc* tiles; const BlockInfo* binfo; const u32* rank; const u32* fstart; const u32* nfac;
    const u32* done; u64* lb; u32* K2; u32* V2; u32* active; u32 h;
};

template <bool CYCLIC>
__global__ void __launch_bounds__(KOLM_THREADS) k_ls_build(LsArgs a) {
    __shared__ u64 s_warp[NWARPS];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(a.lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = a.tiles[tile];
    const BlockInfo bi = a.binfo[td.block];
    const bool done = a.done[td.block] != 0;
    u32 pv[KOLM_IPT], pk[KOLM_IPT];
    u32 amask = 0;
    if (!done) {
        const u32 nf = CYCLIC ? a.nfac[td.block] : 0;
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {
            u32 r = tid * KOLM_IPT + i;
            pv[i] = 0;
            if (r < td.count && !(a.nr[td.start + r] >> 31)) { pv[i] = a.V[td.start + r]; amask |= 1u << i; }
        }
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {
            pk[i] = 0;
            if ((amask >> i) & 1u) {
                u32 lp = pv[i] - bi.pbase, sp;
                if (CYCLIC) { u32 fs, fl; find_factor(a.fstart + bi.pbase, nf, bi.len, lp, fs, fl); { u32 o = lp - fs + a.h % fl; sp = fs + (o >= fl ? o - fl : o); } }
                else sp = lp + a.h;
                pk[i] = bi.pbase + sp;
            }
        }
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) if ((amask >> i) & 1u) pk[i] = a.rank[pk[i]];
    }
    const u32 cnt = __popc(amask);
    u64 tot;
    u64 incl = block_scan_incl((u64)cnt, 0ull, OpAdd(), s_warp, &tot);
    if (tid < 32) {
        u64 e = lb_exclusive(a.lb, tile, (td.flags & 1u) != 0, tot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u32 o = bi.pbase + (u32)s_excl + (u32)(incl - cnt);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) if ((amask >> i) & 1u) { u32 d = o + __popc(amask & ((1u << i) - 1u)); a.K2[d] = pk[i]; a.V2[d] = pv[i]; }
    if (tid == 0 && (td.flags & 2u)) a.active[td.block] = (u32)(s_excl + tot);
}

__global__ void __launch_bounds__(KOLM_THREADS) k_ls_key1(u32* __restrict__ K, const u32* __restrict__ V, const TileDesc* __restrict__ tiles,
                                                          const u32* __restrict__ rank) {
    TileDesc td = tiles[blockIdx.x];
    u32 v[KOLM_IPT];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { u32 x = i * KOLM_THREADS + threadIdx.x; v[i] = x < td.count ? V[td.start + x] : 0xffffffffu; }
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { u32 x = i * KOLM_THREADS + threadIdx.x; if (v[i] != 0xffffffffu) K[td.start + x] = rank[v[i]]; }
}

__global__ void k_round_end(u32* __restrict__ newcls, u32* __restrict__ done, const u32* __restrict__ active, int nblocks, int cyclic) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    if (cyclic && !done[b] && active[b] && newcls[b] == 0) done[b] = 1;   // partition stable: equal rotations remain
    newcls[b] = 0;
}

// ------------------------------------------------------------------------------------------------
// Local refinement rounds (Larsson–Sadakane form).  After a rerank the order `sa` is grouped: the records of an unsettled group
// are consecutive, and doubling the depth only permutes every group inside its own range by rank[succ_h(.)].  Groups are short
// (text at depth 12: a third of the records sit in groups of weighted mean size 38, 0.8 % in groups above 256), so instead of
// streaming the whole order to emit the unsettled records, sorting them globally by their group start (three radix passes over
// 8-byte records) and re-ranking, ONE CTA per static tile of the order
//   * reads grp[] (group start of every order index) for its 4096 indices + LR_GCAP look-ahead: a group belongs to the tile its
//     start lies in; a group longer than LR_GCAP ("big") is left to the global path (its members are flagged in F[]),
//   * fetches rank[succ_h(.)] of the unsettled records of its groups into shared memory,
//   * lets every record count, inside its group, the keys below its own and the equal keys before it: that is its new place,
//     its new group start and whether it is now alone — no sort, no scan, no barrier besides the one after the key fetch,
//   * writes the new order in place and the new group starts to nr[] (by order index).
// k_apply_local publishes nr -> grp / rank after the kernel boundary (every CTA must order against the same snapshot of the
// ranks).  No ticket, no look-back; tiles without unsettled records retire (live[] = 0) and cost one byte load afterwards.
// ------------------------------------------------------------------------------------------------
#define LR_GCAP 256
#define LR_CAP (KOLM_TILE + LR_GCAP)
#define LR_RPT (LR_CAP / KOLM_THREADS)                   // 17 order indices per thread
#define LR_FOREIGN 0xffffu
#define LR_NOHEAD 0xfffeu
#define LR_UNIFORM_MIN 24                                 // groups at least this long are tested for "all keys equal" before the quadratic count
#define LR_SMEM (2 * (LR_CAP + 8) + 2 * LR_CAP + 4 * LR_CAP + 4 * LR_CAP + 2 * LR_CAP + LR_CAP)

struct RefineArgs {
    u32* sa; const u32* rank; const u32* grp; u32* list; u32* list2; u32* lcount; u32* F; const TileDesc* tiles; const BlockInfo* binfo;
    const u32* fstart; const u32* nfac; const u32* done; u8* live; u32* lact; u32* newcls; u32* stats; u32 h;
    u32* gflag; u32 stamp;              // gflag[group start] == stamp: that big group's members do not all carry the same successor rank
    int big_uniform;                    // 1: fetch the keys of big groups' members too and let k_big_emit skip groups without a difference
};

template <bool CYCLIC>
__global__ void __launch_bounds__(KOLM_THREADS, 3) k_refine_local(RefineArgs a) {
    extern __shared__ __align__(16) u8 lr_smem[];
    u32* K2c = reinterpret_cast<u32*>(lr_smem);            // successor rank of compact record i
    u32* CP = K2c + LR_CAP;                                // its position
    u16* G16 = reinterpret_cast<u16*>(CP + LR_CAP);        // grp[] window minus t0 (LR_FOREIGN: the group starts before the tile)   [LR_CAP + 8]
    u16* E = G16 + LR_CAP + 8;                             // end (exclusive, window index) of the group that starts at this window index
    u16* CX = E + LR_CAP;                                  // window index of compact record i
    u8* COFF = reinterpret_cast<u8*>(CX + LR_CAP);         // its offset inside its group
    __shared__ u32 s_wcnt[NWARPS], s_wbig[NWARPS];
    __shared__ u32 s_g0, s_keyprev, s_nchg;
    __shared__ u32 s_flag[4];                              // [0] unsettled in the nominal range, [1] big in the nominal range, [2] straddling-in group is big, [3] changed records
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const u32 tile = blockIdx.x;
    if (!a.live[tile]) return;
    const TileDesc td = a.tiles[tile];
    const BlockInfo bi = a.binfo[td.block];
    if (a.done[td.block]) { if (tid == 0) a.live[tile] = 0; return; }
    const u32 nrec = bi.len, t0 = td.start - bi.pbase, cnt = td.count;
    const u32 navail = min((u32)LR_CAP, nrec - t0);        // order indices this CTA may refine
    const u32 nG = min((u32)LR_CAP + 1, nrec - t0);        // grp[] entries staged
    const u32* grp = a.grp + td.start;
    // window entries past the staged ones: "a head" right after the block's last record, "not a head" otherwise (LR_NOHEAD never
    // equals a window index or a group start), so the classify loop needs no bounds checks
    {
        constexpr int NI = (LR_CAP + 8 + KOLM_THREADS - 1) / KOLM_THREADS;
        u32 gv[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) { const u32 x = tid + i * KOLM_THREADS; gv[i] = x < nG ? grp[x] : 0u; }     // all loads in flight at once
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const u32 x = tid + i * KOLM_THREADS;
            u32 v = LR_NOHEAD;
            if (x < nG) v = gv[i] >= t0 ? gv[i] - t0 : LR_FOREIGN;
            else if (x == nG && t0 + nG == nrec) v = nG;
            if (x < LR_CAP + 8) G16[x] = (u16)v;
        }
    }
    if (tid == 0) {
        const u32 g0 = grp[0];
        s_flag[0] = 0; s_flag[1] = 0; s_flag[3] = 0; s_g0 = g0; s_nchg = 0;
        s_flag[2] = (g0 < t0 && g0 + LR_GCAP < nrec && a.grp[bi.pbase + g0 + LR_GCAP] == g0) ? 1u : 0u;
    }
    __syncthreads();
    const bool big_in = s_flag[2] != 0;
    // ---- classify: warp w looks at window indices [w*LR_RPT*32, (w+1)*LR_RPT*32), iteration k at 32 consecutive ones
    u32 procbits = 0, bigbits = 0, nuns = 0, nbig = 0, nproc = 0, nbigw = 0;
#pragma unroll
    for (int k = 0; k < LR_RPT; ++k) {
        const u32 x = w * (LR_RPT * 32) + k * 32 + lane;
        const u32 g = G16[x], gn = G16[x + 1];
        const bool ends = gn == x + 1;                     // my group ends with me
        const bool own = g < cnt;                          // its start lies in the nominal range (LR_FOREIGN / LR_NOHEAD are larger)
        if (ends && g < LR_CAP) { E[g] = (u16)(x + 1); if (own && x + 1 - g >= LR_UNIFORM_MIN) s_flag[3] = 1; }
        const bool single = ends && g == x;
        const bool big = own ? G16[g + LR_GCAP] == g : (g == LR_FOREIGN && big_in);
        const bool proc = own && !big && !single;
        const bool nb = big && x < cnt;                    // member of a big group in the nominal range
        if (x < cnt) nuns += single ? 0u : 1u;
        if (nb) { ++nbig; bigbits |= 1u << k; }
        if (proc) procbits |= 1u << k;
        nproc += __popc(__ballot_sync(FULL, proc));        // warp totals (the same in every lane)
        nbigw += __popc(__ballot_sync(FULL, nb));
    }
    for (int o = 16; o > 0; o >>= 1) { nuns += __shfl_xor_sync(FULL, nuns, o); nbig += __shfl_xor_sync(FULL, nbig, o); }
    if (lane == 0) { s_wcnt[w] = nproc; s_wbig[w] = nbigw; if (nuns) atomicAdd(&s_flag[0], nuns); if (nbig) atomicAdd(&s_flag[1], nbig); }
    u32 pp[LR_RPT];                                        // positions of my records: in flight across the barrier
    const u32 loadbits = a.big_uniform ? (procbits | bigbits) : procbits;
#pragma unroll
    for (int k = 0; k < LR_RPT; ++k) pp[k] = ((loadbits >> k) & 1u) ? a.sa[td.start + w * (LR_RPT * 32) + k * 32 + lane] : 0u;
    __syncthreads();
    const u32 tot_uns = s_flag[0], tot_big = s_flag[1];
    if (!tot_uns) { if (tid == 0) { a.live[tile] = 0; a.lcount[tile] = 0; } return; }
    if (tid == 0) {
        atomicAdd(a.lact + td.block, tot_uns); atomicAdd(a.stats + 6, tot_uns);
        if (tot_big) { atomicAdd(a.stats + 5, tot_big); atomicAdd(a.stats + 4, tot_big); }    // big members stay unsettled unless the global path splits them
        a.live[tile] = tot_big ? 3 : 1;                    // bit 1: k_big_emit has work here
    }
    if (tot_big) {                                         // only such tiles are read by k_big_emit
#pragma unroll
        for (int k = 0; k < LR_RPT; ++k) {
            const u32 x = w * (LR_RPT * 32) + k * 32 + lane;
            if (x < cnt) a.F[td.start + x] = ((bigbits >> k) & 1u) ? 0u : 0x80000000u;
        }
    }
    u32 wbase = 0, m = 0, bbase = 0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) { const u32 c_ = s_wcnt[i], b_ = s_wbig[i]; if ((u32)i < w) { wbase += c_; bbase += b_; } m += c_; }
    const u32 mall = m + (a.big_uniform ? tot_big : 0u);
    if (!mall) { if (tid == 0) a.lcount[tile] = 0; return; }     // only a foreign small group here
    // ---- compact: my records at [0, m), then the nominal members of big groups at [m, mall) (only their keys are looked at)
    {
        u32 run = wbase, brun = m + bbase;
#pragma unroll
        for (int k = 0; k < LR_RPT; ++k) {
            const bool proc = (procbits >> k) & 1u, nb = a.big_uniform && ((bigbits >> k) & 1u);
            const u32 pm = __ballot_sync(FULL, proc), bm = __ballot_sync(FULL, nb);
            const u32 x = w * (LR_RPT * 32) + k * 32 + lane;
            if (proc) {
                const u32 slot = run + __popc(pm & lanemask_lt());
                CX[slot] = (u16)x; COFF[slot] = (u8)(x - G16[x]); CP[slot] = pp[k];
            }
            if (nb) {
                const u32 slot = brun + __popc(bm & lanemask_lt());
                CX[slot] = (u16)x; CP[slot] = pp[k];
            }
            run += __popc(pm); brun += __popc(bm);
        }
    }
    __syncthreads();
    // ---- successor ranks, four gathers in flight per thread
    {
        const u32 nf = CYCLIC ? a.nfac[td.block] : 0;
        const u32* fst = a.fstart + bi.pbase;
        if (tid == KOLM_THREADS - 1 && big_in && t0 > 0 && a.big_uniform) {         // the key of the order index before the tile: left neighbour of a big group's member at x = 0
            u32 lp = a.sa[td.start - 1] - bi.pbase, sp;
            if (CYCLIC) { u32 fs, fl; find_factor(fst, nf, bi.len, lp, fs, fl); { u32 o = lp - fs + a.h % fl; sp = fs + (o >= fl ? o - fl : o); } }
            else sp = lp + a.h;
            s_keyprev = a.rank[bi.pbase + sp];
        }
        for (u32 i0 = tid; i0 < mall; i0 += 4 * KOLM_THREADS) {
            u32 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const u32 i = i0 + u * KOLM_THREADS;
                q[u] = 0xffffffffu;
                if (i < mall) {
                    u32 lp = CP[i] - bi.pbase, sp;
                    if (CYCLIC) { u32 fs, fl; find_factor(fst, nf, bi.len, lp, fs, fl); { u32 o = lp - fs + a.h % fl; sp = fs + (o >= fl ? o - fl : o); } }
                    else sp = lp + a.h;
                    q[u] = bi.pbase + sp;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (q[u] != 0xffffffffu) q[u] = a.rank[q[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const u32 i = i0 + u * KOLM_THREADS; if (i < mall) K2c[i] = q[u]; }
        }
    }
    __syncthreads();
    // ---- groups whose members all carry the same key stay as they are (periodic data: most groups, for many rounds).  Small
    //      groups: bit 15 of E[group start]; big groups: gflag[group start] = stamp, read by k_big_emit
    const bool utest = s_flag[3] != 0;                     // (read after the barriers that follow its writers)
    for (u32 i = utest ? tid : m + tid; i < mall; i += KOLM_THREADS) {
        if (i < m) {
            const u32 off = COFF[i];
            if (off && K2c[i] != K2c[i - off]) { const u32 gs = (u32)CX[i] - off; atomicOr(reinterpret_cast<u32*>(E) + (gs >> 1), 0x8000u << (16 * (gs & 1u))); }
        } else {
            const u32 x = CX[i], g = G16[x];
            if (g != x) {                                  // not the group's first member: slot i - 1 holds order index x - 1 (same group) unless x = 0
                const u32 left = x ? K2c[i - 1] : s_keyprev;
                if (K2c[i] != left) a.gflag[bi.pbase + (g == LR_FOREIGN ? s_g0 : t0 + g)] = a.stamp;
            }
        }
    }
    __syncthreads();
    if (!m) { if (tid == 0) a.lcount[tile] = 0; return; }
    // ---- every record finds its place inside its group: keys below mine, equal keys before me
    u32 nnew = 0, nsurv = 0;
    for (u32 i = tid; i < m; i += KOLM_THREADS) {
        const u32 x = CX[i], off = COFF[i], cs = i - off, gs = x - off, eg = E[gs], n = (eg & 0x7fffu) - gs, mine = K2c[i];
        if (utest && !(eg >> 15)) { ++nsurv; continue; }   // all keys of the group equal: nothing moves
        // keys are below 2^30: (k - mine) >> 31 counts "k < mine", (k - mine - 1) >> 31 counts "k <= mine" — two adds per key
        // and count; the keys before me and the ones from me on are counted apart, their "<=" minus "<" are the equal ones
        const u32* kp = K2c + cs;
        const u32 m1 = mine + 1u;
        u32 ltb = 0, leb = 0, lta = 0, lea = 0, j = 0;
        if (n <= 4) {                                      // most groups of text: no loops (slots past the group are masked out)
            const u32 k0 = kp[0], k1 = kp[1], k2 = n > 2 ? kp[2] : 0x7fffffffu, k3 = n > 3 ? kp[3] : 0x7fffffffu;
            const u32 l0 = (k0 - mine) >> 31, l1 = (k1 - mine) >> 31, l2 = (k2 - mine) >> 31, l3 = (k3 - mine) >> 31;
            const u32 e0 = (k0 - m1) >> 31, e1 = (k1 - m1) >> 31, e2 = (k2 - m1) >> 31, e3 = (k3 - m1) >> 31;
            lta = l0 + l1 + l2 + l3; lea = e0 + e1 + e2 + e3;
            ltb = (off > 0 ? l0 : 0u) + (off > 1 ? l1 : 0u) + (off > 2 ? l2 : 0u);
            leb = (off > 0 ? e0 : 0u) + (off > 1 ? e1 : 0u) + (off > 2 ? e2 : 0u);
            lta -= ltb; lea -= leb;
            j = n;
        }
        for (; j + 4 <= off && j < n; j += 4) {
            const u32 k0 = kp[j], k1 = kp[j + 1], k2 = kp[j + 2], k3 = kp[j + 3];
            ltb += ((k0 - mine) >> 31) + ((k1 - mine) >> 31) + ((k2 - mine) >> 31) + ((k3 - mine) >> 31);
            leb += ((k0 - m1) >> 31) + ((k1 - m1) >> 31) + ((k2 - m1) >> 31) + ((k3 - m1) >> 31);
        }
        for (; j < off && j < n; ++j) { const u32 kj = kp[j]; ltb += (kj - mine) >> 31; leb += (kj - m1) >> 31; }
        for (; j + 4 <= n; j += 4) {
            const u32 k0 = kp[j], k1 = kp[j + 1], k2 = kp[j + 2], k3 = kp[j + 3];
            lta += ((k0 - mine) >> 31) + ((k1 - mine) >> 31) + ((k2 - mine) >> 31) + ((k3 - mine) >> 31);
            lea += ((k0 - m1) >> 31) + ((k1 - m1) >> 31) + ((k2 - m1) >> 31) + ((k3 - m1) >> 31);
        }
        for (; j < n; ++j) { const u32 kj = kp[j]; lta += (kj - mine) >> 31; lea += (kj - m1) >> 31; }
        const u32 lt = ltb + lta, eqb = leb - ltb, eq = eqb + (lea - lta);
        const u32 np = gs + lt + eqb;
        a.sa[td.start + np] = CP[i];
        if (lt) {                                          // group start moved: k_apply_local publishes it
            const u32 idx = atomicAdd(&s_nchg, 1u);
            const u32 e = np | ((gs + lt) << 16);
            if (idx < KOLM_TILE) a.list[td.start + idx] = e; else a.list2[td.start + idx - KOLM_TILE] = e;
        }
        nnew += (eqb == 0 && lt > 0) ? 1u : 0u;           // a new group start that was none before
        nsurv += eq > 1 ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) { nnew += __shfl_xor_sync(FULL, nnew, o); nsurv += __shfl_xor_sync(FULL, nsurv, o); }
    if (lane == 0) { if (nnew) atomicAdd(a.newcls + td.block, nnew); if (nsurv) atomicAdd(a.stats + 4, nsurv); }
    __syncthreads();
    if (tid == 0) a.lcount[tile] = s_nchg;
}

// records of the tiles k_refine_local marked, appended to the block's record list as (rank[succ_h(v)], v) in any order (the LS
// path sorts by both keys); runs after k_apply_local so that all keys of the global path come from one snapshot.
//   ALL = false: members of big groups (F == 0) in tiles with live == 3;  ALL = true: every unsettled record of every live tile
//   (hand-over to the compacted LS rounds once few records are left).
template <bool CYCLIC, bool ALL>
__global__ void __launch_bounds__(KOLM_THREADS) k_big_emit(const u32* __restrict__ sa, const u32* __restrict__ rank, const u32* __restrict__ F, const u32* __restrict__ grp,
                                                           const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo, const u32* __restrict__ fstart,
                                                           const u32* __restrict__ nfac, const u32* __restrict__ done, const u8* __restrict__ live, u32* __restrict__ active,
                                                           u32* __restrict__ K2, u32* __restrict__ V2, u32 h, const u32* __restrict__ gflag, u32 stamp) {
    __shared__ u32 s_w[NWARPS];
    __shared__ u32 s_base;
    const u8 lv = live[blockIdx.x];
    if (ALL ? lv == 0 : lv != 3) return;
    const TileDesc td = tiles[blockIdx.x];
    if (done[td.block]) return;
    const BlockInfo bi = binfo[td.block];
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const u32 t0 = td.start - bi.pbase;
    if (ALL) {
        // sixteen consecutive order indices per thread: four 16-byte loads of grp[] (+ the next thread's first value), the
        // unsettled ones read off in registers; any order inside the block's list will do (the LS path sorts by both keys)
        const u32 x0 = tid * KOLM_IPT;
        u32 g[KOLM_IPT + 1];
        if (x0 + KOLM_IPT <= td.count) {
            const uint4* gp = reinterpret_cast<const uint4*>(grp + td.start + x0);
#pragma unroll
            for (int q = 0; q < KOLM_IPT / 4; ++q) { const uint4 v = gp[q]; g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w; }
        } else {
#pragma unroll
            for (int k = 0; k < KOLM_IPT; ++k) g[k] = x0 + k < td.count ? grp[td.start + x0 + k] : 0u;
        }
        g[KOLM_IPT] = (x0 + KOLM_IPT < td.count || t0 + x0 + KOLM_IPT < bi.len) && x0 + KOLM_IPT <= td.count ? grp[td.start + x0 + KOLM_IPT] : 0xffffffffu;
        u32 bits = 0;
#pragma unroll
        for (int k = 0; k < KOLM_IPT; ++k) {
            const u32 j = t0 + x0 + k;                           // block-local order index
            const bool single = g[k] == j && (j + 1 == bi.len || g[k + 1] == j + 1);
            if (x0 + k < td.count && !single) bits |= 1u << k;
        }
        const u32 c = __popc(bits);
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(FULL, incl, o); if (lane >= (u32)o) incl += t; }
        if (lane == 31) s_w[w] = incl;
        __syncthreads();
        if (tid == 0) { u32 t = 0; for (int i = 0; i < NWARPS; ++i) { const u32 c_ = s_w[i]; s_w[i] = t; t += c_; } s_base = t ? atomicAdd(active + td.block, t) : 0u; }
        __syncthreads();
        u32 d = bi.pbase + s_base + s_w[w] + incl - c;
        const u32 nf = CYCLIC ? nfac[td.block] : 0;
        u32 b = bits;
        while (b) {
            const u32 k = __ffs(b) - 1; b &= b - 1;
            const u32 p = sa[td.start + x0 + k];
            u32 lp = p - bi.pbase, sp;
            if (CYCLIC) { u32 fs, fl; find_factor(fstart + bi.pbase, nf, bi.len, lp, fs, fl); { u32 o = lp - fs + h % fl; sp = fs + (o >= fl ? o - fl : o); } }
            else sp = lp + h;
            K2[d] = rank[bi.pbase + sp]; V2[d] = p; ++d;
        }
        return;
    }
    u32 bits = 0, mycnt = 0;
#pragma unroll
    for (int k = 0; k < KOLM_IPT; ++k) {
        const u32 x = w * (KOLM_IPT * 32) + k * 32 + lane;
        bool take = false;
        if (x < td.count) {
            if (ALL) {
                const u32 g = grp[td.start + x];
                take = !(g == t0 + x && (t0 + x + 1 == bi.len || grp[td.start + x + 1] == t0 + x + 1));
            } else take = F[td.start + x] == 0u && (!gflag || gflag[bi.pbase + grp[td.start + x]] == stamp);   // gflag: only big groups some of whose keys differ
        }
        if (take) bits |= 1u << k;
        mycnt += __popc(__ballot_sync(FULL, take));        // warp total so far (same in every lane)
    }
    if (lane == 0) s_w[w] = mycnt;
    __syncthreads();
    if (tid == 0) { u32 t = 0; for (int i = 0; i < NWARPS; ++i) { const u32 c_ = s_w[i]; s_w[i] = t; t += c_; } s_base = t ? atomicAdd(active + td.block, t) : 0u; }
    __syncthreads();
    u32 run = bi.pbase + s_base + s_w[w];
    const u32 nf = CYCLIC ? nfac[td.block] : 0;
#pragma unroll 1
    for (int k = 0; k < KOLM_IPT; ++k) {
        const bool take = (bits >> k) & 1u;
        const u32 bm = __ballot_sync(FULL, take);
        if (take) {
            const u32 x = w * (KOLM_IPT * 32) + k * 32 + lane;
            const u32 p = sa[td.start + x];
            u32 lp = p - bi.pbase, sp;
            if (CYCLIC) { u32 fs, fl; find_factor(fstart + bi.pbase, nf, bi.len, lp, fs, fl); { u32 o = lp - fs + h % fl; sp = fs + (o >= fl ? o - fl : o); } }
            else sp = lp + h;
            const u32 d = run + __popc(bm & lanemask_lt());
            K2[d] = rank[bi.pbase + sp]; V2[d] = p;
        }
        run += __popc(bm);
    }
}

// publish a local round: grp / rank take the new group starts of the records k_refine_local listed
__global__ void __launch_bounds__(KOLM_THREADS) k_apply_local(const u32* __restrict__ sa, const u32* __restrict__ list, const u32* __restrict__ list2,
                                                              const u32* __restrict__ lcount, u32* __restrict__ grp, u32* __restrict__ rank,
                                                              const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo, const u8* __restrict__ live) {
    if (!live[blockIdx.x]) return;
    const u32 n = lcount[blockIdx.x];
    if (!n) return;
    const TileDesc td = tiles[blockIdx.x];
    const u32 t0 = td.start - binfo[td.block].pbase;
    for (u32 i0 = threadIdx.x; i0 < n; i0 += 8 * KOLM_THREADS) {
        u32 e[8], v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const u32 i = i0 + u * KOLM_THREADS; e[u] = i < n ? (i < KOLM_TILE ? list[td.start + i] : list2[td.start + i - KOLM_TILE]) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = e[u] != 0xffffffffu ? sa[td.start + (e[u] & 0xffffu)] : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (e[u] != 0xffffffffu) { const u32 g = t0 + (e[u] >> 16); grp[td.start + (e[u] & 0xffffu)] = g; rank[v[u]] = g; }
    }
}

// local rounds: a block whose unsettled groups produced no new class in a whole round is stable (equal rotations)
__global__ void k_round_end_local(u32* __restrict__ newcls, u32* __restrict__ done, u32* __restrict__ lact, int nblocks, int cyclic) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    if (cyclic && !done[b] && lact[b] && newcls[b] == 0) done[b] = 1;
    newcls[b] = 0; lact[b] = 0;
}

// ------------------------------------------------------------------------------------------------
// Lyndon starts = strict prefix minima of the ISA (a1).  Emits the compacted factor-start list per
// block (block-local positions), the factor count, and optionally one flag byte per position.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KOLM_THREADS) k_lyndon(const u32* __restrict__ rank, const TileDesc* __restrict__ tiles,
                                                         const BlockInfo* __restrict__ binfo, u64* lb, u32* __restrict__ fstart,
                                                         u32* __restrict__ nfac, u8* __restrict__ flags_out) {
    __shared__ u64 s_warp[NWARPS];
    __shared__ u64 s_last[NWARPS];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u64 IDENT = 0x7fffffffull << 31;
    u32 rk[KOLM_IPT];
    u32 mn = 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        rk[i] = r < td.count ? rank[td.start + r] : 0x7fffffffu;
        mn = min(mn, rk[i]);
    }
    // pass 1: prefix minimum only (count lane 0)
    u64 tot;
    u64 incl = block_scan_incl((u64)mn << 31, IDENT, OpMinAdd(), s_warp, &tot);
    u64 prev = __shfl_up_sync(FULL, incl, 1);
    if ((tid & 31) == 31) s_last[tid >> 5] = incl;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : IDENT;
    // tile-local start candidates need the exclusive min from previous tiles first: do the look-back
    // on (min, 0), then a second block scan for the count and a second look-back for the offsets.
    if (tid < 32) {
        u64 e = lb_exclusive(lb, tile, (td.flags & 1u) != 0, tot, IDENT, OpMinAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u32 run = min((u32)(s_excl >> 31), (u32)(prev >> 31));
    u32 startbits = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        if (r < td.count && rk[i] < run) { startbits |= 1u << i; ++cnt; run = rk[i]; }
    }
    __syncthreads();
    // second scan: counts.  A separate look-back array region (lb2 = lb + 8 + gridDim.x) keeps the two scans apart.
    u64* lb2 = lb + 8 + gridDim.x;
    u64 ctot;
    u64 cincl = block_scan_incl((u64)cnt, 0ull, OpAdd(), s_warp, &ctot);
    if (tid < 32) {
        u64 e = lb_exclusive(lb2 - 8, tile, (td.flags & 1u) != 0, ctot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u32 o = bi.pbase + (u32)s_excl + (u32)(cincl - cnt);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        if (r < td.count) {
            bool st = (startbits >> i) & 1u;
            if (st) fstart[o++] = td.start + r - bi.pbase;
            if (flags_out) flags_out[bi.ioff + (td.start + r - bi.pbase)] = st ? 1 : 0;
        }
    }
    if (tid == 0 && (td.flags & 2u)) nfac[td.block] = (u32)(s_excl + ctot);
}


// ------------------------------------------------------------------------------------------------
// Lyndon fast path.  Factor starts are the strict prefix minima of the suffix order, so a position can
// only start a factor if its 6-byte prefix (+ length code) is <= every earlier one.  k_lyn_cand finds
// those candidates with one look-back min-scan over 51-bit keys; k_lyn_resolve (one warp per block) walks
// them in order and settles ties against the current champion by direct suffix comparison.  Text-like data
// has a handful of candidates per block; degenerate data (long runs / periods) exceeds the work budget and
// the whole batch takes the robust ISA path (plain suffix sort + k_lyndon) instead.  Both paths are exact.
// ------------------------------------------------------------------------------------------------
#define LYN_MAX_CAND 4096u
#define LYN_MAX_ITERS 8192u
#define LYN_KEY_INF ((1ull << 52) - 1)

__device__ __forceinline__ u64 lyn_key(const u8* __restrict__ src, u32 lp, u32 len) {
    u32 rem = len - lp;
    u64 k = 0;
#pragma unroll
    for (int t = 0; t < 6; ++t) k = (k << 8) | (u64)((u32)t < rem ? src[lp + t] : 0);
    return (k << 3) | (u64)(rem < 7 ? rem : 7);
}

__global__ void __launch_bounds__(KOLM_THREADS) k_lyn_cand(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u64* lb, u32* __restrict__ cand,
                                                           u32* __restrict__ ncand) {
    __shared__ u64 s_warp[NWARPS];
    __shared__ u64 s_last[NWARPS];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff;
    const u32 t0 = td.start - bi.pbase;
    u64 key[KOLM_IPT];
    u64 mn = LYN_KEY_INF;
    {
        // my 16 positions and the 5 bytes after them: one 16-byte and one 8-byte load when the window lies inside the block and
        // is aligned (byte loads at a 16-byte lane stride cost 4 L1 wavefronts each, 96 of them per thread), funnel shifts
        // cut the 6-byte prefixes out of the registers; threads at the end of a block or on unaligned input build them bytewise
        const u32 r0 = tid * KOLM_IPT;
        const u8* p = src + t0 + r0;
        if (KOLM_IPT == 16 && r0 + KOLM_IPT <= td.count && t0 + r0 + 24u <= bi.len && ((uintptr_t)p & 15) == 0) {
            const uint4 a = *reinterpret_cast<const uint4*>(p);
            const uint2 b = *reinterpret_cast<const uint2*>(p + 16);
            const u32 w[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
#pragma unroll
            for (int i = 0; i < KOLM_IPT; ++i) {
                const int q = i >> 2, sh = 8 * (i & 3);
                const u32 x0 = __funnelshift_r(w[q], w[q + 1], sh), x1 = __funnelshift_r(w[q + 1], w[(q + 2) < 6 ? q + 2 : 5], sh);
                const u64 k = ((u64)__byte_perm(x0, 0, 0x0123) << 16) | (u64)__byte_perm(x1, 0, 0x4401);
                key[i] = (k << 3) | 7ull;                    // at least 9 bytes remain: length code 7
                mn = key[i] < mn ? key[i] : mn;
            }
        } else {
#pragma unroll
            for (int i = 0; i < KOLM_IPT; ++i) {
                u32 r = r0 + i;
                key[i] = r < td.count ? lyn_key(src, t0 + r, bi.len) : LYN_KEY_INF;
                mn = key[i] < mn ? key[i] : mn;
            }
        }
    }
    struct OpMin { __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a < b ? a : b; } };
    u64 tot;
    u64 incl = block_scan_incl(mn, LYN_KEY_INF, OpMin(), s_warp, &tot);
    u64 prev = __shfl_up_sync(FULL, incl, 1);
    if ((tid & 31) == 31) s_last[tid >> 5] = incl;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : LYN_KEY_INF;
    if (tid < 32) {
        u64 e = lb_exclusive(lb, tile, (td.flags & 1u) != 0, tot, LYN_KEY_INF, OpMin());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u64 run = s_excl < prev ? s_excl : prev;
    u32 cmask = 0, smask = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if (key[i] != LYN_KEY_INF || tid * KOLM_IPT + i < td.count) {
            if (tid * KOLM_IPT + i < td.count) {
                if (key[i] < run) { cmask |= 1u << i; smask |= 1u << i; run = key[i]; }
                else if (key[i] == run) cmask |= 1u << i;
            }
        }
    }
    __syncthreads();
    u64* lb2 = lb + gridDim.x;
    const u32 cnt = __popc(cmask);
    u64 ctot;
    u64 cincl = block_scan_incl((u64)cnt, 0ull, OpAdd(), s_warp, &ctot);
    if (tid < 32) {
        u64 e = lb_exclusive(lb2, tile, (td.flags & 1u) != 0, ctot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u32 o = (u32)s_excl + (u32)(cincl - cnt);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if ((cmask >> i) & 1u) {
            if (o < LYN_MAX_CAND + 1) cand[bi.pbase + o] = (t0 + tid * KOLM_IPT + i) | (((smask >> i) & 1u) << 31);
            ++o;
        }
    }
    if (tid == 0 && (td.flags & 2u)) ncand[td.block] = (u32)(s_excl + ctot);
}

// one warp per block
__global__ void __launch_bounds__(128) k_lyn_resolve(const u8* __restrict__ in, const BlockInfo* __restrict__ binfo, const u32* __restrict__ cand,
                                                     const u32* __restrict__ ncand, u32* __restrict__ fstart, u32* __restrict__ nfac,
                                                     u8* __restrict__ flags_out, u32* __restrict__ fallback, u32* __restrict__ blockfail, int nblocks) {
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const BlockInfo bi = binfo[b];
    if (lane == 0) blockfail[b] = 0;
    if (bi.len == 0) { if (lane == 0) nfac[b] = 0; return; }
    const u32 nc = ncand[b];
    if (nc > LYN_MAX_CAND || (bi.len > LYN_MAX_CAND && nc > bi.len / 8)) { if (lane == 0) { blockfail[b] = 1; atomicExch(fallback, 1u); } return; }
    const u8* src = in + bi.ioff;
    u32 champ = 0, nf = 0, iters = 0;
    for (u32 k = 0; k < nc; ++k) {
        u32 cv = cand[bi.pbase + k];
        u32 pos = cv & 0x7fffffffu;
        bool win = (cv >> 31) != 0;
        if (!win) {
            // tie on the 6-byte key (both suffixes have >= 7 bytes): compare from offset 6, 32 bytes per step
            u32 o = 6;
            for (;;) {
                if (++iters > LYN_MAX_ITERS) { if (lane == 0) { blockfail[b] = 1; atomicExch(fallback, 1u); } return; }
                u32 x = o + lane;
                bool inb = pos + x < bi.len;                 // pos > champ, so pos's suffix ends first
                u32 a = inb ? src[pos + x] : 0, c = inb ? src[champ + x] : 0;
                u32 ne = __ballot_sync(FULL, !inb || a != c);
                if (ne) {
                    u32 l = __ffs(ne) - 1;
                    u32 aa = __shfl_sync(FULL, a, l), cc = __shfl_sync(FULL, c, l);
                    bool ended = !__shfl_sync(FULL, (u32)inb, l);
                    win = ended || aa < cc;                  // a proper prefix is the smaller suffix
                    break;
                }
                o += 32;
            }
        }
        if (win) {
            champ = pos;
            if (lane == 0) fstart[bi.pbase + nf] = pos;
            ++nf;
        }
    }
    if (lane == 0) nfac[b] = nf;
    if (flags_out) for (u32 t = lane; t < nf; t += 32) flags_out[bi.ioff + fstart[bi.pbase + t]] = 1;   // only once the block succeeded
}

// Blocks the candidate path gave up on (long runs / periods): Duval's walk (kolm_final.py:200-225) by one warp, with the
// run of equal comparisons s[k]==s[j] taken 128 bytes at a time.  Degenerate data is exactly where those runs are long, so
// the walk needs few steps there; an iteration cap sends pathological blocks to the ISA path.
#define DUVAL_MAX_BLOCK_DEFAULT (64u << 20)   // KOLM_DUVAL_MAX_MIB: larger failed blocks send the batch to the ISA path
__global__ void __launch_bounds__(128) k_lyn_duval(const u8* __restrict__ in, const BlockInfo* __restrict__ binfo, const u32* __restrict__ blockfail,
                                                   u32* __restrict__ fstart, u32* __restrict__ nfac, u8* __restrict__ flags_out,
                                                   u32* __restrict__ fallback, int nblocks, u32 max_block) {
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblocks || !blockfail[b]) return;
    const BlockInfo bi = binfo[b];
    const u8* s = in + bi.ioff;
    const u32 n = bi.len;
    // one warp walks the block serially: only worth it for blocks it can finish in a few ms; larger ones go to the ISA path
    if (n > max_block) { if (lane == 0) atomicExch(fallback, 2u); return; }
    const u32 max_iters = (1u << 18) + n / 16, max_siters = (1u << 20) + n / 8;   // work caps (about 0.1-0.3 s worst case): beyond them the ISA path decides
    u32 i = 0, nf = 0, iters = 0, siters = 0;
    while (i < n) {
        u32 j = i + 1, k = i;
        const u32 si = s[i];
        u32 run = 0, adv = 0;                                // consecutive equal comparisons / consecutive plain advances
        u32 jw = 0, jwbase = 0xffffffffu;                    // 4-byte window of s around j (one load per 4 scalar steps)
        for (;;) {
            if (j >= n) break;
            if (run >= 16) {
                // long common prefix of s[k..] and s[j..]: skip it 128 bytes per step with the whole warp
                for (;;) {
                    if (++iters > max_iters) { if (lane == 0) atomicExch(fallback, 2u); return; }
                    u32 x = lane * 4, m = 4;
#pragma unroll
                    for (int t = 3; t >= 0; --t) { u32 jj = j + x + t; if (jj >= n || s[k + x + t] != s[jj]) m = t; }
                    u32 bal = __ballot_sync(FULL, m < 4);
                    if (bal) { u32 l = __ffs(bal) - 1; u32 adv = l * 4 + __shfl_sync(FULL, m, l); k += adv; j += adv; break; }
                    k += 128; j += 128;
                }
                run = 0;
                if (j >= n) break;
            }
            if (k == i && adv >= 8) {                       // only after a streak of plain advances: dense stops would cost a step each
                // fresh comparison against the factor's first byte: every byte > s[i] only advances j, so the warp skips to the
                // first byte <= s[i], 512 bytes (32 aligned 16-byte loads) per step
                const u8* sa16 = reinterpret_cast<const u8*>(reinterpret_cast<uintptr_t>(s + j) & ~(uintptr_t)15);
                const i64 rel0 = (i64)(sa16 - s);                      // block offset of the aligned base (may be negative by < 16)
                const u32 pat = si * 0x01010101u;
                bool found = false;
                for (i64 base = rel0;; base += 512) {
                    if (++iters > max_iters) { if (lane == 0) atomicExch(fallback, 2u); return; }
                    const i64 o = base + lane * 16;
                    u32 first = 16;
                    if (o < (i64)n && o + 16 > (i64)j) {
                        // the aligned 16 bytes may start before the block or end after it: stay inside [0, n) of the batch buffer's
                        // allocation by reading bytes when the vector would cross either end
                        u32 w[4];
                        if (o >= 0 && o + 16 <= (i64)n) { const uint4 q = *reinterpret_cast<const uint4*>(s + o); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
                        else { w[0] = w[1] = w[2] = w[3] = 0xffffffffu; for (int t = 0; t < 16; ++t) { i64 x = o + t; if (x >= 0 && x < (i64)n) { w[t >> 2] = (w[t >> 2] & ~(0xFFu << (8 * (t & 3)))) | ((u32)s[x] << (8 * (t & 3))); } } }
#pragma unroll
                        for (int q4 = 3; q4 >= 0; --q4) {
                            u32 le = __vcmpleu4(w[q4], pat);                       // 0xFF in every byte <= s[i]
#pragma unroll
                            for (int t = 3; t >= 0; --t) { const i64 x = o + q4 * 4 + t; if (((le >> (8 * t)) & 1u) && x >= (i64)j && x < (i64)n) first = q4 * 4 + t; }
                        }
                    }
                    const u32 bal = __ballot_sync(FULL, first < 16);
                    if (bal) { const u32 l = __ffs(bal) - 1; j = (u32)(base + l * 16 + __shfl_sync(FULL, first, l)); found = true; break; }
                    if (base + 512 >= (i64)n) break;
                }
                if (!found) { j = n; break; }
                adv = 0;
            }
            // scalar step (all lanes redundantly: uniform control flow, broadcast loads)
            if (++siters > max_siters) { if (lane == 0) atomicExch(fallback, 2u); return; }
            if ((j & ~3u) != jwbase) { jwbase = j & ~3u; jw = 0; for (u32 t = 0; t < 4 && jwbase + t < n; ++t) jw |= (u32)s[jwbase + t] << (8 * t); }
            const u32 cj = (jw >> (8 * (j & 3))) & 0xFF;
            const u32 ck = (k == i) ? si : (u32)s[k];
            if (ck == cj) { ++k; ++j; ++run; adv = 0; continue; }
            run = 0;
            if (ck < cj) { k = i; ++j; ++adv; } else break;
        }
        const u32 p = j - k;
        const u32 cnt = (k - i) / p + 1;                     // while i <= k: emit i; i += p
        for (u32 t = lane; t < cnt; t += 32) { u32 st = i + t * p; fstart[bi.pbase + nf + t] = st; if (flags_out) flags_out[bi.ioff + st] = 1; }
        iters += cnt >> 5;
        nf += cnt; i += cnt * p;
    }
    if (lane == 0) nfac[b] = nf;
}

// out[j] = byte preceding rotation sa[j] inside its factor  (kolm_final.py:321).  Three kernels: prevb[p] = in[p - 1] for every
// position (streaming; indexed like sa's values, i.e. by padded position), the factor starts patched to their factor's last byte
// (one thread per factor), then the emit is one byte gather per order index — no factor search in the 268 M-record kernel.
__global__ void __launch_bounds__(KOLM_THREADS) k_prev_bytes(const u8* __restrict__ in, u8* __restrict__ prevb, const TileDesc* __restrict__ tiles,
                                                             const BlockInfo* __restrict__ binfo) {
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase;
    const u8* src = in + bi.ioff + t0;                     // src[x - 1] is the predecessor of tile index x (x = 0 at t0 = 0: patched below)
    u8* dst = prevb + td.start;                            // 128-byte aligned
    const bool al = (((uintptr_t)src) & 3) == 0;
    for (u32 x = threadIdx.x * 16; x < td.count; x += KOLM_THREADS * 16) {
        u32 w[4] = {0, 0, 0, 0};
        const u32 nb = min(16u, td.count - x);
        if (al && nb == 16) {                               // five aligned words in[x-4 .. x+15], shifted down by three bytes
            const u32* sw = reinterpret_cast<const u32*>(src + x);
            u32 v[5];
            v[0] = (t0 + x) ? sw[-1] : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i + 1] = sw[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = __funnelshift_r(v[i], v[i + 1], 24);
        } else
        for (u32 i = 0; i < nb; ++i) { const u32 q = x + i; const u32 b = (t0 + q) ? src[(int)q - 1] : 0u; w[i >> 2] |= b << (8 * (i & 3)); }
        if (nb == 16) *reinterpret_cast<uint4*>(dst + x) = make_uint4(w[0], w[1], w[2], w[3]);
        else for (u32 i = 0; i < nb; ++i) dst[x + i] = (u8)(w[i >> 2] >> (8 * (i & 3)));
    }
}

__global__ void __launch_bounds__(256) k_prev_patch(const u8* __restrict__ in, u8* __restrict__ prevb, const BlockInfo* __restrict__ binfo,
                                                    const u32* __restrict__ fstart, const u32* __restrict__ nfac, int nblocks) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const BlockInfo bi = binfo[b];
        if (!bi.len) continue;                             // an empty block has no tiles: nothing wrote its factor count
        const u32 nf = nfac[b];
        const u32* fst = fstart + bi.pbase;
        const u8* src = in + bi.ioff;
        for (u32 f = threadIdx.x; f < nf; f += blockDim.x) {
            const u32 fs = fst[f], fe = f + 1 < nf ? fst[f + 1] : bi.len;
            prevb[bi.pbase + fs] = src[fe - 1];
        }
    }
}

__global__ void __launch_bounds__(KOLM_THREADS) k_bbwt_emit(const u8* __restrict__ prevb, u8* __restrict__ out, const u32* __restrict__ sa,
                                                            const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo) {
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    // sixteen consecutive outputs per thread: four 16-byte loads of the order, sixteen byte gathers in flight, one 16-byte store
    // (td.start is a multiple of 32 elements, so the order is 16-byte aligned; the output is when the block is)
    u8* const dst = out + bi.ioff + (td.start - bi.pbase);
    const u32 n16 = (((uintptr_t)dst & 15) == 0) ? (td.count & ~15u) : 0u;
    for (u32 x = threadIdx.x * 16; x < n16; x += KOLM_THREADS * 16) {
        const uint4* sp = reinterpret_cast<const uint4*>(sa + td.start + x);
        uint4 s4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s4[k] = sp[k];
        u32 w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 b0 = prevb[s4[k].x], b1 = prevb[s4[k].y], b2 = prevb[s4[k].z], b3 = prevb[s4[k].w];
            w[k] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        *reinterpret_cast<uint4*>(dst + x) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (u32 x = n16 + threadIdx.x; x < td.count; x += KOLM_THREADS) dst[x] = prevb[sa[td.start + x]];
}

// ================================================================================================
// host drivers
// ================================================================================================
static inline int ceil_log2_u32(u32 v) { int b = 0; while ((1ull << b) < v) ++b; return b; }

// stable LSD radix sort of (K,V) records over `ntiles` tiles; result pointers returned in *Kr,*Vr.
static int radix_sort(kolm_ctx* c, const TileDesc* tiles, int ntiles, i64 nrec, const u32* tile0, const u32* tilen, int bits,
                      u32* Ka, u32* Va, u32* Kb, u32* Vb, u32** Kr, u32** Vr, cudaStream_t s, const u32* rekey_last = nullptr) {
    int passes = (bits + 7) / 8; if (passes < 1) passes = 1;
    int dbits = (bits + passes - 1) / passes; if (dbits < 1) dbits = 1;
    u32 mask = (1u << dbits) - 1;
    int sgrid = c->nblocks < 4 * c->sm_count ? c->nblocks : 4 * c->sm_count;
    for (int p = 0; p < passes; ++p) {
        int shift = p * dbits;
        KL(c, KC_HIST, nrec * 4 + (i64)ntiles * 1024, s, k_radix_hist<<<ntiles, KOLM_THREADS, 0, s>>>(Ka, tiles, c->d_thist, shift, mask));
        KL(c, KC_SCAN, (i64)ntiles * 2048, s, k_radix_scan<<<sgrid, 256 * RSCAN_GROUPS, 0, s>>>(c->d_thist, tile0, tilen, c->nblocks));
        KL(c, KC_SCATTER, nrec * 16 + (i64)ntiles * 1024, s, k_radix_scatter<<<ntiles, KOLM_THREADS, 0, s>>>(Ka, Va, Kb, Vb, tiles, c->d_binfo, c->d_thist, shift, mask, p + 1 == passes ? rekey_last : nullptr));
        u32* t = Ka; Ka = Kb; Kb = t; t = Va; Va = Vb; Vb = t;
    }
    CUDA_TRY(cudaGetLastError());
    *Kr = Ka; *Vr = Va;
    return KOLM_OK;
}

__global__ void k_lb_header(u64* lb, u64 rows, u64 nb, const u32* tile0, const u32* tilen, u64 group) {
    lb[1] = rows; lb[2] = nb; lb[3] = (u64)tile0; lb[4] = (u64)tilen; lb[5] = group;
}

// mode 0: sort kernels (rerank / gather) — block-major tickets by default: measured on B200, interleaving blocks there loses
//         more L2 locality of the rank gathers than it saves look-back steps (KOLM_LB_GROUP overrides).
// mode 1: cheap streaming scans (Rice cost/pack, residual/LZ emit, Lyndon, inverse offsets) — row-major over all blocks, so
//         that each block's look-back chain is one window deep instead of tiles_per_block/32 dependent steps.
int kolm_lb_reset_mode(kolm_ctx* c, bool active, int ntiles, int* grid, int mode, cudaStream_t s) {
    static int G0 = -1, G1 = -1;
    if (G0 < 0) { const char* e = getenv("KOLM_LB_GROUP"); G0 = e ? atoi(e) : 0; }
    if (G1 < 0) { const char* e = getenv("KOLM_LB_GROUP_STREAM"); G1 = e ? atoi(e) : (1 << 30); }
    int G = mode ? G1 : G0;
    if (G > c->nblocks) G = c->nblocks;
    u64 rows = active ? c->active_rows : c->static_rows;
    u64 ngroups = G > 0 ? ((u64)c->nblocks + G - 1) / G : 0;
    u64 g = rows * (u64)G * ngroups;
    if (G < 2 || rows < 2 || c->nblocks < 2 || g > 4ull * (u64)ntiles) { rows = 0; g = (u64)ntiles; }   // ragged batch: keep block-major order
    CUDA_TRY(cudaMemsetAsync(c->d_lb, 0, (size_t)(16 + 2 * g) * sizeof(u64), s));
    if (rows) k_lb_header<<<1, 1, 0, s>>>(c->d_lb, rows, (u64)c->nblocks, active ? c->d_atile0 : c->d_btile0, active ? c->d_atilen : c->d_btilen, (u64)G);
    *grid = (int)g;
    return KOLM_OK;
}
int kolm_lb_reset(kolm_ctx* c, bool active, int ntiles, int* grid, cudaStream_t s) { return kolm_lb_reset_mode(c, active, ntiles, grid, 0, s); }

// Sort all suffixes (plain) or all rotations (cyclic) of every block of the current batch.
// On return c->d_sa holds the order and c->d_rank the (group-start) ranks.  *rounds_out = doubling rounds run.
static int sort_batch(kolm_ctx* c, const u8* in, bool cyclic, int* rounds_out, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    if (!nt) { if (rounds_out) *rounds_out = 0; return KOLM_OK; }
    const size_t words = (c->total_elems + 31) / 32 + 1;
    CUDA_TRY(cudaMemsetAsync(c->d_single, 0, words * 4, s));
    CUDA_TRY(cudaMemsetAsync(c->d_done, 0, (size_t)nb * 4, s));
    CUDA_TRY(cudaMemsetAsync(c->d_newcls, 0, (size_t)nb * 4, s));
    // k_gather / k_ls_build write a block's active count from its LAST TILE: a block without tiles (empty) is never written, and
    // the array doubles as scratch of the Lyndon pass — without this an empty block enters k_plan_active with whatever the memory held
    CUDA_TRY(cudaMemsetAsync(c->d_active, 0, (size_t)nb * 4, s));
    int bgrid = nb < 1024 ? nb : 1024;
    // ---- bootstrap
    const i64 N = c->total_bytes;
    u32 h0 = cyclic ? 4 : 3;
    static int alpha = -1;
    if (alpha < 0) { const char* e = getenv("KOLM_ALPHA_BOOT"); alpha = e ? atoi(e) : 1; }
    if (cyclic && alpha) {
        // dense symbol ranks: more than 4 symbols per 32-bit bootstrap key when the batch's alphabets are small
        CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * 64 * 8, s));
        CUDA_TRY(cudaMemsetAsync(c->d_stats + 9, 0xff, 4, s));
        KL(c, KC_BOOT, N, s, k_alpha_mask<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_bacc));
        KL(c, KC_BOOT, (i64)nb * 512, s, k_alpha_map<<<nb, 256, 0, s>>>(c->d_bacc, c->d_binfo, c->d_stats + 9, nb));
        CUDA_TRY(cudaMemcpyAsync(c->h_stats + 9, c->d_stats + 9, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        u32 syms = c->h_stats[9];
        if (syms != 0xffffffffu && syms > 4) h0 = syms > 32 ? 32 : syms;
    }
    static int deep_mode = -1;
    if (deep_mode < 0) { const char* e = getenv("KOLM_DEEP_BOOT"); deep_mode = e ? atoi(e) : 1; }   // 0 never, 1 compressed alphabets, 2 always
    const bool deep = cyclic && (deep_mode == 2 || (deep_mode == 1 && h0 > 4));
    u32* K0 = deep ? c->d_nr : c->d_k0;                      // deep: the keys by position stay in d_nr (idle until the first round)
    if (cyclic && h0 > 4) KL(c, KC_BOOT, N * 9, s, k_boot_keys_alpha<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_binfo, c->d_tiles, c->d_fstart, c->d_nfac, c->d_bacc, h0, K0, c->d_v0));
    else if (cyclic) KL(c, KC_BOOT, N * 9, s, k_boot_keys<true><<<nt, KOLM_THREADS, 0, s>>>(in, c->d_binfo, c->d_tiles, c->d_fstart, c->d_nfac, K0, c->d_v0));
    else KL(c, KC_BOOT, N * 9, s, k_boot_keys<false><<<nt, KOLM_THREADS, 0, s>>>(in, c->d_binfo, c->d_tiles, c->d_fstart, c->d_nfac, c->d_k0, c->d_v0));
    u32 *K, *V;
    if (deep) {
        KL(c, KC_BOOT, N * 12, s, k_boot_lo<<<nt, KOLM_THREADS, 0, s>>>(c->d_nr, c->d_binfo, c->d_tiles, c->d_fstart, c->d_nfac, h0, c->d_k0, c->d_lo));
        // the last pass writes KH[V[j]] as its key column (k_boot_rekey fused into the scatter: one read of V and one write of K less)
        KOLM_TRY(radix_sort(c, c->d_tiles, nt, N, c->d_btile0, c->d_btilen, 32, c->d_k0, c->d_v0, c->d_k1, c->d_v1, &K, &V, s, c->d_nr));
        u32* Ko = (K == c->d_k0) ? c->d_k1 : c->d_k0;
        u32* Vo = (V == c->d_v0) ? c->d_v1 : c->d_v0;
        KOLM_TRY(radix_sort(c, c->d_tiles, nt, N, c->d_btile0, c->d_btilen, 32, K, V, Ko, Vo, &K, &V, s));
    } else {
        KOLM_TRY(radix_sort(c, c->d_tiles, nt, N, c->d_btile0, c->d_btilen, cyclic ? 32 : 27, c->d_k0, c->d_v0, c->d_k1, c->d_v1, &K, &V, s));
    }
    int lgrid = nt;
    KOLM_TRY(kolm_lb_reset(c, false, nt, &lgrid, s));
    static int local_rounds = -1;
    if (local_rounds < 0) { const char* e = getenv("KOLM_LOCAL_ROUNDS"); local_rounds = e ? atoi(e) : 1; }
    RerankArgs ra;
    ra.K = K; ra.V = V; ra.tiles = c->d_tiles; ra.binfo = c->d_binfo; ra.active = c->d_active; ra.fstart = c->d_fstart; ra.nfac = c->d_nfac;
    ra.lb = c->d_lb; ra.sa = c->d_sa; ra.rank = c->d_rank; ra.nr = c->d_nr; ra.single = c->d_single; ra.newcls = c->d_newcls; ra.h = deep ? h0 : 0; ra.lo = c->d_lo; ra.grp = local_rounds ? c->d_grp : nullptr;
    if (local_rounds) ra.single = nullptr;                   // the settled-bit map is read by k_gather only, which the local rounds never run
    if (deep) { KL(c, KC_RERANK, N * 20, s, k_rerank<2, true><<<lgrid, KOLM_THREADS, 0, s>>>(ra)); h0 *= 2; c->counters[4] += N; }   // second full sort
    else if (cyclic) KL(c, KC_RERANK, N * 16, s, k_rerank<1, true><<<lgrid, KOLM_THREADS, 0, s>>>(ra));
    else KL(c, KC_RERANK, N * 16, s, k_rerank<1, false><<<lgrid, KOLM_THREADS, 0, s>>>(ra));
    CUDA_TRY(cudaMemsetAsync(c->d_newcls, 0, (size_t)nb * 4, s));
    CUDA_TRY(cudaGetLastError());
    const int kbits = ceil_log2_u32(c->max_len > 1 ? c->max_len : 2);
    int rounds = 0;
    static int ls_div = -1;
    if (ls_div < 0) { const char* e = getenv("KOLM_LS_DIV"); ls_div = e ? atoi(e) : 8; }     // LS rounds once survivors < N / ls_div (0: never)
    ra.survivors = c->d_stats + 4;
    bool use_ls = false;
    u32 *Kprev = nullptr, *Vprev = nullptr;                  // sorted records of the previous round (aligned with d_nr)
    u64 hstart = h0;
    bool prebuilt = false;                                   // the first global round finds its (second key, position) records already emitted
    if (local_rounds) {
        // ---- local refinement rounds (k_refine_local) while the unsettled records are dense; groups longer than LR_GCAP go
        //      through the global LS path of the same round; once few records are left the compacted LS rounds below take over
        static bool attr_done = false;
        if (!attr_done) {
            CUDA_TRY(cudaFuncSetAttribute(k_refine_local<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LR_SMEM));
            CUDA_TRY(cudaFuncSetAttribute(k_refine_local<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LR_SMEM));
            attr_done = true;
        }
        CUDA_TRY(cudaMemsetAsync(c->d_live, 1, (size_t)nt, s));
        CUDA_TRY(cudaMemsetAsync(c->d_lact, 0, (size_t)nb * 4, s));
        RefineArgs fa;
        fa.sa = c->d_sa; fa.rank = c->d_rank; fa.grp = c->d_grp; fa.list = c->d_nr; fa.list2 = c->d_k0; fa.lcount = c->d_thist; fa.F = c->d_lo;
        fa.tiles = c->d_tiles; fa.binfo = c->d_binfo; fa.fstart = c->d_fstart; fa.nfac = c->d_nfac; fa.done = c->d_done; fa.live = c->d_live;
        fa.lact = c->d_lact; fa.newcls = c->d_newcls; fa.stats = c->d_stats; fa.gflag = c->d_v1;
        static int big_uniform = -1;
        if (big_uniform < 0) { const char* e = getenv("KOLM_BIG_UNIFORM"); big_uniform = e ? atoi(e) : 0; }
        fa.big_uniform = big_uniform;
        const u32 serial = ++c->sort_serial;
        static int trace = -1;
        if (trace < 0) { const char* e = getenv("KOLM_TRACE_ROUNDS"); trace = e ? atoi(e) : 0; }
        bool finished = false;
        u64 h = h0;
        for (; rounds < 64; h <<= 1) {
            if (h > 0x7fffffffull) h = 0x7fffffffull;
            fa.h = (u32)h;
            fa.stamp = 0x80000000u | ((serial & 0xffffffu) << 7) | (u32)(rounds & 127);   // never a value a sort leaves in d_v1 (positions < 2^31); a stale hit only costs work
            CUDA_TRY(cudaMemsetAsync(c->d_stats + 4, 0, 12, s));      // [4] survivors, [5] members of big groups, [6] unsettled at the start
            if (cyclic) KL(c, KC_RERANK, N * 4, s, k_refine_local<true><<<nt, KOLM_THREADS, LR_SMEM, s>>>(fa));
            else KL(c, KC_RERANK, N * 4, s, k_refine_local<false><<<nt, KOLM_THREADS, LR_SMEM, s>>>(fa));
            KL(c, KC_APPLY, N / 2, s, k_apply_local<<<nt, KOLM_THREADS, 0, s>>>(c->d_sa, c->d_nr, c->d_k0, c->d_thist, c->d_grp, c->d_rank, c->d_tiles, c->d_binfo, c->d_live));
            CUDA_TRY(cudaMemcpyAsync(c->h_stats + 4, c->d_stats + 4, 12, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            const u64 uns = c->h_stats[6], big = c->h_stats[5];
            if (trace) fprintf(stderr, "[kolm sort] %s local round %d h=%llu unsettled=%llu of %lld (%.3f) big=%llu survivors(local)=%u\n", cyclic ? "cyclic" : "plain",
                               rounds + 1, (unsigned long long)h, (unsigned long long)uns, (long long)N, (double)uns / (double)N, (unsigned long long)big, c->h_stats[4]);
            if (uns == 0) { finished = true; break; }
            ++rounds;
            c->counters[4] += (i64)uns;
            if (big) {
                // members of big groups (flagged in d_lo): (rank[succ_h(v)], v) per block -> sort -> (rank[v], v) -> stable sort -> rerank
                CUDA_TRY(cudaMemsetAsync(c->d_active, 0, (size_t)nb * 4, s));
                if (cyclic) KL(c, KC_GATHER, (i64)big * 16, s, k_big_emit<true, false><<<nt, KOLM_THREADS, 0, s>>>(c->d_sa, c->d_rank, c->d_lo, c->d_grp, c->d_tiles, c->d_binfo, c->d_fstart, c->d_nfac, c->d_done, c->d_live, c->d_active, c->d_k0, c->d_v0, (u32)h, fa.big_uniform ? c->d_v1 : nullptr, fa.stamp));
                else KL(c, KC_GATHER, (i64)big * 16, s, k_big_emit<false, false><<<nt, KOLM_THREADS, 0, s>>>(c->d_sa, c->d_rank, c->d_lo, c->d_grp, c->d_tiles, c->d_binfo, c->d_fstart, c->d_nfac, c->d_done, c->d_live, c->d_active, c->d_k0, c->d_v0, (u32)h, fa.big_uniform ? c->d_v1 : nullptr, fa.stamp));
                KL(c, KC_PLAN, (i64)nb * 16, s, k_plan_active<<<1, 1024, 0, s>>>(c->d_active, c->d_done, c->d_atile0, c->d_atilen, c->d_stats, nb));
                CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, 12, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                const int ant = (int)c->h_stats[0];
                const i64 M = (i64)c->h_stats[1];
                c->active_rows = c->h_stats[2];
                if (ant) {
                    KL(c, KC_TILES, (i64)ant * 16, s, k_build_tiles<<<bgrid, 128, 0, s>>>(c->d_binfo, c->d_atile0, c->d_atilen, c->d_active, c->d_atiles, nb));
                    u32 *K1, *V1;
                    KOLM_TRY(radix_sort(c, c->d_atiles, ant, M, c->d_atile0, c->d_atilen, kbits, c->d_k0, c->d_v0, c->d_k1, c->d_v1, &K1, &V1, s));
                    KL(c, KC_GATHER, M * 12, s, k_ls_key1<<<ant, KOLM_THREADS, 0, s>>>(K1, V1, c->d_atiles, c->d_rank));
                    u32* K1o = (K1 == c->d_k0) ? c->d_k1 : c->d_k0;
                    u32* V1o = (V1 == c->d_v0) ? c->d_v1 : c->d_v0;
                    KOLM_TRY(radix_sort(c, c->d_atiles, ant, M, c->d_atile0, c->d_atilen, kbits, K1, V1, K1o, V1o, &K, &V, s));
                    KOLM_TRY(kolm_lb_reset(c, true, ant, &lgrid, s));
                    ra.K = K; ra.V = V; ra.tiles = c->d_atiles; ra.h = (u32)h;
                    if (cyclic) KL(c, KC_RERANK, M * 20, s, k_rerank<0, true><<<lgrid, KOLM_THREADS, 0, s>>>(ra));
                    else KL(c, KC_RERANK, M * 20, s, k_rerank<0, false><<<lgrid, KOLM_THREADS, 0, s>>>(ra));
                    KL(c, KC_APPLY, M * 12, s, k_apply<<<ant, KOLM_THREADS, 0, s>>>(V, c->d_nr, c->d_atiles, c->d_rank, c->d_single));
                }
            }
            KL(c, KC_PLAN, (i64)nb * 12, s, k_round_end_local<<<(nb + 255) / 256, 256, 0, s>>>(c->d_newcls, c->d_done, c->d_lact, nb, cyclic ? 1 : 0));
            CUDA_TRY(cudaGetLastError());
            if (h >= 0x7fffffffull) { finished = true; break; }
            u64 surv = c->h_stats[4];
            if (big) {
                CUDA_TRY(cudaMemcpyAsync(c->h_stats + 4, c->d_stats + 4, 4, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                surv = c->h_stats[4];
            }
            if (surv == 0) { finished = true; break; }
            if (ls_div > 0 && surv * (u64)ls_div < (u64)N) {
                // few records left, spread over most tiles: hand the unsettled records (read off grp[]) to the compacted LS rounds
                h <<= 1;
                if (h > 0x7fffffffull) h = 0x7fffffffull;
                CUDA_TRY(cudaMemsetAsync(c->d_active, 0, (size_t)nb * 4, s));
                if (cyclic) KL(c, KC_GATHER, N * 4 + (i64)surv * 16, s, k_big_emit<true, true><<<nt, KOLM_THREADS, 0, s>>>(c->d_sa, c->d_rank, c->d_lo, c->d_grp, c->d_tiles, c->d_binfo, c->d_fstart, c->d_nfac, c->d_done, c->d_live, c->d_active, c->d_k0, c->d_v0, (u32)h, fa.big_uniform ? c->d_v1 : nullptr, fa.stamp));
                else KL(c, KC_GATHER, N * 4 + (i64)surv * 16, s, k_big_emit<false, true><<<nt, KOLM_THREADS, 0, s>>>(c->d_sa, c->d_rank, c->d_lo, c->d_grp, c->d_tiles, c->d_binfo, c->d_fstart, c->d_nfac, c->d_done, c->d_live, c->d_active, c->d_k0, c->d_v0, (u32)h, fa.big_uniform ? c->d_v1 : nullptr, fa.stamp));
                prebuilt = true; use_ls = true; Kprev = c->d_k1; Vprev = c->d_v1;      // "the other pair" of the LS branch is (d_k0, d_v0)
                break;
            }
        }
        if (finished || rounds >= 64) {
            if (rounds_out) *rounds_out = rounds + (deep ? 1 : 0);
            return KOLM_OK;
        }
        hstart = h;
        ra.grp = nullptr;                                    // the compacted rounds do not keep grp[] up to date (nothing reads it any more)
    }
    for (u64 h = hstart; rounds < 80; h <<= 1) {
        if (h > 0x7fffffffull) h = 0x7fffffffull;
        if (prebuilt) {
            // records emitted by k_big_emit<.., true>
        } else if (!use_ls) {
            // ---- Manber–Myers gather: stream the order, emit unsettled predecessors
            KOLM_TRY(kolm_lb_reset(c, false, nt, &lgrid, s));
            GatherArgs ga;
            ga.sa = c->d_sa; ga.rank = c->d_rank; ga.single = c->d_single; ga.tiles = c->d_tiles; ga.binfo = c->d_binfo; ga.fstart = c->d_fstart;
            ga.nfac = c->d_nfac; ga.done = c->d_done; ga.lb = c->d_lb; ga.K = c->d_k0; ga.V = c->d_v0; ga.active = c->d_active; ga.h = (u32)h;
            if (cyclic) KL(c, KC_GATHER, N * 4, s, k_gather<true><<<lgrid, KOLM_THREADS, 0, s>>>(ga));
            else KL(c, KC_GATHER, N * 4, s, k_gather<false><<<lgrid, KOLM_THREADS, 0, s>>>(ga));
        } else {
            // ---- Larsson–Sadakane build from the survivors of the previous round (old active tile map)
            const int pant = (int)c->h_stats[0];
            u32* Ko = (Kprev == c->d_k0) ? c->d_k1 : c->d_k0;
            u32* Vo = (Vprev == c->d_v0) ? c->d_v1 : c->d_v0;
            KOLM_TRY(kolm_lb_reset(c, true, pant, &lgrid, s));
            CUDA_TRY(cudaMemsetAsync(c->d_active, 0, (size_t)nb * 4, s));
            LsArgs la;
            la.V = Vprev; la.nr = c->d_nr; la.tiles = c->d_atiles; la.binfo = c->d_binfo; la.rank = c->d_rank; la.fstart = c->d_fstart;
            la.nfac = c->d_nfac; la.done = c->d_done; la.lb = c->d_lb; la.K2 = Ko; la.V2 = Vo; la.active = c->d_active; la.h = (u32)h;
            if (cyclic) KL(c, KC_GATHER, (i64)c->h_stats[1] * 12, s, k_ls_build<true><<<lgrid, KOLM_THREADS, 0, s>>>(la));
            else KL(c, KC_GATHER, (i64)c->h_stats[1] * 12, s, k_ls_build<false><<<lgrid, KOLM_THREADS, 0, s>>>(la));
        }
        KL(c, KC_PLAN, (i64)nb * 16, s, k_plan_active<<<1, 1024, 0, s>>>(c->d_active, c->d_done, c->d_atile0, c->d_atilen, c->d_stats, nb));
        CUDA_TRY(cudaMemcpyAsync(c->h_stats, c->d_stats, 12, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        int ant = (int)c->h_stats[0];
        const i64 M = (i64)c->h_stats[1];
        c->active_rows = c->h_stats[2];
        if (ant == 0) break;
        ++rounds;
        static int trace = -1;
        if (trace < 0) { const char* e = getenv("KOLM_TRACE_ROUNDS"); trace = e ? atoi(e) : 0; }
        if (trace) fprintf(stderr, "[kolm sort] %s round %d h=%llu active=%lld of %lld (%.3f) tiles=%d %s\n", cyclic ? "cyclic" : "plain", rounds,
                           (unsigned long long)h, (long long)M, (long long)N, (double)M / (double)N, ant, use_ls ? "LS" : "MM");
        c->counters[4] += M;                                  // active records summed over rounds
        KL(c, KC_TILES, (i64)ant * 16, s, k_build_tiles<<<bgrid, 128, 0, s>>>(c->d_binfo, c->d_atile0, c->d_atilen, c->d_active, c->d_atiles, nb));
        if (!use_ls) {
            KOLM_TRY(radix_sort(c, c->d_atiles, ant, M, c->d_atile0, c->d_atilen, kbits, c->d_k0, c->d_v0, c->d_k1, c->d_v1, &K, &V, s));
        } else {
            u32* Ko = (Kprev == c->d_k0) ? c->d_k1 : c->d_k0;
            u32* Vo = (Vprev == c->d_v0) ? c->d_v1 : c->d_v0;
            u32 *K1, *V1;
            KOLM_TRY(radix_sort(c, c->d_atiles, ant, M, c->d_atile0, c->d_atilen, kbits, Ko, Vo, Kprev, Vprev, &K1, &V1, s));
            KL(c, KC_GATHER, M * 12, s, k_ls_key1<<<ant, KOLM_THREADS, 0, s>>>(K1, V1, c->d_atiles, c->d_rank));
            u32* K1o = (K1 == c->d_k0) ? c->d_k1 : c->d_k0;
            u32* V1o = (V1 == c->d_v0) ? c->d_v1 : c->d_v0;
            KOLM_TRY(radix_sort(c, c->d_atiles, ant, M, c->d_atile0, c->d_atilen, kbits, K1, V1, K1o, V1o, &K, &V, s));
        }
        KOLM_TRY(kolm_lb_reset(c, true, ant, &lgrid, s));
        CUDA_TRY(cudaMemsetAsync(c->d_stats + 4, 0, 4, s));
        ra.K = K; ra.V = V; ra.tiles = c->d_atiles; ra.h = (u32)h;
        if (cyclic) KL(c, KC_RERANK, M * 20, s, k_rerank<0, true><<<lgrid, KOLM_THREADS, 0, s>>>(ra));
        else KL(c, KC_RERANK, M * 20, s, k_rerank<0, false><<<lgrid, KOLM_THREADS, 0, s>>>(ra));
        KL(c, KC_APPLY, M * 12, s, k_apply<<<ant, KOLM_THREADS, 0, s>>>(V, c->d_nr, c->d_atiles, c->d_rank, c->d_single));
        KL(c, KC_PLAN, (i64)nb * 12, s, k_round_end<<<(nb + 255) / 256, 256, 0, s>>>(c->d_newcls, c->d_done, c->d_active, nb, cyclic ? 1 : 0));
        CUDA_TRY(cudaGetLastError());
        if (h >= 0x7fffffffull) break;
        // survivors decide how the next round starts (and whether there is one)
        CUDA_TRY(cudaMemcpyAsync(c->h_stats + 4, c->d_stats + 4, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        const u64 surv = c->h_stats[4];
        if (surv == 0) break;
        Kprev = K; Vprev = V;
        prebuilt = false;
        use_ls = local_rounds ? true : (ls_div > 0 && surv * (u64)ls_div < (u64)N);   // after local rounds there is no settled-bit map for k_gather
    }
    if (rounds_out) *rounds_out = rounds + (deep ? 1 : 0);    // doublings of the sorted depth: the deep bootstrap is one (h0 -> 2*h0)
    return KOLM_OK;
}

// a1: Lyndon factorisation of every block.  Leaves factor lists in the context for the cyclic sort.
int kolm_lyndon_impl(kolm_ctx* c, const u8* in, u8* flags_out, int* rounds_out, cudaStream_t s) {
    if (!c->ntiles) return KOLM_OK;
    static int fast = -1;
    if (fast < 0) { const char* e = getenv("KOLM_LYNDON_FAST"); fast = e ? atoi(e) : 1; }
    int lgrid = c->ntiles;
    if (fast) {
        // fast path: candidate scan + per-block tie resolution; falls back to the ISA path if any block is degenerate
        u32* cand = c->d_nr;                                  // free until the first sort round
        u32* ncand = c->d_active;
        CUDA_TRY(cudaMemsetAsync(c->d_stats + 8, 0, 4, s));
        if (flags_out && c->total_bytes) CUDA_TRY(cudaMemsetAsync(flags_out + c->h_binfo[0].ioff, 0, (size_t)c->total_bytes, s));
        KOLM_TRY(kolm_lb_reset_mode(c, false, c->ntiles, &lgrid, 1, s));
        KL(c, KC_LYNDON, c->total_bytes * 2, s, k_lyn_cand<<<lgrid, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_lb, cand, ncand));
        u32* blockfail = c->d_atilen;                          // free until the first sort round
        KL(c, KC_LYNDON, c->total_bytes / 64, s, k_lyn_resolve<<<(c->nblocks + 3) / 4, 128, 0, s>>>(in, c->d_binfo, cand, ncand, c->d_fstart, c->d_nfac,
                                                                                                  flags_out, c->d_stats + 8, blockfail, c->nblocks));
        CUDA_TRY(cudaMemcpyAsync(c->h_stats + 8, c->d_stats + 8, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(cudaGetLastError());
        if (c->h_stats[8] == 0) { if (rounds_out) *rounds_out = 0; return KOLM_OK; }
        static int duval = -1;
        if (duval < 0) { const char* e = getenv("KOLM_LYNDON_DUVAL"); duval = e ? atoi(e) : 1; }
        static u32 duval_max = 0;
        if (!duval_max) { const char* e = getenv("KOLM_DUVAL_MAX_MIB"); duval_max = e ? (u32)atoi(e) << 20 : DUVAL_MAX_BLOCK_DEFAULT; }
        if (duval) {
            // second chance for the blocks that failed: vectorised Duval walk, one warp per block
            CUDA_TRY(cudaMemsetAsync(c->d_stats + 8, 0, 4, s));
            KL(c, KC_LYNDON, c->total_bytes / 8, s, k_lyn_duval<<<(c->nblocks + 3) / 4, 128, 0, s>>>(in, c->d_binfo, blockfail, c->d_fstart, c->d_nfac, flags_out,
                                                                                                   c->d_stats + 8, c->nblocks, duval_max));
            CUDA_TRY(cudaMemcpyAsync(c->h_stats + 8, c->d_stats + 8, 4, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            CUDA_TRY(cudaGetLastError());
            if (c->h_stats[8] == 0) { if (rounds_out) *rounds_out = 0; return KOLM_OK; }
        }
    }
    KOLM_TRY(sort_batch(c, in, false, rounds_out, s));
    KOLM_TRY(kolm_lb_reset_mode(c, false, c->ntiles, &lgrid, 1, s));
    KL(c, KC_LYNDON, c->total_bytes * 4, s, k_lyndon<<<lgrid, KOLM_THREADS, 0, s>>>(c->d_rank, c->d_tiles, c->d_binfo, c->d_lb, c->d_fstart, c->d_nfac, flags_out));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}

// a2: BBWT of every block
int kolm_bbwt_fwd_impl(kolm_ctx* c, const u8* in, u8* out, int* rounds_plain, int* rounds_cyclic, cudaStream_t s) {
    if (!c->ntiles) return KOLM_OK;
    KOLM_TRY(kolm_lyndon_impl(c, in, nullptr, rounds_plain, s));
    KOLM_TRY(sort_batch(c, in, true, rounds_cyclic, s));
    u8* prevb = reinterpret_cast<u8*>(c->d_k1);              // the sort buffers are idle again
    KL(c, KC_EMIT, c->total_bytes * 2, s, k_prev_bytes<<<c->ntiles, KOLM_THREADS, 0, s>>>(in, prevb, c->d_tiles, c->d_binfo));
    KL(c, KC_EMIT, (i64)c->nblocks * 64, s, k_prev_patch<<<c->nblocks < 1024 ? c->nblocks : 1024, 256, 0, s>>>(in, prevb, c->d_binfo, c->d_fstart, c->d_nfac, c->nblocks));
    KL(c, KC_EMIT, c->total_bytes * 6, s, k_bbwt_emit<<<c->ntiles, KOLM_THREADS, 0, s>>>(prevb, out, c->d_sa, c->d_tiles, c->d_binfo));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}
