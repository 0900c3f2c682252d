// mtf.cu — move-to-front encode / decode.
//
// Replaces mtf_encode / mtf_decode (kolm_final.py:375-405 == kolm_final_researched_v2-2.py:460-478).
//
// Encode uses the timestamp form: with last[c] = time of the last occurrence of c
// (never seen: -c, i.e. identity order behind every seen symbol), the MTF index of s at time t is
// #{c : last[c] > last[s]}.  Each tile (<= 4096 bytes of one block) is handled by one warp:
//   pass A  per-tile last occurrence of every symbol,
//   pass B  per-block exclusive max-scan over tiles  -> the table on entry to every tile,
//   pass C  the warp walks its tile 32 bytes at a time; only run heads (s[t] != s[t-1]) cost work:
//           8 compares per lane + one REDUX add; run bodies are zeros.
// Decode: pass A computes the permutation each tile applies to the list, pass B composes them per
// block, pass C replays the tile from its true entry list.
#include "common.cuh"

#define MTF_WARPS 4

// Sub-tiles: the encode passes can cut every tile into 2^ss pieces of (KOLM_TILE >> ss) bytes (piece x = tile (x >> ss), part
// (x & mask); parts beyond the tile's count are empty).  More pieces = more independent list walks in k_mtf_enc2, which is
// bound by the latency of its serial walk and not by traffic (r1 ncu: 10-20 % of the warp slots occupied at 4096 bytes/thread).
__device__ __forceinline__ void mtf_piece(const TileDesc& td, u32 x, u32 ss, u32& start, u32& count) {
    const u32 part = x & ((1u << ss) - 1u), plen = KOLM_TILE >> ss, o = part * plen;
    start = td.start + o;
    count = td.count > o ? min(plen, td.count - o) : 0u;
}

// pass A (encode): tlast[piece][c] = 1 + block-local position of the last c in the piece, 0 if absent.
// With G > 1 the pieces of a block form G consecutive groups whose column maxima are collected in gt[block][group][c]
// (pass B then scans the groups independently).
__global__ void __launch_bounds__(KOLM_THREADS) k_mtf_last(const u8* __restrict__ in, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                           const u32* __restrict__ tile0, const u32* __restrict__ tilen, u32* __restrict__ tlast,
                                                           u32* __restrict__ gt, u32 ss, u32 G) {
    __shared__ u32 last[256];
    TileDesc td = tiles[blockIdx.x >> ss];
    BlockInfo bi = binfo[td.block];
    u32 start, count;
    mtf_piece(td, blockIdx.x, ss, start, count);
    last[threadIdx.x] = 0;
    __syncthreads();
    u32 t0 = start - bi.pbase;
    const u8* src = in + bi.ioff + t0;
    if ((((uintptr_t)src) & 15) == 0) {                        // 16 bytes per thread and load: only run ends touch shared memory
        for (u32 x = threadIdx.x * 16; x < count; x += KOLM_THREADS * 16) {
            const u32 nb = min(16u, count - x);
            u32 w[5] = {0, 0, 0, 0, 0};
            if (nb == 16) { const uint4 q = *reinterpret_cast<const uint4*>(src + x); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
            else for (u32 i = 0; i < nb; ++i) w[i >> 2] |= (u32)src[x + i] << (8 * (i & 3));
            const bool more = x + nb < count;                   // the byte after mine decides whether my last byte ends a run
            if (more) w[4] = src[x + nb];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((u32)i < nb) {
                    const u32 sb = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                    const bool last_of_mine = (u32)i + 1 == nb;
                    const u32 nx = last_of_mine ? (w[4] & 0xFFu) : ((w[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFFu);
                    if ((last_of_mine && !more) || nx != sb) atomicMax(&last[sb], t0 + x + i + 1);
                }
            }
        }
    } else {
        for (u32 x = threadIdx.x; x < count; x += KOLM_THREADS) {
            u8 s = src[x];
            if (x + 1 == count || src[x + 1] != s) atomicMax(&last[s], t0 + x + 1);
        }
    }
    __syncthreads();
    const u32 v = last[threadIdx.x];
    tlast[(size_t)blockIdx.x * 256 + threadIdx.x] = v;
    if (G > 1 && v) {
        const u32 np = tilen[td.block] << ss, Q = (np + G - 1) / G;
        const u32 g = (blockIdx.x - (tile0[td.block] << ss)) / Q;
        atomicMax(gt + ((size_t)td.block * G + g) * 256 + threadIdx.x, v);
    }
}

// pass B (encode): exclusive running max over the pieces of each block; CTA (block, group) starts from the maxima of the
// groups before it and walks its own rows, eight in flight per step
__global__ void __launch_bounds__(256) k_mtf_scan_max(u32* __restrict__ tlast, const u32* __restrict__ tile0, const u32* __restrict__ tilen,
                                                      const u32* __restrict__ gt, int nblocks, u32 ss, u32 G) {
    for (u32 x = blockIdx.x; x < (u32)nblocks * G; x += gridDim.x) {
        const u32 b = x / G, g = x - b * G;
        const u32 np = tilen[b] << ss, Q = (np + G - 1) / G;
        const u32 r0 = g * Q, r1 = min(np, r0 + Q);
        if (r0 >= r1) continue;
        u32 run = 0;
        for (u32 k = 0; k < g; ++k) run = max(run, gt[((size_t)b * G + k) * 256 + threadIdx.x]);
        u32* base = tlast + ((size_t)tile0[b] << ss) * 256 + threadIdx.x;
        for (u32 t = r0; t < r1; t += 8) {
            u32 v[8];
#pragma unroll
            for (u32 k = 0; k < 8; ++k) v[k] = t + k < r1 ? base[(size_t)(t + k) * 256] : 0u;
#pragma unroll
            for (u32 k = 0; k < 8; ++k) if (t + k < r1) { base[(size_t)(t + k) * 256] = run; run = max(run, v[k]); }
        }
    }
}

// pass C (encode): one warp per tile
__global__ void __launch_bounds__(MTF_WARPS * 32) k_mtf_enc(const u8* __restrict__ in, u8* __restrict__ out, const TileDesc* __restrict__ tiles,
                                                            const BlockInfo* __restrict__ binfo, const u32* __restrict__ tlast, int ntiles) {
    __shared__ int wts[MTF_WARPS][256];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * MTF_WARPS + w;
    if (tile >= ntiles) return;
    TileDesc td = tiles[tile];
    BlockInfo bi = binfo[td.block];
    int ts[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        u32 c = lane * 8 + k;
        u32 e = tlast[(size_t)tile * 256 + c];
        ts[k] = e ? (int)e : -(int)c;
        wts[w][c] = ts[k];
    }
    __syncwarp();
    const u32 t0 = td.start - bi.pbase;
    const u8* src = in + bi.ioff + t0;
    u8* dst = out + bi.ioff + t0;
    u32 prevlast = t0 ? (u32)src[-1] : 0x100u;            // byte before the tile (0x100: none)
    for (u32 x0 = 0; x0 < td.count; x0 += 32) {
        u32 x = x0 + lane;
        bool valid = x < td.count;
        u32 s = valid ? src[x] : 0;
        u32 sp = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane == 0) sp = prevlast;
        u32 heads = __ballot_sync(0xffffffffu, valid && s != sp);
        u32 myidx = 0;
        while (heads) {
            u32 pos = __ffs(heads) - 1; heads &= heads - 1;
            u32 hs = __shfl_sync(0xffffffffu, s, pos);
            int thr = wts[w][hs];
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) cnt += (ts[k] > thr) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            int nts = (int)(t0 + x0 + pos + 1);
            if (lane == (hs >> 3)) {
#pragma unroll
                for (int k = 0; k < 8; ++k) if ((u32)k == (hs & 7u)) ts[k] = nts;
            }
            __syncwarp();
            if (lane == 0) wts[w][hs] = nts;
            __syncwarp();
            if (lane == pos) myidx = (u32)cnt;
        }
        if (valid) dst[x] = (u8)myidx;
        prevlast = __shfl_sync(0xffffffffu, s, 31);
    }
}


// pass C (encode), v2: one THREAD per tile, 128 tiles per CTA.
//   step 1 (warp-cooperative): turn each tile's entry timestamps into its entry list — seen symbols by last occurrence
//           (rank by counting among the seen ones only), then the never-seen ones in identity order;
//   step 2 (thread-serial): the classic list walk, but on a packed list (4 symbols per 32-bit word, word 0 in a register,
//           words 1..63 in shared memory laid out [word][thread] = conflict free): zero-byte test finds the symbol inside a
//           word, funnel shift + byte permute moves it to the front.  BWT output is dominated by small indices, so most
//           symbols never leave the register word.
#define MTF2_THREADS 128
__device__ __forceinline__ u32 mtf_zero_byte(u32 x) { return (x - 0x01010101u) & ~x & 0x80808080u; }

__global__ void __launch_bounds__(MTF2_THREADS) k_mtf_enc2(const u8* __restrict__ in, u8* __restrict__ out, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, const u32* __restrict__ tlast, int ntiles, u32 ss) {
    __shared__ u32 lst[64][MTF2_THREADS];                  // packed entry lists, word k of thread t at lst[k][t]
    __shared__ u32 s_ts[MTF2_THREADS / 32][256];           // per warp: timestamps of the tile being converted
    __shared__ u8 s_sym[MTF2_THREADS / 32][256];
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    u8* lb = reinterpret_cast<u8*>(&lst[0][0]);
    // ---- step 1: warp w builds the lists of tiles (blockIdx*128 + w*32 + q), q = 0..31
    for (u32 q = 0; q < 32; ++q) {
        const u32 T = w * 32 + q;                          // thread column that will own this tile
        const int tile = blockIdx.x * MTF2_THREADS + T;
        if (tile >= ntiles) break;
        u32 ts[8]; u32 seenmask = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ts[k] = tlast[(size_t)tile * 256 + k * 32 + lane]; if (ts[k]) seenmask |= 1u << k; }   // symbol c = k*32+lane
        // compact the seen symbols (any order) into smem
        u32 nseen = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            u32 bal = __ballot_sync(0xffffffffu, ts[k] != 0);
            if (ts[k]) { u32 o = nseen + __popc(bal & lanemask_lt()); s_ts[w][o] = ts[k]; s_sym[w][o] = (u8)(k * 32 + lane); }
            nseen += __popc(bal);
        }
        __syncwarp();
        // rank by counting: position of a seen symbol = #seen symbols with a larger timestamp (timestamps are distinct)
        for (u32 e = lane; e < nseen; e += 32) {
            u32 my = s_ts[w][e], r = 0;
            for (u32 f = 0; f < nseen; ++f) r += (s_ts[w][f] > my);
            u32 p = r;
            lb[(((p >> 2) * MTF2_THREADS + T) << 2) + (p & 3)] = s_sym[w][e];
        }
        // never-seen symbols keep identity order behind the seen ones: position = nseen + #unseen symbols smaller than c
        u32 below = 0;                                      // unseen symbols in rows k' < k (all lanes) handled incrementally
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            u32 bal = __ballot_sync(0xffffffffu, ts[k] == 0);
            if (!ts[k]) { u32 p = nseen + below + __popc(bal & lanemask_lt()); lb[(((p >> 2) * MTF2_THREADS + T) << 2) + (p & 3)] = (u8)(k * 32 + lane); }
            below += __popc(bal);
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- step 2
    const int tile = blockIdx.x * MTF2_THREADS + tid;          // piece index (ntiles = number of pieces)
    if (tile >= ntiles) return;
    const TileDesc td = tiles[(u32)tile >> ss];
    const BlockInfo bi = binfo[td.block];
    u32 pstart, pcount;
    mtf_piece(td, (u32)tile, ss, pstart, pcount);
    const u32 t0 = pstart - bi.pbase;
    const u8* src = in + bi.ioff + t0;
    u8* dst = out + bi.ioff + t0;
    u32 w0 = lst[0][tid];
    const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    for (u32 x0 = 0; x0 < pcount; x0 += 16) {
        u32 inw[4], outw[4] = {0, 0, 0, 0};
        const u32 nb = min(16u, pcount - x0);
        if (aligned && nb == 16) { uint4 v = *reinterpret_cast<const uint4*>(src + x0); inw[0] = v.x; inw[1] = v.y; inw[2] = v.z; inw[3] = v.w; }
        else { inw[0] = inw[1] = inw[2] = inw[3] = 0; for (u32 i = 0; i < nb; ++i) inw[i >> 2] |= (u32)src[x0 + i] << (8 * (i & 3)); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((u32)i < nb) {
                const u32 b = (inw[i >> 2] >> (8 * (i & 3))) & 0xFF;
                u32 idx = 0;
                if (b != (w0 & 0xFF)) {
                    const u32 bbbb = b * 0x01010101u;
                    u32 z = mtf_zero_byte(w0 ^ bbbb);
                    if (z) {                                // inside the register word: positions 1..3
                        idx = (__ffs(z) - 1) >> 3;
                        w0 = __byte_perm(w0, b, idx == 1 ? 0x3204 : idx == 2 ? 0x3104 : 0x2104);
                    } else {
                        u32 carry = w0 >> 24;
                        w0 = (w0 << 8) | b;
                        for (u32 k = 1;; ++k) {
                            u32 wk = lst[k][tid];
                            z = mtf_zero_byte(wk ^ bbbb);
                            if (!z) { lst[k][tid] = (wk << 8) | carry; carry = wk >> 24; continue; }
                            u32 j = (__ffs(z) - 1) >> 3;
                            idx = 4 * k + j;
                            lst[k][tid] = __byte_perm(wk, carry, j == 0 ? 0x3214 : j == 1 ? 0x3204 : j == 2 ? 0x3104 : 0x2104);
                            break;
                        }
                    }
                }
                outw[i >> 2] |= idx << (8 * (i & 3));
            }
        }
        if (aligned && nb == 16) *reinterpret_cast<uint4*>(dst + x0) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
        else for (u32 i = 0; i < nb; ++i) dst[x0 + i] = (u8)(outw[i >> 2] >> (8 * (i & 3)));
    }
}

// ---------------------------------------------------------------------------------------------
// encode, v3 (default): local lists, list-prefix scan, miss fix-up.  No timestamp tables, no per-piece entry lists.
//
//   k_mtf3_walk   one THREAD per 1 KiB piece walks its bytes from the EMPTY list.  A symbol already seen in the piece has its true
//                 MTF index (everything between its two occurrences lies inside the piece); a first occurrence ("miss") is
//                 recorded (symbol, position) and patched later.  The walk alternates three warp-uniform phases per 32 bytes:
//                   A   all lanes run the same branch-free code on the four front entries (one register): the entry is found
//                       there (index 0..3) or not; either way the register word is final — a symbol not found in front is pushed
//                       on the front and the fourth entry drops to the next level — so the phase never waits for a search;
//                   A2  the bytes that missed the front (19 % on BWT of text) are drained in order against entries 4..15, three
//                       more register words searched and rotated branch-free by all lanes (the loop runs as often as the lane
//                       with the most such bytes needs); what drops out of entry 15 goes on;
//                   B   the rest (3 %) searches the deep part (packed words in shared memory, [word][thread]) in order.
//                 The divergence of the old walk (every byte step cost the warp its slowest lane's search) is confined to phase B.
//   k_mtf3_agg    one warp per group of 32 pieces: the group's list = fold of its pieces' final lists under
//                 later (+) earlier = later, then earlier without the symbols of later.
//   k_mtf3_prefix one warp per block: entry list of every group (identity list at the block start).
//   k_mtf3_fix    one warp per group: entry list E of every piece; the m-th miss of a piece (symbol c, entry index r in E) has
//                 index r + #{earlier misses of the piece with a larger entry index} (the m symbols seen so far sit in front, the
//                 rest of E keeps its order).
// ---------------------------------------------------------------------------------------------
#define M3_THREADS 128
#define M3_SLOT 1024                                       // scratch bytes per piece: final list [256] | misses in order [256] | their positions [256] u16
#define M3_GROUP_TILES 8                                   // tiles per scan group (32 pieces of 1 KiB)

template <bool WHOLE>
__device__ __forceinline__ void m3_byte(const u32 b, const u32 i, const u32 nb, u32& w0, u32& vmask, u32& dm, u32& ow) {
    const u32 x = w0 ^ (b * 0x01010101u);
    const u32 z = (x - 0x01010101u) & ~x & vmask;          // lowest set bit = the entry that equals b (bits above it may be false)
    const bool ok = WHOLE || i < nb;
    u32 M = z ^ (z - 1u);                                  // bytes 0..idx; all ones when b is not among the front entries
    if (!WHOLE && !ok) M = 0u;
    const u32 sh = __byte_perm(w0, b, 0x2104);            // (w0 << 8) | b
    const u32 e = w0 >> 24;                                // the entry that drops out of the register when b is pushed
    w0 = (w0 & ~M) | (sh & M);
    const u32 val = z ? (u32)__popc(M & 0x01010100u) : e;
    ow |= val << (8 * (i & 3));
    if (ok && !z) { dm |= 1u << i; vmask = (vmask << 8) | 0x80u; }
}

__global__ void __launch_bounds__(M3_THREADS) k_mtf3_walk(const u8* __restrict__ in, u8* __restrict__ out, const TileDesc* __restrict__ tiles,
                                                          const BlockInfo* __restrict__ binfo, u8* __restrict__ slots, u32* __restrict__ npc, int np) {
    __shared__ u32 lst[60][M3_THREADS];                    // deep part: list word k >= 4 (entries 4k..4k+3) of thread t at lst[k-4][t]
    __shared__ u32 ist[8][M3_THREADS];                     // the group's input bytes
    __shared__ u32 ost[8][M3_THREADS];                     // the group's output bytes (phase B patches them)
    const u32 tid = threadIdx.x;
    const int piece = blockIdx.x * M3_THREADS + tid;
    u32 pcount = 0; const u8* src = nullptr; u8* dst = nullptr;
    if (piece < np) {
        const TileDesc td = tiles[(u32)piece >> 2];
        const BlockInfo bi = binfo[td.block];
        u32 pstart;
        mtf_piece(td, (u32)piece, 2u, pstart, pcount);
        const u32 t0 = pstart - bi.pbase;
        src = in + bi.ioff + t0; dst = out + bi.ioff + t0;
    }
    u8* slot = slots + (size_t)(piece < np ? piece : 0) * M3_SLOT;
    const bool al_in = (((uintptr_t)src) & 15) == 0, al_out = (((uintptr_t)dst) & 15) == 0;
    u8* ib = reinterpret_cast<u8*>(&ist[0][0]) + 4 * tid;  // byte i of my group at ib[(i >> 2) * 4 * M3_THREADS + (i & 3)]
    u8* ob = reinterpret_cast<u8*>(&ost[0][0]) + 4 * tid;
    u32 w0 = 0, vmask = 0, n = 0;
    u32 w1 = 0, w2 = 0, w3 = 0, vm1 = 0, vm2 = 0, vm3 = 0;   // entries 4..15 and their valid-byte masks
    const u32 maxc = __reduce_max_sync(0xffffffffu, pcount);
    for (u32 x0 = 0; x0 < maxc; x0 += 32) {
        const u32 nb = pcount > x0 ? min(32u, pcount - x0) : 0u;
        u32 inw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) inw[j] = 0;
        if (nb) {
            const u8* g = src + x0;
            if (al_in) {                                   // a 16-byte load that holds one valid byte stays inside the allocation
                const uint4 q0 = *reinterpret_cast<const uint4*>(g);
                inw[0] = q0.x; inw[1] = q0.y; inw[2] = q0.z; inw[3] = q0.w;
                if (nb > 16) { const uint4 q1 = *reinterpret_cast<const uint4*>(g + 16); inw[4] = q1.x; inw[5] = q1.y; inw[6] = q1.z; inw[7] = q1.w; }
            } else {                                       // aligned words + funnel shift (same argument for the words at the ends)
                const u32 s = (u32)((uintptr_t)g & 3);
                const u32* gw = reinterpret_cast<const u32*>(g - s);
                u32 a[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) a[j] = (4u * j < nb + s) ? gw[j] : 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) inw[j] = __funnelshift_r(a[j], a[j + 1], 8 * s);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) ist[j][tid] = inw[j];
        // ---- phase A
        u32 dm = 0;
        if (__all_sync(0xffffffffu, nb == 32)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                u32 ow = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) m3_byte<true>((inw[j] >> (8 * q)) & 0xFFu, 4 * j + q, 32u, w0, vmask, dm, ow);
                ost[j][tid] = ow;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                u32 ow = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) m3_byte<false>((inw[j] >> (8 * q)) & 0xFFu, 4 * j + q, nb, w0, vmask, dm, ow);
                ost[j][tid] = ow;
            }
        }
        // ---- phase A2: the bytes that were not among the front four, in order, against entries 4..15 (three more register
        //      words).  Every lane runs the same branch-free search / rotate; 97 % of BWT-of-text bytes end here at the latest.
        u32 dm2 = 0;
        while (__any_sync(0xffffffffu, dm != 0)) {
            const bool act = dm != 0;
            u32 i = 0, bo = 0, c = 0, e = 0;
            if (act) { i = __ffs(dm) - 1; dm &= dm - 1; bo = (i >> 2) * (4 * M3_THREADS) + (i & 3); c = ib[bo]; e = ob[bo]; }
            const bool l1 = act && n >= 4;                 // with fewer than four entries everything sits in the front word
            const u32 bbbb = c * 0x01010101u;
            const u32 x1 = w1 ^ bbbb, x2 = w2 ^ bbbb, x3 = w3 ^ bbbb;
            const u32 z1 = (x1 - 0x01010101u) & ~x1 & vm1, z2 = (x2 - 0x01010101u) & ~x2 & vm2, z3 = (x3 - 0x01010101u) & ~x3 & vm3;
            const u32 M1 = l1 ? (z1 ^ (z1 - 1u)) : 0u;
            const u32 M2 = (l1 && !z1) ? (z2 ^ (z2 - 1u)) : 0u;
            const u32 M3 = (l1 && !z1 && !z2) ? (z3 ^ (z3 - 1u)) : 0u;
            const u32 c1 = w1 >> 24, c2 = w2 >> 24, e3 = w3 >> 24;
            w1 = (w1 & ~M1) | (((w1 << 8) | e) & M1);
            w2 = (w2 & ~M2) | (((w2 << 8) | c1) & M2);
            w3 = (w3 & ~M3) | (((w3 << 8) | c2) & M3);
            if (act) {
                if (l1 && (z1 | z2 | z3)) {
                    ob[bo] = (u8)(z1 ? 4 + __popc(M1 & 0x01010100u) : z2 ? 8 + __popc(M2 & 0x01010100u) : 12 + __popc(M3 & 0x01010100u));
                } else if (n < 16) {                       // first occurrence in the piece (the deep part is still empty): k_mtf3_fix writes its index
                    slot[256 + n] = (u8)c;
                    reinterpret_cast<u16*>(slot + 512)[n] = (u16)(x0 + i);
                    if (n >= 4) { const u32 j = n - 4, bit = 0x80u << (8 * (j & 3)); if (j < 4) vm1 |= bit; else if (j < 8) vm2 |= bit; else vm3 |= bit; }
                    ++n;
                } else { ob[bo] = (u8)e3; dm2 |= 1u << i; }   // on to the deep part, with the entry that dropped out of the registers
            }
        }
        // ---- phase B: the group's deep searches (entries 16 and up, shared memory), in order
        while (__any_sync(0xffffffffu, dm2 != 0)) {
            if (dm2) {
                const u32 i = __ffs(dm2) - 1; dm2 &= dm2 - 1;
                const u32 bo = (i >> 2) * (4 * M3_THREADS) + (i & 3);
                const u32 c = ib[bo];
                bool miss = true;
                {
                    const u32 bbbb = c * 0x01010101u, last = n >> 2;      // n >= 16 here
                    u32 carry = ob[bo];
                    for (u32 k = 4;; ++k) {
                        const u32 wk = lst[k - 4][tid];
                        const u32 vm = k == last ? (0x00808080u >> (8 * (3 - (n & 3)))) : 0x80808080u;   // word `last` holds n & 3 entries
                        const u32 x = wk ^ bbbb;
                        const u32 z = (x - 0x01010101u) & ~x & vm;
                        const u32 M = z ^ (z - 1u);
                        const u32 sh = (wk << 8) | carry;
                        lst[k - 4][tid] = (wk & ~M) | (sh & M);
                        if (z) { ob[bo] = (u8)(4 * k + __popc(M & 0x01010100u)); miss = false; break; }
                        if (k == last) break;
                        carry = wk >> 24;
                    }
                }
                if (miss) {                                // first occurrence in the piece: k_mtf3_fix writes its index
                    slot[256 + n] = (u8)c;
                    reinterpret_cast<u16*>(slot + 512)[n] = (u16)(x0 + i);
                    ++n;
                }
            }
        }
        // ---- store
        if (nb) {
            u8* g = dst + x0;
            if (al_out && nb == 32) {
                *reinterpret_cast<uint4*>(g) = make_uint4(ost[0][tid], ost[1][tid], ost[2][tid], ost[3][tid]);
                *reinterpret_cast<uint4*>(g + 16) = make_uint4(ost[4][tid], ost[5][tid], ost[6][tid], ost[7][tid]);
            } else {
                for (u32 i = 0; i < nb; ++i) g[i] = ob[(i >> 2) * (4 * M3_THREADS) + (i & 3)];
            }
        }
    }
    if (piece < np) {
        npc[piece] = n;
        u32* Lw = reinterpret_cast<u32*>(slot);
        if (n) Lw[0] = w0;
        if (n > 4) Lw[1] = w1;
        if (n > 8) Lw[2] = w2;
        if (n > 12) Lw[3] = w3;
        for (u32 k = 4; 4 * k < n; ++k) Lw[k] = lst[k - 4][tid];
    }
}

struct __align__(16) M3Warp { u8 a[256]; u8 b[256]; u8 inv[256]; u8 r[256]; u32 mask[8]; };

// cur (cnt entries) <- L (n entries, all distinct) followed by the entries of cur that are not in L; l0 = L[lane] (prefetched, lane < n).
// The current list is W.a or W.b, chosen by `sel` (an index, not a pair of swapped pointers: with the swap nvcc 12.9 addressed
// W.mask relative to the loop-carried `cur`, i.e. 256 bytes too far in every second call, and cleared the next warp's list instead).
__device__ __forceinline__ void m3_combine(M3Warp& W, u32& sel, u32& cnt, const u8* __restrict__ L, u32 n, u32 l0) {
    const u32 lane = lane_id();
    if (!n) return;
    u8* const cur = sel ? W.b : W.a;
    u8* const nxt = sel ? W.a : W.b;
    u32* const mask = W.mask;
    if (lane < 8) mask[lane] = 0;
    __syncwarp();
    if (lane < n) { nxt[lane] = (u8)l0; atomicOr(&mask[l0 >> 5], 1u << (l0 & 31)); }
    for (u32 i = lane + 32; i < n; i += 32) { const u32 c = L[i]; nxt[i] = (u8)c; atomicOr(&mask[c >> 5], 1u << (c & 31)); }
    __syncwarp();
    const uint2 mine = *reinterpret_cast<const uint2*>(cur + 8 * lane);
    u32 keep = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const u32 c = ((j < 4 ? mine.x : mine.y) >> (8 * (j & 3))) & 0xFFu;
        if (8 * lane + j < cnt && !((mask[c >> 5] >> (c & 31)) & 1u)) keep |= 1u << j;
    }
    u32 incl = __popc(keep);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    u32 pos = n + incl - __popc(keep);
#pragma unroll
    for (int j = 0; j < 8; ++j) if ((keep >> j) & 1u) nxt[pos++] = (u8)(((j < 4 ? mine.x : mine.y) >> (8 * (j & 3))) & 0xFFu);
    cnt = n + __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    sel ^= 1u;
}

// warp g looks at the tiles 8g .. 8g+7: every tile that starts a group (its distance from the block's first tile is a multiple of
// eight) is folded by this warp — exactly one group per warp when blocks are whole multiples of 32 KiB, a few short ones otherwise
__global__ void __launch_bounds__(256) k_mtf3_agg(const TileDesc* __restrict__ tiles, const u32* __restrict__ tile0, const u32* __restrict__ tilen,
                                                  const u8* __restrict__ slots, const u32* __restrict__ npc, u8* __restrict__ agg, u32* __restrict__ aggn, int nt) {
    __shared__ M3Warp sw[8];
    const u32 lane = lane_id(), w = threadIdx.x >> 5;
    const int g = blockIdx.x * 8 + w;
    for (int t = g * M3_GROUP_TILES; t < nt && t < (g + 1) * M3_GROUP_TILES; ++t) {
        const u32 b = tiles[t].block, rel = (u32)t - tile0[b];
        if (rel % M3_GROUP_TILES) continue;
        const u32 np = 4 * min((u32)M3_GROUP_TILES, tilen[b] - rel), p0 = 4 * (u32)t;
        const u32 myn = lane < np ? npc[p0 + lane] : 0u;
        u32 sel = 0, cnt = 0;
        u32 n = __shfl_sync(0xffffffffu, myn, 0);
        u32 l0 = lane < n ? slots[(size_t)p0 * M3_SLOT + lane] : 0u;
        for (u32 q = 0; q < np; ++q) {
            const u32 n1 = q + 1 < np ? __shfl_sync(0xffffffffu, myn, q + 1) : 0u;
            const u32 l1 = lane < n1 ? slots[(size_t)(p0 + q + 1) * M3_SLOT + lane] : 0u;      // next piece's list, in flight during this fold
            m3_combine(sw[w], sel, cnt, slots + (size_t)(p0 + q) * M3_SLOT, n, l0);
            n = n1; l0 = l1;
        }
        const u8* cur = sel ? sw[w].b : sw[w].a;
        for (u32 i = lane; i < cnt; i += 32) agg[(size_t)t * 256 + i] = cur[i];
        if (lane == 0) aggn[t] = cnt;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_mtf3_prefix(const u32* __restrict__ tile0, const u32* __restrict__ tilen, const u8* __restrict__ agg,
                                                     const u32* __restrict__ aggn, u8* __restrict__ pre, int nblocks) {
    __shared__ M3Warp sw[8];
    const u32 lane = lane_id(), w = threadIdx.x >> 5;
    const int b = blockIdx.x * 8 + w;
    if (b >= nblocks) return;
    u32 sel = 0, cnt = 256;
    for (u32 i = lane; i < 256; i += 32) sw[w].a[i] = (u8)i;
    __syncwarp();
    const u32 t0 = tile0[b], tn = tilen[b];
    u32 n = tn ? aggn[t0] : 0u;
    u32 l0 = lane < n ? agg[(size_t)t0 * 256 + lane] : 0u;
    for (u32 rel = 0; rel < tn; rel += M3_GROUP_TILES) {
        const size_t t = (size_t)t0 + rel;
        const u32 n1 = rel + M3_GROUP_TILES < tn ? aggn[t + M3_GROUP_TILES] : 0u;
        const u32 l1 = lane < n1 ? agg[(t + M3_GROUP_TILES) * 256 + lane] : 0u;
        reinterpret_cast<uint2*>(pre + t * 256)[lane] = *reinterpret_cast<const uint2*>((sel ? sw[w].b : sw[w].a) + 8 * lane);
        m3_combine(sw[w], sel, cnt, agg + t * 256, n, l0);
        n = n1; l0 = l1;
    }
}

__global__ void __launch_bounds__(256) k_mtf3_fix(u8* __restrict__ out, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                  const u32* __restrict__ tile0, const u32* __restrict__ tilen, const u8* __restrict__ slots,
                                                  const u32* __restrict__ npc, const u8* __restrict__ pre, int nt) {
    __shared__ M3Warp sw[8];
    const u32 lane = lane_id(), w = threadIdx.x >> 5;
    const int g = blockIdx.x * 8 + w;
    M3Warp& W = sw[w];
    for (int t = g * M3_GROUP_TILES; t < nt && t < (g + 1) * M3_GROUP_TILES; ++t) {
        const TileDesc td0 = tiles[t];
        const u32 b = td0.block, rel = (u32)t - tile0[b];
        if (rel % M3_GROUP_TILES) continue;
        const BlockInfo bi = binfo[b];
        const u32 np = 4 * min((u32)M3_GROUP_TILES, tilen[b] - rel), p0 = 4 * (u32)t;
        const u32 myn = lane < np ? npc[p0 + lane] : 0u;
        u32 sel = 0, cnt = 256;
        *reinterpret_cast<uint2*>(W.a + 8 * lane) = reinterpret_cast<const uint2*>(pre + (size_t)t * 256)[lane];
        __syncwarp();
        u8* dst0 = out + bi.ioff + (td0.start - bi.pbase);   // tiles of a block are consecutive and full except the last
        u32 n = __shfl_sync(0xffffffffu, myn, 0);
        u32 l0 = 0, f0 = 0, q0 = 0;
        if (lane < n) { const u8* sl = slots + (size_t)p0 * M3_SLOT; l0 = sl[lane]; f0 = sl[256 + lane]; q0 = reinterpret_cast<const u16*>(sl + 512)[lane]; }
        for (u32 q = 0; q < np; ++q) {
            const u32 n1 = q + 1 < np ? __shfl_sync(0xffffffffu, myn, q + 1) : 0u;
            u32 l1 = 0, f1 = 0, q1 = 0;                          // the next piece's first 32 entries, in flight during this step
            if (lane < n1) { const u8* sl = slots + (size_t)(p0 + q + 1) * M3_SLOT; l1 = sl[lane]; f1 = sl[256 + lane]; q1 = reinterpret_cast<const u16*>(sl + 512)[lane]; }
            if (n) {
                const u8* slot = slots + (size_t)(p0 + q) * M3_SLOT;
                const u16* Q = reinterpret_cast<const u16*>(slot + 512);
                u8* dst = dst0 + (size_t)q * (KOLM_TILE / 4);
                const uint2 mine = *reinterpret_cast<const uint2*>((sel ? W.b : W.a) + 8 * lane);
#pragma unroll
                for (int j = 0; j < 8; ++j) W.inv[((j < 4 ? mine.x : mine.y) >> (8 * (j & 3))) & 0xFFu] = (u8)(8 * lane + j);
                __syncwarp();
                if (lane < n) W.r[lane] = W.inv[f0];
                for (u32 i = lane + 32; i < n; i += 32) W.r[i] = W.inv[slot[256 + i]];
                __syncwarp();
                for (u32 i = lane; i < n; i += 32) {
                    const u32 r = W.r[i], rrrr = r * 0x01010101u;
                    u32 gt = 0, k = 0;
                    for (; k + 4 <= i; k += 4) gt += __popc(__vcmpgtu4(*reinterpret_cast<const u32*>(W.r + k), rrrr)) >> 3;
                    for (; k < i; ++k) gt += W.r[k] > r;
                    dst[i < 32 ? q0 : (u32)Q[i]] = (u8)(r + gt);
                }
                __syncwarp();
                m3_combine(W, sel, cnt, slot, n, l0);
            }
            n = n1; l0 = l1; f0 = f1; q0 = q1;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------
// Replays the list operations of one tile on `list` (256 bytes in shared memory, private to the warp).
// If out != nullptr the decoded symbols are written.
__device__ __forceinline__ void mtf_replay_tile(const u8* __restrict__ src, u8* __restrict__ dst, u32 count, u8* list, u8* res) {
    const u32 lane = threadIdx.x & 31;
    for (u32 x0 = 0; x0 < count; x0 += 32) {
        u32 x = x0 + lane;
        bool valid = x < count;
        u32 idx = valid ? src[x] : 0;
        u32 nz = __ballot_sync(0xffffffffu, idx != 0);
        u32 front0 = list[0];
        u32 todo = nz;
        while (todo) {
            u32 pos = __ffs(todo) - 1; todo &= todo - 1;
            u32 i = __shfl_sync(0xffffffffu, idx, pos);
            u32 v = list[i];
            __syncwarp();
            // shift list[0..i-1] -> list[1..i], highest chunk first so reads precede overwrites
            for (int base = (int)((i - 1) / 32) * 32; base >= 0; base -= 32) {
                u32 j = base + lane;
                u32 tmp = (j < i) ? list[j] : 0;
                __syncwarp();
                if (j < i) list[j + 1] = (u8)tmp;
                __syncwarp();
            }
            if (lane == 0) { list[0] = (u8)v; res[pos] = (u8)v; }
            __syncwarp();
        }
        if (dst && valid) {
            u32 below = nz & ((2u << lane) - 1u);               // ops at positions <= lane
            u32 sym = below ? res[31 - __clz(below)] : front0;
            dst[x] = (u8)sym;
        }
        __syncwarp();
    }
}

// pass A (decode): permutation of the tile: perm[i] = entry-list index that ends at position i
__global__ void __launch_bounds__(MTF_WARPS * 32) k_mtf_dec_perm(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                                 const BlockInfo* __restrict__ binfo, u8* __restrict__ tperm, int ntiles) {
    __shared__ u8 list[MTF_WARPS][256];
    __shared__ u8 res[MTF_WARPS][32];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * MTF_WARPS + w;
    if (tile >= ntiles) return;
    TileDesc td = tiles[tile];
    BlockInfo bi = binfo[td.block];
    for (int k = 0; k < 8; ++k) list[w][lane * 8 + k] = (u8)(lane * 8 + k);
    __syncwarp();
    mtf_replay_tile(in + bi.ioff + (td.start - bi.pbase), nullptr, td.count, list[w], res[w]);
    for (int k = 0; k < 8; ++k) tperm[(size_t)tile * 256 + lane * 8 + k] = list[w][lane * 8 + k];
}

// pass B (decode): entry list of every tile = composition of the previous tiles' permutations
__global__ void __launch_bounds__(256) k_mtf_dec_compose(u8* __restrict__ tperm, const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks, u32 ss) {
    __shared__ u8 cur[256];
    __shared__ u8 nxt[256];
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        u32 nt = tilen[b] << ss;                           // rows = pieces of 4096 >> ss bytes
        u8* base = tperm + ((size_t)tile0[b] << ss) * 256;
        cur[threadIdx.x] = (u8)threadIdx.x;
        __syncthreads();
        for (u32 t = 0; t < nt; ++t) {
            u8 p = base[(size_t)t * 256 + threadIdx.x];
            nxt[threadIdx.x] = cur[p];
            base[(size_t)t * 256 + threadIdx.x] = cur[threadIdx.x];   // entry list of tile t
            __syncthreads();
            cur[threadIdx.x] = nxt[threadIdx.x];
            __syncthreads();
        }
    }
}

// the same composition with ONE WARP per block (eight list entries per lane, warp-level synchronisation only): the CTA form above
// pays two block barriers per row, 2048 per 1 MiB block at 1 KiB pieces (0.53 ms per 256 MiB; this form 0.1)
__global__ void __launch_bounds__(128) k_mtf_dec_compose_w(u8* __restrict__ tperm, const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks, u32 ss) {
    __shared__ __align__(8) u8 cur[4][256];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x * 4 + w;
    if (b >= nblocks) return;
    const u32 nt = tilen[b] << ss;
    uint2* base = reinterpret_cast<uint2*>(tperm + ((size_t)tile0[b] << ss) * 256) + lane;
    u8* c = cur[w];
    uint2 mine = make_uint2(0x03020100u + 0x08080808u * lane, 0x07060504u + 0x08080808u * lane);    // identity entries 8*lane .. 8*lane+7
    uint2 p = nt ? base[0] : make_uint2(0, 0);
    for (u32 t = 0; t < nt; ++t) {
        const uint2 pn = t + 1 < nt ? base[(size_t)(t + 1) * 32] : make_uint2(0, 0);     // next row in flight
        reinterpret_cast<uint2*>(c)[lane] = mine;
        base[(size_t)t * 32] = mine;                       // entry list of row t
        __syncwarp();
        uint2 nx;
        nx.x = (u32)c[p.x & 0xFF] | ((u32)c[(p.x >> 8) & 0xFF] << 8) | ((u32)c[(p.x >> 16) & 0xFF] << 16) | ((u32)c[p.x >> 24] << 24);
        nx.y = (u32)c[p.y & 0xFF] | ((u32)c[(p.y >> 8) & 0xFF] << 8) | ((u32)c[(p.y >> 16) & 0xFF] << 16) | ((u32)c[p.y >> 24] << 24);
        __syncwarp();
        mine = nx; p = pn;
    }
}

// pass C (decode)
__global__ void __launch_bounds__(MTF_WARPS * 32) k_mtf_dec(const u8* __restrict__ in, u8* __restrict__ out, const TileDesc* __restrict__ tiles,
                                                            const BlockInfo* __restrict__ binfo, const u8* __restrict__ tperm, int ntiles) {
    __shared__ u8 list[MTF_WARPS][256];
    __shared__ u8 res[MTF_WARPS][32];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * MTF_WARPS + w;
    if (tile >= ntiles) return;
    TileDesc td = tiles[tile];
    BlockInfo bi = binfo[td.block];
    for (int k = 0; k < 8; ++k) list[w][lane * 8 + k] = tperm[(size_t)tile * 256 + lane * 8 + k];
    __syncwarp();
    u32 t0 = td.start - bi.pbase;
    mtf_replay_tile(in + bi.ioff + t0, out + bi.ioff + t0, td.count, list[w], res[w]);
}


// decode, v2: thread-serial replay on the packed list (same layout as k_mtf_enc2).  PERM: start from the identity list and
// store the final list (= the permutation the tile applies); otherwise start from the tile's entry list and write symbols.
template <bool PERM>
__global__ void __launch_bounds__(MTF2_THREADS) k_mtf_dec2(const u8* __restrict__ in, u8* __restrict__ out, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u8* __restrict__ tperm, int ntiles) {
    __shared__ u32 lst[64][MTF2_THREADS];
    const u32 tid = threadIdx.x;
    const int tile = blockIdx.x * MTF2_THREADS + tid;
    if (tile >= ntiles) return;
    const u32* row = reinterpret_cast<const u32*>(tperm + (size_t)tile * 256);
#pragma unroll 8
    for (u32 k = 0; k < 64; ++k) lst[k][tid] = PERM ? (0x03020100u + 0x04040404u * k) : row[k];
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase;
    const u8* src = in + bi.ioff + t0;
    u8* dst = PERM ? nullptr : out + bi.ioff + t0;
    u32 w0 = lst[0][tid];
    const bool aligned = ((uintptr_t)src & 15) == 0 && (PERM || ((uintptr_t)dst & 15) == 0);
    for (u32 x0 = 0; x0 < td.count; x0 += 16) {
        u32 inw[4], outw[4] = {0, 0, 0, 0};
        const u32 nb = min(16u, td.count - x0);
        if (aligned && nb == 16) { uint4 v = *reinterpret_cast<const uint4*>(src + x0); inw[0] = v.x; inw[1] = v.y; inw[2] = v.z; inw[3] = v.w; }
        else { inw[0] = inw[1] = inw[2] = inw[3] = 0; for (u32 i = 0; i < nb; ++i) inw[i >> 2] |= (u32)src[x0 + i] << (8 * (i & 3)); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((u32)i < nb) {
                const u32 idx = (inw[i >> 2] >> (8 * (i & 3))) & 0xFF;
                u32 sym;
                if (idx < 4) {
                    sym = (w0 >> (8 * idx)) & 0xFF;
                    if (idx) w0 = __byte_perm(w0, sym, idx == 1 ? 0x3204 : idx == 2 ? 0x3104 : 0x2104);
                } else {
                    const u32 k = idx >> 2, j = idx & 3;
                    const u32 wk = lst[k][tid];
                    sym = (wk >> (8 * j)) & 0xFF;
                    u32 carry = w0 >> 24;
                    w0 = (w0 << 8) | sym;
                    for (u32 q = 1; q < k; ++q) { u32 w = lst[q][tid]; lst[q][tid] = (w << 8) | carry; carry = w >> 24; }
                    lst[k][tid] = __byte_perm(wk, carry, j == 0 ? 0x3214 : j == 1 ? 0x3204 : j == 2 ? 0x3104 : 0x2104);
                }
                outw[i >> 2] |= sym << (8 * (i & 3));
            }
        }
        if (!PERM) {
            if (aligned && nb == 16) *reinterpret_cast<uint4*>(dst + x0) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
            else for (u32 i = 0; i < nb; ++i) dst[x0 + i] = (u8)(outw[i >> 2] >> (8 * (i & 3)));
        }
    }
    if (PERM) {
        u32* orow = reinterpret_cast<u32*>(tperm + (size_t)tile * 256);
        orow[0] = w0;
        for (u32 k = 1; k < 64; ++k) orow[k] = lst[k][tid];
    }
}


// ---------------------------------------------------------------------------------------------
// decode, v3 (default): ONE walk instead of two.  The list operations of a piece depend on the index stream only, so the
// walk from the identity list yields, besides the piece's permutation, for every byte the SLOT of the entry list its symbol
// sits in (what the identity start "decodes" to).  After the per-block composition of the permutations has produced the
// true entry lists, the symbols are a table lookup — no second serial walk.  Pieces of 1 KiB (four per tile) instead of whole
// tiles: four times the independent walks for a kernel that is bound by the latency of its serial chain.
//   k_mtf_dec3_walk   thread per piece: identity-start replay (packed list as in k_mtf_dec2), slots to `out`, permutation to tperm
//   k_mtf_dec_compose per block: entry list of every piece (tperm rewritten in place)
//   k_mtf_dec3_map    CTA per tile: out[i] = entry_list[piece(i)][out[i]]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MTF2_THREADS) k_mtf_dec3_walk(const u8* __restrict__ in, u8* __restrict__ out, const TileDesc* __restrict__ tiles,
                                                                const BlockInfo* __restrict__ binfo, u8* __restrict__ tperm, int np) {
    __shared__ u32 lst[64][MTF2_THREADS];
    const u32 tid = threadIdx.x;
    const int piece = blockIdx.x * MTF2_THREADS + tid;
    if (piece >= np) return;
#pragma unroll 8
    for (u32 k = 0; k < 64; ++k) lst[k][tid] = 0x03020100u + 0x04040404u * k;
    const TileDesc td = tiles[(u32)piece >> 2];
    const BlockInfo bi = binfo[td.block];
    u32 pstart, pcount;
    mtf_piece(td, (u32)piece, 2u, pstart, pcount);
    const u32 t0 = pstart - bi.pbase;
    const u8* src = in + bi.ioff + t0;
    u8* dst = out + bi.ioff + t0;
    u32 w0 = lst[0][tid];
    const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    for (u32 x0 = 0; x0 < pcount; x0 += 16) {
        u32 inw[4], outw[4] = {0, 0, 0, 0};
        const u32 nb = min(16u, pcount - x0);
        if (aligned && nb == 16) { uint4 v = *reinterpret_cast<const uint4*>(src + x0); inw[0] = v.x; inw[1] = v.y; inw[2] = v.z; inw[3] = v.w; }
        else { inw[0] = inw[1] = inw[2] = inw[3] = 0; for (u32 i = 0; i < nb; ++i) inw[i >> 2] |= (u32)src[x0 + i] << (8 * (i & 3)); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((u32)i < nb) {
                const u32 idx = (inw[i >> 2] >> (8 * (i & 3))) & 0xFF;
                u32 sym;
                if (idx < 4) {
                    sym = (w0 >> (8 * idx)) & 0xFF;
                    if (idx) w0 = __byte_perm(w0, sym, idx == 1 ? 0x3204 : idx == 2 ? 0x3104 : 0x2104);
                } else {
                    const u32 k = idx >> 2, j = idx & 3;
                    const u32 wk = lst[k][tid];
                    sym = (wk >> (8 * j)) & 0xFF;
                    u32 carry = w0 >> 24;
                    w0 = (w0 << 8) | sym;
                    for (u32 q = 1; q < k; ++q) { u32 w = lst[q][tid]; lst[q][tid] = (w << 8) | carry; carry = w >> 24; }
                    lst[k][tid] = __byte_perm(wk, carry, j == 0 ? 0x3214 : j == 1 ? 0x3204 : j == 2 ? 0x3104 : 0x2104);
                }
                outw[i >> 2] |= sym << (8 * (i & 3));
            }
        }
        if (aligned && nb == 16) *reinterpret_cast<uint4*>(dst + x0) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
        else for (u32 i = 0; i < nb; ++i) dst[x0 + i] = (u8)(outw[i >> 2] >> (8 * (i & 3)));
    }
    u32* orow = reinterpret_cast<u32*>(tperm + (size_t)piece * 256);
    orow[0] = w0;
    for (u32 k = 1; k < 64; ++k) orow[k] = lst[k][tid];
}

__global__ void __launch_bounds__(KOLM_THREADS) k_mtf_dec3_map(u8* __restrict__ out, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                               const u8* __restrict__ tperm) {
    __shared__ u32 E[4][64];                               // entry lists of the tile's four pieces
    const u32 tid = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    (&E[0][0])[tid] = reinterpret_cast<const u32*>(tperm + (size_t)blockIdx.x * 1024)[tid];
    __syncthreads();
    const u8* e = reinterpret_cast<const u8*>(&E[tid >> 6][0]);   // 16 bytes per thread: thread t sits in piece t / 64
    u8* dst = out + bi.ioff + (td.start - bi.pbase) + 16 * tid;
    const u32 x = 16 * tid;
    if (x >= td.count) return;
    if (x + 16 <= td.count && (((uintptr_t)dst) & 15) == 0) {
        uint4 v = *reinterpret_cast<const uint4*>(dst);
        u32 wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = (u32)e[wv[i] & 0xFF] | ((u32)e[(wv[i] >> 8) & 0xFF] << 8) | ((u32)e[(wv[i] >> 16) & 0xFF] << 16) | ((u32)e[wv[i] >> 24] << 24);
        *reinterpret_cast<uint4*>(dst) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    } else {
        const u32 nb = min(16u, td.count - x);
        for (u32 i = 0; i < nb; ++i) dst[i] = e[dst[i]];
    }
}

// KOLM_MTF3_DBG=2: host re-computation of every intermediate of the v3 encode (pieces' local lists and misses, group lists,
// group entry lists) compared with what the kernels left in scratch — a debugging aid, never on in production.
#include <vector>
static int m3_host_check(kolm_ctx* c, const u8* in, const u8* d_slots, const u32* d_npc, const u8* d_agg, const u32* d_aggn, const u8* d_pre, int stage) {
    const int nb = c->nblocks, nt = c->ntiles, np = nt * 4;
    std::vector<u8> slots((size_t)np * M3_SLOT), agg((size_t)nt * 256), pre((size_t)nt * 256);
    std::vector<u32> npc(np), aggn(nt);
    cudaMemcpy(slots.data(), d_slots, slots.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(npc.data(), d_npc, npc.size() * 4, cudaMemcpyDeviceToHost);
    if (stage >= 2) { cudaMemcpy(agg.data(), d_agg, agg.size(), cudaMemcpyDeviceToHost); cudaMemcpy(aggn.data(), d_aggn, aggn.size() * 4, cudaMemcpyDeviceToHost); }
    if (stage >= 3) cudaMemcpy(pre.data(), d_pre, pre.size(), cudaMemcpyDeviceToHost);
    int bad = 0; u32 t = 0;
    auto fold = [](std::vector<u8>& cur, const u8* L, u32 n) {
        std::vector<u8> nx(L, L + n); bool in[256] = {false};
        for (u32 i = 0; i < n; ++i) in[L[i]] = true;
        for (u8 x : cur) if (!in[x]) nx.push_back(x);
        cur.swap(nx);
    };
    for (int b = 0; b < nb && bad < 8; ++b) {
        const u32 len = c->h_binfo[b].len, tn = (len + KOLM_TILE - 1) / KOLM_TILE;
        std::vector<u8> data(len);
        if (len) cudaMemcpy(data.data(), in + c->h_binfo[b].ioff, len, cudaMemcpyDeviceToHost);
        std::vector<u8> entry(256); for (int i = 0; i < 256; ++i) entry[i] = (u8)i;
        for (u32 rel = 0; rel < tn && bad < 8; rel += M3_GROUP_TILES) {
            const u32 gtiles = std::min<u32>(M3_GROUP_TILES, tn - rel), gp = 4 * gtiles, p0 = 4 * (t + rel);
            std::vector<u8> g;
            for (u32 q = 0; q < gp; ++q) {
                const u32 a = (rel * 4 + q) * 1024, e = std::min<u32>(len, a + 1024), piece = p0 + q;
                const u8* sl = slots.data() + (size_t)piece * M3_SLOT;
                const u32 n = npc[piece];
                if (stage == 1) {
                    std::vector<u8> L, ms; std::vector<u16> mp;
                    for (u32 i = a; i < e && a < len; ++i) {
                        const u8 x = data[i]; size_t k = 0;
                        while (k < L.size() && L[k] != x) ++k;
                        if (k == L.size()) { ms.push_back(x); mp.push_back((u16)(i - a)); } else L.erase(L.begin() + k);
                        L.insert(L.begin(), x);
                    }
                    bool ok = n == L.size();
                    for (u32 i = 0; ok && i < n; ++i) ok = sl[i] == L[i] && sl[256 + i] == ms[i] && reinterpret_cast<const u16*>(sl + 512)[i] == mp[i];
                    if (!ok) { ++bad; fprintf(stderr, "mtf3 check: walk piece %u (block %d len %u, bytes %u..%u): n gpu %u cpu %zu\n", piece, b, len, a, e, n, L.size()); }
                }
                if (n <= 256) fold(g, sl, n);
            }
            const u32 gt = t + rel;
            if (stage == 2) {
                bool ok = aggn[gt] == g.size();
                for (size_t i = 0; ok && i < g.size(); ++i) ok = agg[(size_t)gt * 256 + i] == g[i];
                if (!ok) { ++bad; fprintf(stderr, "mtf3 check: agg tile %u (block %d rel %u): n gpu %u cpu %zu\n", gt, b, rel, aggn[gt], g.size()); }
            }
            if (stage == 3) {
                bool ok = true;
                for (int i = 0; ok && i < 256; ++i) ok = pre[(size_t)gt * 256 + i] == entry[i];
                if (!ok) { ++bad; fprintf(stderr, "mtf3 check: prefix tile %u (block %d rel %u)\n", gt, b, rel); }
            }
            fold(entry, g.data(), (u32)g.size());
        }
        t += tn;
    }
    fprintf(stderr, "mtf3 check stage %d: %s\n", stage, bad ? "MISMATCH" : "ok");
    return bad;
}

int kolm_mtf_impl(kolm_ctx* c, const u8* in, u8* out, bool decode, cudaStream_t s) {
    const int nt = c->ntiles, nb = c->nblocks;
    if (!nt) return KOLM_OK;
    int sgrid = nb < 4 * c->sm_count ? nb : 4 * c->sm_count;
    int wgrid = (nt + MTF_WARPS - 1) / MTF_WARPS;
    if (!decode) {
        const i64 N = c->total_bytes;
        static int v2 = -1, sub = -1, v3 = -1;
        if (v2 < 0) { const char* e = getenv("KOLM_MTF_V2"); v2 = e ? atoi(e) : 1; }
        if (v3 < 0) { const char* e = getenv("KOLM_MTF_V3"); v3 = e ? atoi(e) : 1; }
        if (v3 && v2 && (size_t)nt * 1024 <= c->max_elems) {
            // scratch in the idle sort buffers: 1 KB per piece (d_k0), the pieces' list lengths (d_v0), per group of eight tiles its
            // list (d_k1, indexed by the group's first tile), length (d_v1) and entry list (d_sa)
            const int np = nt * 4;
            u8* slots = (u8*)c->d_k0; u32* npc = c->d_v0; u8* agg = (u8*)c->d_k1; u32* aggn = c->d_v1; u8* pre = (u8*)c->d_sa;
            static int dbg = -1;
            if (dbg < 0) { const char* e = getenv("KOLM_MTF3_DBG"); dbg = e ? atoi(e) : 0; }
#define M3_CHECK(name) do { if (dbg) { cudaError_t e_ = cudaStreamSynchronize(s); if (e_ == cudaSuccess) e_ = cudaGetLastError(); \
                                        fprintf(stderr, "mtf3 %s: %s (nt %d nb %d)\n", name, cudaGetErrorString(e_), nt, nb); if (e_ != cudaSuccess) return KOLM_E_CUDA; } } while (0)
            KL(c, KC_MTF_MAIN, 2 * N, s, k_mtf3_walk<<<(np + M3_THREADS - 1) / M3_THREADS, M3_THREADS, 0, s>>>(in, out, c->d_tiles, c->d_binfo, slots, npc, np));
            M3_CHECK("walk");
            if (dbg >= 2 && m3_host_check(c, in, slots, npc, agg, aggn, pre, 1)) return KOLM_E_CUDA;
            KL(c, KC_MTF_SCAN, (i64)np * 64, s, k_mtf3_agg<<<((nt + M3_GROUP_TILES - 1) / M3_GROUP_TILES + 7) / 8, 256, 0, s>>>(c->d_tiles, c->d_btile0, c->d_btilen, slots, npc, agg, aggn, nt));
            M3_CHECK("agg");
            if (dbg >= 2 && m3_host_check(c, in, slots, npc, agg, aggn, pre, 2)) return KOLM_E_CUDA;
            KL(c, KC_MTF_SCAN, (i64)nt * 64, s, k_mtf3_prefix<<<(nb + 7) / 8, 256, 0, s>>>(c->d_btile0, c->d_btilen, agg, aggn, pre, nb));
            M3_CHECK("prefix");
            if (dbg >= 2 && m3_host_check(c, in, slots, npc, agg, aggn, pre, 3)) return KOLM_E_CUDA;
            KL(c, KC_MTF_PRE, (i64)np * 128, s, k_mtf3_fix<<<((nt + M3_GROUP_TILES - 1) / M3_GROUP_TILES + 7) / 8, 256, 0, s>>>(out, c->d_tiles, c->d_binfo, c->d_btile0, c->d_btilen, slots, npc, pre, nt));
            M3_CHECK("fix");
            CUDA_TRY(cudaGetLastError());
            return KOLM_OK;
        }
        if (sub < 0) { const char* e = getenv("KOLM_MTF_SUB"); sub = e ? atoi(e) : 2; if (sub < 0 || sub > 3) sub = 2; }
        // pieces of 4096 >> ss bytes; their last-occurrence tables (1 KB each) live in the idle sort buffer d_k0 when they
        // outgrow d_thist (batches of many tiny blocks keep whole tiles)
        u32 ss = v2 ? (u32)sub : 0u;
        if (((size_t)nt << ss) * 256 > c->max_elems) ss = 0;
        u32* tl = ss ? c->d_k0 : c->d_thist;
        const int np = nt << ss;
        // groups per block for the scan: their maxima sit in d_thist, which is free once the tables moved to d_k0
        u32 G = ss ? (u32)std::min<i64>(8, (i64)c->max_tiles / nb) : 1u;
        if (G < 1) G = 1;
        u32* gt = c->d_thist;
        if (G > 1) CUDA_TRY(cudaMemsetAsync(gt, 0, (size_t)nb * G * 1024, s));
        const int sg = (int)std::min<i64>((i64)nb * G, 16 * (i64)c->sm_count);
        KL(c, KC_MTF_PRE, N + (i64)np * 1024, s, k_mtf_last<<<np, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_btile0, c->d_btilen, tl, gt, ss, G));
        KL(c, KC_MTF_SCAN, (i64)np * 2048, s, k_mtf_scan_max<<<sg, 256, 0, s>>>(tl, c->d_btile0, c->d_btilen, gt, nb, ss, G));
        if (v2) KL(c, KC_MTF_MAIN, 2 * N + (i64)np * 1024, s, k_mtf_enc2<<<(np + MTF2_THREADS - 1) / MTF2_THREADS, MTF2_THREADS, 0, s>>>(in, out, c->d_tiles, c->d_binfo, tl, np, ss));
        else KL(c, KC_MTF_MAIN, 2 * N + (i64)nt * 1024, s, k_mtf_enc<<<wgrid, MTF_WARPS * 32, 0, s>>>(in, out, c->d_tiles, c->d_binfo, tl, nt));
    } else {
        u8* tperm = (u8*)c->d_thist;
        const i64 N = c->total_bytes;
        static int v2 = -1;
        if (v2 < 0) { const char* e = getenv("KOLM_MTF_V2"); v2 = e ? atoi(e) : 1; }
        const int g2 = (nt + MTF2_THREADS - 1) / MTF2_THREADS;
        static int v3 = -1;
        if (v3 < 0) { const char* e = getenv("KOLM_MTF_DEC_V3"); v3 = e ? atoi(e) : 1; }
        if (v3 && v2) {
            const int np = nt * 4;
            KL(c, KC_MTF_MAIN, 2 * N + (i64)np * 256, s, k_mtf_dec3_walk<<<(np + MTF2_THREADS - 1) / MTF2_THREADS, MTF2_THREADS, 0, s>>>(in, out, c->d_tiles, c->d_binfo, tperm, np));
            KL(c, KC_MTF_SCAN, (i64)np * 512, s, k_mtf_dec_compose_w<<<(nb + 3) / 4, 128, 0, s>>>(tperm, c->d_btile0, c->d_btilen, nb, 2u));
            KL(c, KC_MTF_PRE, 2 * N + (i64)np * 256, s, k_mtf_dec3_map<<<nt, KOLM_THREADS, 0, s>>>(out, c->d_tiles, c->d_binfo, tperm));
            CUDA_TRY(cudaGetLastError());
            return KOLM_OK;
        }
        if (v2) KL(c, KC_MTF_PRE, N + (i64)nt * 256, s, k_mtf_dec2<true><<<g2, MTF2_THREADS, 0, s>>>(in, nullptr, c->d_tiles, c->d_binfo, tperm, nt));
        else KL(c, KC_MTF_PRE, N + (i64)nt * 256, s, k_mtf_dec_perm<<<wgrid, MTF_WARPS * 32, 0, s>>>(in, c->d_tiles, c->d_binfo, tperm, nt));
        KL(c, KC_MTF_SCAN, (i64)nt * 512, s, k_mtf_dec_compose<<<sgrid, 256, 0, s>>>(tperm, c->d_btile0, c->d_btilen, nb, 0u));
        if (v2) KL(c, KC_MTF_MAIN, 2 * N + (i64)nt * 256, s, k_mtf_dec2<false><<<g2, MTF2_THREADS, 0, s>>>(in, out, c->d_tiles, c->d_binfo, tperm, nt));
        else KL(c, KC_MTF_MAIN, 2 * N + (i64)nt * 256, s, k_mtf_dec<<<wgrid, MTF_WARPS * 32, 0, s>>>(in, out, c->d_tiles, c->d_binfo, tperm, nt));
    }
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}
