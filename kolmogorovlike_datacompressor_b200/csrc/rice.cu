// rice.cu — entropy-coder stage.
//
//   KF model 2  (kolm_final.py:499-529 cost_gamma/cost_rice/choose_rice_grid, :636-691 pack, :762-798 unpack):
//       tokens = zero runs r>=1 (tag 0) and non-zeros x=v-1 (tag 1); k0,k1 = argmin over k in 0..6,
//       Rice vs Elias-gamma per class; stream = 2b flags, 4b k0, 4b k1, then tag + code per token.
//   V22 models 2-6 (kolm_final_researched_v2-2.py:1100-1120 bit-plane, :1413-1421 rice_encode,
//       :1650-1680 nibble/bit-reverse/Gray, :2044-2065): Rice(k=2) of every transformed MTF byte.
//
// Encode = cost pass (all candidate parameterisations in one read, per-block 64-bit sums)
//        -> device-side parameter choice + payload offsets (exclusive scan over blocks)
//        -> pack pass: per-position bit length, block-wide exclusive prefix sum with a look-back
//           across tiles, MSB-first bit scatter.
// Bitstreams are MSB-first, zero padded to a byte (BitWriter.getbytes / _BitWriter.pad_to_byte).
#include "common.cuh"

#define RB_STRIDE 64
// per-block accumulator slots (u64)
#define RB_KF_Z 0      // [0..6] rice cost of zero runs for k, [7] gamma cost
#define RB_KF_N 8      // [8..14] rice cost of non-zeros,      [15] gamma cost
#define RB_KF_NZ 16    // number of zero-run tokens
#define RB_KF_NN 17    // number of non-zero tokens
#define RB_K2 20       // [20..24] Rice(k=2) bit sums for flags {0,1,4,8,16}
#define RB_BYTES 32    // chosen payload size in bytes
#define RB_OFF 33      // exclusive byte offset of the payload
#define RB_PARAM 34    // KF: k0 | k1<<8 | urz<<16 | urn<<17

__device__ __forceinline__ u32 bitlen32(u32 v) { return 32u - __clz(v); }
__device__ __forceinline__ u8 dev_bitrev8(u32 b) { return (u8)(__brev(b) >> 24); }

// 8x8 bit transpose of one group: out byte `bit` collects bit (7-bit) of every input byte, input i -> bit (7-i)
__device__ __forceinline__ u64 bitplane8(u64 g /* byte i of the group in bits [8i, 8i+8) */) {
    u64 r = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        u32 v = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) v |= (u32)((g >> (8 * i + (7 - bit))) & 1ull) << (7 - i);
        r |= (u64)v << (8 * bit);
    }
    return r;
}
__device__ __forceinline__ u32 v22_xform(u32 b, int flags) {
    if (flags & 4) b = ((b & 0x0F) << 4) | ((b & 0xF0) >> 4);
    if (flags & 8) b = dev_bitrev8(b);
    if (flags & 16) b = (b ^ (b >> 1)) & 0xFF;
    return b;
}

// ---------------------------------------------------------------------------------------------
// MSB-first bit scatter into 32-bit words (output region must be zero).  `base` is 4-byte aligned;
// bit 0 of the stream is the MSB of byte 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_word(u32* base, u64 word, u32 be_bits) {
    if (be_bits) atomicOr(base + word, __byte_perm(be_bits, 0, 0x0123));
}
__device__ __forceinline__ void put_bits(u32* base, u64 bitpos, u32 value, u32 n) {   // n in 1..32, value < 2^n
    u64 w = bitpos >> 5; u32 o = (u32)bitpos & 31;
    u64 v = (u64)value << (64 - n - o);                   // left-aligned at bit offset o of a 64-bit window
    put_word(base, w, (u32)(v >> 32));
    put_word(base, w + 1, (u32)v);
}
__device__ __forceinline__ void put_ones(u32* base, u64 bitpos, u64 q) {
    while (q >= 32) { put_bits(base, bitpos, 0xffffffffu, 32); bitpos += 32; q -= 32; }
    if (q) put_bits(base, bitpos, (1u << q) - 1u, (u32)q);
}
__device__ __forceinline__ void put_rice(u32* base, u64 bitpos, u64 x, u32 k) {        // q ones, 0, k bits
    u64 q = x >> k;
    put_ones(base, bitpos, q);
    if (k) put_bits(base, bitpos + q + 1, (u32)(x & ((1u << k) - 1u)), k);
}
__device__ __forceinline__ void put_gamma(u32* base, u64 bitpos, u32 x) {              // (b-1) zeros then b bits
    u32 b = bitlen32(x);
    put_bits(base, bitpos + (b - 1), x, b);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory staging of one tile's bit range: codes are OR-ed into shared words, the tile's interior words are then
// stored coalesced and only the two boundary words (shared with the neighbouring tiles) use a global atomic.
// A tile whose bit range does not fit (long unary runs) scatters straight to global memory.
// ---------------------------------------------------------------------------------------------
#define STAGE_WORDS 8704
struct BitStage {
    u32* sm; u32* out; u64 first_word; u32 nwords; bool use;
    __device__ __forceinline__ void begin(u32* smem, u32* out_, u64 bit0, u64 nbits) {
        sm = smem; out = out_;
        first_word = bit0 >> 5;
        u64 nw = nbits ? ((bit0 + nbits - 1) >> 5) - first_word + 1 : 0;
        use = nw <= STAGE_WORDS; nwords = use ? (u32)nw : 0;
        for (u32 i = threadIdx.x; i < nwords; i += KOLM_THREADS) sm[i] = 0;
        __syncthreads();
    }
    __device__ __forceinline__ void word(u64 w, u32 be_bits) const {
        if (!be_bits) return;
        if (use) atomicOr(sm + (u32)(w - first_word), be_bits);
        else atomicOr(out + w, __byte_perm(be_bits, 0, 0x0123));
    }
    __device__ __forceinline__ void bits(u64 bitpos, u32 value, u32 n) const {      // n in 1..32, value < 2^n
        u64 w = bitpos >> 5; u32 o = (u32)bitpos & 31;
        if (o + n <= 32) { word(w, value << (32 - n - o)); return; }                  // inside one word: one OR
        u64 v = (u64)value << (64 - n - o);
        word(w, (u32)(v >> 32)); word(w + 1, (u32)v);
    }
    // whole tokens in one OR when they fit 32 bits: [tag] q ones, 0, k bits   /   [tag] (b-1) zeros, b bits
    __device__ __forceinline__ void rice_tok(u64 bitpos, u32 tagbits, u32 tag, u64 x, u32 k) const {
        u64 q = x >> k; u64 total = tagbits + q + 1 + k;
        if (total <= 32) {
            u32 code = (tag << (total - tagbits)) | ((((u32)1 << q) - 1u) << (k + 1)) | (u32)(x & ((1u << k) - 1u));
            bits(bitpos, code, (u32)total);
        } else { if (tagbits && tag) bits(bitpos, 1, 1); rice(bitpos + tagbits, x, k); }
    }
    __device__ __forceinline__ void gamma_tok(u64 bitpos, u32 tag, u32 x) const {      // 1 tag bit + gamma(x)
        u32 b = bitlen32(x);
        if (2 * b <= 32) bits(bitpos, (tag << (2 * b - 1)) | x, 2 * b);
        else { if (tag) bits(bitpos, 1, 1); gamma(bitpos + 1, x); }
    }
    __device__ __forceinline__ void ones(u64 bitpos, u64 q) const {
        while (q >= 32) { bits(bitpos, 0xffffffffu, 32); bitpos += 32; q -= 32; }
        if (q) bits(bitpos, (1u << q) - 1u, (u32)q);
    }
    __device__ __forceinline__ void rice(u64 bitpos, u64 x, u32 k) const {
        u64 q = x >> k; ones(bitpos, q);
        if (k) bits(bitpos + q + 1, (u32)(x & ((1u << k) - 1u)), k);
    }
    __device__ __forceinline__ void gamma(u64 bitpos, u32 x) const { u32 b = bitlen32(x); bits(bitpos + (b - 1), x, b); }
    __device__ __forceinline__ void flush() const {
        __syncthreads();
        for (u32 i = threadIdx.x; i < nwords; i += KOLM_THREADS) {
            u32 v = sm[i];
            if (!v) continue;
            v = __byte_perm(v, 0, 0x0123);
            if (i == 0 || i == nwords - 1) atomicOr(out + first_word + i, v); else out[first_word + i] = v;
        }
    }
};

// Per-thread bit accumulator over a staged tile: a thread's tokens are contiguous in the stream, so they are shifted into a
// 64-bit register and leave as whole words.  Words that lie completely inside the thread's bit range are stored plainly; only
// the first word it touches and the trailing partial word can be shared with a neighbouring thread and use an atomic OR.
struct BitAcc {
    u64 acc; u32 fill, w; bool first;
    __device__ __forceinline__ void init(const BitStage& st, u64 bp) {
        w = (u32)((bp >> 5) - st.first_word); fill = (u32)bp & 31u; acc = 0; first = true;
    }
    __device__ __forceinline__ u64 bitpos(const BitStage& st) const { return ((st.first_word + w) << 5) + fill; }
    __device__ __forceinline__ void push(const BitStage& st, u32 code, u32 n) {      // n in 1..32, code < 2^n
        acc |= (u64)code << (64u - fill - n); fill += n;
        if (fill >= 32u) {
            u32 word = (u32)(acc >> 32);
            if (first) { if (word) atomicOr(st.sm + w, word); first = false; } else st.sm[w] = word;
            ++w; acc <<= 32; fill -= 32u;
        }
    }
    __device__ __forceinline__ void finish(const BitStage& st) {
        u32 word = (u32)(acc >> 32);
        if (word) atomicOr(st.sm + w, word);
        acc = 0;
    }
};
// tokens longer than 32 bits (long unary parts, gamma of runs >= 2^16): rare, kept out of line
__device__ __noinline__ void long_rice_tok(const BitStage* st, u64 bp, u32 tagbits, u32 tag, u32 x, u32 k) { st->rice_tok(bp, tagbits, tag, x, k); }
__device__ __noinline__ void long_gamma_tok(const BitStage* st, u64 bp, u32 tag, u32 x) { st->gamma_tok(bp, tag, x); }

// ---------------------------------------------------------------------------------------------
// shared tile prologue: symbols of my IPT items plus the one after (run-end detection)
// ---------------------------------------------------------------------------------------------
#define V_END 0x100u      // beyond the end of the block
__device__ __forceinline__ void load_items(const u8* __restrict__ src, u32 t0, u32 count, u32 blen, u32 (&v)[KOLM_IPT + 1]) {
    const u32 r0 = threadIdx.x * KOLM_IPT;
    if (r0 + KOLM_IPT <= count && ((uintptr_t)(src + r0) & 15) == 0) {      // full, 16-byte aligned: one 128-bit load
        uint4 q = *reinterpret_cast<const uint4*>(src + r0);
        u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) v[i] = (w[i >> 2] >> (8 * (i & 3))) & 0xFF;
        u32 r = r0 + KOLM_IPT;
        v[KOLM_IPT] = (r < count || (r == count && t0 + r < blen)) ? (u32)src[r] : V_END;
        return;
    }
#pragma unroll
    for (int i = 0; i <= KOLM_IPT; ++i) {
        u32 r = r0 + i;
        v[i] = (r < count || (r == count && t0 + r < blen)) ? (u32)src[r] : V_END;
    }
}
// exclusive "last non-zero position (1-based, block-local)" for this thread: block scan + look-back
__device__ __forceinline__ u32 scan_last_nonzero(u32 mine, u64* lb, u32 tile, bool first, u64* s_warp, u64* s_last, u64* s_excl) {
    const u32 tid = threadIdx.x;
    u64 tot;
    u64 incl = block_scan_incl((u64)mine, 0ull, OpMax(), s_warp, &tot);
    u64 prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((tid & 31) == 31) s_last[tid >> 5] = incl;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : 0ull;
    if (tid < 32) {
        u64 e = lb_exclusive(lb, tile, first, tot, 0ull, OpMax());
        if (tid == 0) *s_excl = e;
    }
    __syncthreads();
    u32 r = (u32)max(*s_excl, prev);
    __syncthreads();
    return r;
}

// the same with the tile's exclusive value already known (recorded by the cost pass): no look-back, no ticket
__device__ __forceinline__ u32 scan_last_nonzero_known(u32 mine, u32 excl, u64* s_warp, u64* s_last) {
    const u32 tid = threadIdx.x;
    u64 tot;
    u64 incl = block_scan_incl((u64)mine, 0ull, OpMax(), s_warp, &tot);
    u64 prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((tid & 31) == 31) s_last[tid >> 5] = incl;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : 0ull;
    __syncthreads();
    return (u32)max((u64)excl, prev);
}

// per-tile scratch slots next to the 25 cost sums (tacc[tile*32 + slot])
#define RT_LN0 25      // KF: last non-zero position (1-based, block-local) before the tile
#define RT_BITOFF 26   // exclusive bit offset of the tile inside its block's token stream (k_rice_tile_offsets)
#define RT_LEAD 27     // KF: in-tile length of the zero run that ends in the tile with no non-zero before it in the tile (0: none)
#define RT_TRAIL 28    // KF: zeros at the end of the tile whose run does not end in the tile
#define RT_ANYNZ 29    // KF: the tile holds a non-zero

// ---------------------------------------------------------------------------------------------
// cost pass: every candidate parameterisation in one read of the MTF bytes
// ---------------------------------------------------------------------------------------------
template <bool KF, bool K2>
__global__ void __launch_bounds__(KOLM_THREADS, 4) k_rice_cost(const u8* __restrict__ mtf, const TileDesc* __restrict__ tiles,
                                                            const BlockInfo* __restrict__ binfo, u64* lb, u64* __restrict__ tacc) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_last[KOLM_THREADS / 32];
    __shared__ unsigned long long s_acc[25];
    __shared__ u32 s_ln0, s_lead, s_trail;
    const u32 tid = threadIdx.x;
    const u32 tile = blockIdx.x;                            // no look-back: runs that cross tiles are settled by k_rice_kf_fixup
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase;
    const u8* src = mtf + bi.ioff + t0;
    if (tid < 25) s_acc[tid] = 0;
    if (tid == 0) { s_ln0 = 0; s_lead = 0; s_trail = 0; }
    // s_ptab[x] = (x>>0) | (x>>1)<<12 | (x>>2)<<23 | (x>>3)<<33 | (x>>4)<<42 | (x>>5)<<50 | (x>>6)<<57 : field k is wide enough for
    // the sum of a thread's 16 items (x <= 254), so one 64-bit add per non-zero symbol replaces seven shift/add pairs
    __shared__ u64 s_ptab[256];
    if (KF) {
        u64 x = tid;
        s_ptab[tid] = x | ((x >> 1) << 12) | ((x >> 2) << 23) | ((x >> 3) << 33) | ((x >> 4) << 42) | ((x >> 5) << 50) | ((x >> 6) << 57);
    } else {
        // K2: Rice(k=2) length (t >> 2) + 3 <= 66 of the plain / nibble-swapped / bit-reversed / Gray byte in 11-bit fields (16 items/thread)
        const u32 b = tid;
        s_ptab[tid] = (u64)((b >> 2) + 3) | ((u64)((v22_xform(b, 4) >> 2) + 3) << 11) | ((u64)((v22_xform(b, 8) >> 2) + 3) << 22) |
                      ((u64)((v22_xform(b, 16) >> 2) + 3) << 33);
        __syncthreads();
    }
    u64 packn = 0, packz = 0, packk = 0;
    u32 v[KOLM_IPT + 1];
    load_items(src, t0, td.count, bi.len, v);
    u32 accz[8] = {0, 0, 0, 0, 0, 0, 0, 0};                  // zero runs are disjoint ranges of one block (< 2^30 bytes): warp sums fit 32 bits
    u32 accn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 nzt = 0, nnt = 0;
    u32 k2[5] = {0, 0, 0, 0, 0};
    if (KF) {
        u32 lastnz = 0;
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) { u32 r = tid * KOLM_IPT + i; if (r < td.count && v[i]) lastnz = t0 + r + 1; }
        u32 ln = scan_last_nonzero_known(lastnz, 0u, s_warp, s_last);     // last non-zero of THIS tile before my items (0: none)
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {
            u32 r = tid * KOLM_IPT + i;
            if (r < td.count) {
                u32 pos1 = t0 + r + 1;
                if (v[i]) {                                 // all seven Rice quotients of x = v-1 in one table word
                    ln = pos1;
                    packn += s_ptab[v[i] - 1];
                    accn[7] += 2 * bitlen32(v[i]) - 1;
                    ++nnt;
                } else if (v[i + 1] != 0) {                 // zero run ends here (next is non-zero or end of block)
                    if (ln == 0) {                          // it began before the tile (or at its start): length known only to the fix-up
                        s_lead = pos1 - t0;                 // at most one such run end per tile: a single writer
                        continue;
                    }
                    u32 run = pos1 - ln;
                    if (run < 256u) packz += s_ptab[run];   // same packed quotient sums as the non-zeros
                    else {
#pragma unroll
                        for (int k = 0; k < 7; ++k) accz[k] += run >> k;
                    }
                    accz[7] += 2 * bitlen32(run) - 1;
                    ++nzt;
                }
            }
        }
        {   // the thread that owns the tile's last byte describes the tile's tail (ln now covers its own items too)
            const u32 rl = td.count - 1;
            if (rl / KOLM_IPT == tid) {
                const int il = (int)(rl % KOLM_IPT);
                u32 vl = 0, vn = 0;
#pragma unroll
                for (int i = 0; i < KOLM_IPT; ++i) if (i == il) { vl = v[i]; vn = v[i + 1]; }
                s_ln0 = ln;                                 // 0: the tile holds no non-zero
                s_trail = (vl != 0 || vn != 0) ? 0u : (ln ? t0 + td.count - ln : td.count);
            }
        }
    }
    if (KF) {                                                // unpack the quotient sums, add the per-token constants 1 + k
        const int off[7] = {0, 12, 23, 33, 42, 50, 57}, wid[7] = {12, 11, 10, 9, 8, 7, 6};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            accn[k] = (u32)((packn >> off[k]) & ((1u << wid[k]) - 1u)) + nnt * (1 + k);
            accz[k] += (u32)((packz >> off[k]) & ((1u << wid[k]) - 1u)) + nzt * (1 + k);
        }
    }
    if (K2) {
#pragma unroll
        for (int gi = 0; gi < KOLM_IPT / 8; ++gi) {
            u32 r0 = tid * KOLM_IPT + gi * 8;
            if (r0 < td.count) {
                u64 g = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) if (r0 + i < td.count) g |= (u64)v[gi * 8 + i] << (8 * i);
                u64 tp = bitplane8(g);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    u32 b = (u32)(g >> (8 * i)) & 0xFF, t = (u32)(tp >> (8 * i)) & 0xFF;
                    if (r0 + i < td.count) packk += s_ptab[b];       // four variants' code lengths in one packed add
                    k2[1] += (t >> 2) + 3;                  // bit-plane variant codes the zero-padded group (V22.py:1112-1113)
                }
            }
        }
    }
    if (K2) { k2[0] = (u32)(packk & 0x7FF); k2[2] = (u32)((packk >> 11) & 0x7FF); k2[3] = (u32)((packk >> 22) & 0x7FF); k2[4] = (u32)((packk >> 33) & 0x7FF); }
    __syncthreads();
    auto red = [&](u32 x, int slot) {                        // REDUX.ADD: one instruction per warp sum
        x = __reduce_add_sync(0xffffffffu, x);
        if ((tid & 31) == 0 && x) atomicAdd(&s_acc[slot], (unsigned long long)x);
    };
    if (KF) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { red(accz[k], k); red(accn[k], 8 + k); }
        red(nzt, 16); red(nnt, 17);
    }
    if (K2) {
#pragma unroll
        for (int k = 0; k < 5; ++k) red(k2[k], 20 + k);
    }
    __syncthreads();
    // per-tile partial sums; k_tile_reduce adds them per block (same-address global atomics from 256 tiles serialise in L2)
    if (tid < 32) tacc[(size_t)tile * 32 + tid] = tid < 25 ? s_acc[tid] : !KF ? 0ull : tid == RT_LEAD ? (u64)s_lead : tid == RT_TRAIL ? (u64)s_trail :
                                                  tid == RT_ANYNZ ? (u64)(s_ln0 != 0) : 0ull;
}

// KF: zero runs that cross tile boundaries.  One thread per block walks its tiles in order with the number of pending zeros:
// the run that ends in a tile with no non-zero before it there (RT_LEAD) gets its full length and its costs are added to that
// tile's sums; every tile learns the position of the last non-zero before it (RT_LN0, what the pack pass starts from).
__global__ void k_rice_kf_fixup(u64* __restrict__ tacc, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks) {
    // one WARP per block: 32 tiles' records are loaded at once (the walk itself is a short serial chain over registers)
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const u32 t0 = tile0[b], nt = tilen[b], pbase = binfo[b].pbase;
    u32 carry = 0;
    for (u32 base = 0; base < nt; base += 32) {
        const u32 t = base + lane;
        u32 lead = 0, trail = 0, any = 0, cnt = 0, start = 0;
        u64* a = tacc + (size_t)(t0 + (t < nt ? t : 0)) * 32;
        if (t < nt) { const TileDesc td = tiles[t0 + t]; lead = (u32)a[RT_LEAD]; trail = (u32)a[RT_TRAIL]; any = a[RT_ANYNZ] != 0; cnt = td.count; start = td.start - pbase; }
        const u32 n = min(32u, nt - base);
        u32 my_carry = 0;
        for (u32 k = 0; k < n; ++k) {                       // carry entering tile base+k, broadcast from the lanes' registers
            if (lane == k) my_carry = carry;
            const u32 l = __shfl_sync(0xffffffffu, lead, k), tr = __shfl_sync(0xffffffffu, trail, k);
            const u32 an = __shfl_sync(0xffffffffu, any, k), c = __shfl_sync(0xffffffffu, cnt, k);
            carry = (!l && !an) ? carry + c : tr;
        }
        if (t < nt) {
            a[RT_LN0] = (u64)(start - my_carry);
            if (lead) {
                const u32 run = my_carry + lead;
                for (int k = 0; k < 7; ++k) a[RB_KF_Z + k] += (u64)(run >> k) + 1 + k;
                a[RB_KF_Z + 7] += 2 * bitlen32(run) - 1;
                a[RB_KF_NZ] += 1;
            }
        }
    }
}

// After the plan: every tile's bit count under the chosen parameters from its cost sums, exclusive scan over the block's tiles
// -> tacc[tile*32 + RT_BITOFF].  mode 1 = KF (bacc[RB_PARAM] holds k0, k1, Rice-vs-gamma flags), 2 = K2 (slot of the chosen variant).
__global__ void __launch_bounds__(256) k_rice_tile_offsets(u64* __restrict__ tacc, const u32* __restrict__ tile0, const u32* __restrict__ tilen,
                                                           const u64* __restrict__ bacc, int mode, int k2slot) {
    __shared__ u64 s_w[8];
    __shared__ u64 s_carry;
    const u32 b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const u32 t0 = tile0[b], nt = tilen[b];
    const u32 prm = (u32)bacc[(size_t)b * RB_STRIDE + RB_PARAM];
    const u32 k0 = prm & 0xff, k1 = (prm >> 8) & 0xff; const bool urz = (prm >> 16) & 1, urn = (prm >> 17) & 1;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (u32 base = 0; base < nt; base += 256) {
        const u32 t = base + tid;
        u64 bits = 0;
        if (t < nt) {
            const u64* a = tacc + (size_t)(t0 + t) * 32;
            bits = mode == 1 ? a[RB_KF_NZ] + a[RB_KF_NN] + (urz ? a[RB_KF_Z + k0] : a[RB_KF_Z + 7]) + (urn ? a[RB_KF_N + k1] : a[RB_KF_N + 7]) : a[RB_K2 + k2slot];
        }
        u64 v = bits;
        for (int o = 1; o < 32; o <<= 1) { u64 n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v += n; }
        if (lane == 31) s_w[w] = v;
        __syncthreads();
        u64 pre = 0;
        for (u32 i = 0; i < w; ++i) pre += s_w[i];
        const u64 carry = s_carry;
        if (t < nt) tacc[(size_t)(t0 + t) * 32 + RT_BITOFF] = carry + pre + v - bits;
        __syncthreads();
        if (tid == 255) s_carry = carry + pre + v;
        __syncthreads();
    }
}

// bacc[b*64 + slot] (+)= sum over the tiles of block b of tacc[tile*32 + slot], slot < 32
__global__ void __launch_bounds__(256) k_tile_reduce(const u64* __restrict__ tacc, const u32* __restrict__ tile0, const u32* __restrict__ tilen,
                                                     u64* __restrict__ bacc, int slot0, int nslots) {
    __shared__ u64 s[8][32];
    const u32 b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 t0 = tile0[b], nt = tilen[b];
    u64 acc = 0;
    for (u32 t = w; t < nt; t += 8) acc += tacc[(size_t)(t0 + t) * 32 + lane];
    s[w][lane] = acc;
    __syncthreads();
    if (w == 0 && (int)lane >= slot0 && (int)lane < slot0 + nslots) {
        u64 tot = 0;
        for (int i = 0; i < 8; ++i) tot += s[i][lane];
        bacc[(size_t)b * RB_STRIDE + lane] = tot;
    }
}

// single CTA: per-block parameter choice, payload size, exclusive offsets.  mode 1 = KF, 2 = K2 (slot)
__global__ void k_rice_plan(u64* __restrict__ bacc, const BlockInfo* __restrict__ binfo, i64* __restrict__ poff, int* __restrict__ params,
                            i64* __restrict__ sizes5, int nblocks, int mode, int k2slot) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        int b = base + threadIdx.x;
        u64 bytes = 0;
        if (b < nblocks) {
            u64* a = bacc + (size_t)b * RB_STRIDE;
            if (mode == 1) {
                int k0 = 0, k1 = 0; u64 c0 = 0, c1 = 0;
                if (a[RB_KF_NZ]) { c0 = a[RB_KF_Z]; for (int k = 1; k < 7; ++k) if (a[RB_KF_Z + k] < c0) { c0 = a[RB_KF_Z + k]; k0 = k; } }
                if (a[RB_KF_NN]) { c1 = a[RB_KF_N]; for (int k = 1; k < 7; ++k) if (a[RB_KF_N + k] < c1) { c1 = a[RB_KF_N + k]; k1 = k; } }
                u64 g0 = a[RB_KF_Z + 7], g1 = a[RB_KF_N + 7];
                int urz = c0 < g0, urn = c1 < g1;
                u64 bits = 10 + a[RB_KF_NZ] + a[RB_KF_NN] + (urz ? c0 : g0) + (urn ? c1 : g1);
                bytes = (bits + 7) >> 3;
                a[RB_PARAM] = (u64)k0 | ((u64)k1 << 8) | ((u64)urz << 16) | ((u64)urn << 17);
                params[4 * b] = k0; params[4 * b + 1] = k1; params[4 * b + 2] = urz; params[4 * b + 3] = urn;
            } else {
                bytes = (a[RB_K2 + k2slot] + 7) >> 3;
                for (int k = 0; k < 5; ++k) sizes5[5 * (size_t)b + k] = (i64)((a[RB_K2 + k] + 7) >> 3);
            }
            a[RB_BYTES] = bytes;
        }
        u64 v = bytes;
        for (int o = 1; o < 32; o <<= 1) { u64 n = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (u32)o) v += n; }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = v;
        __syncthreads();
        u64 pre = 0;
        for (u32 i = 0; i < (threadIdx.x >> 5); ++i) pre += s_w[i];
        u64 carry = s_carry;
        if (b < nblocks) { poff[b] = (i64)(carry + pre + v - bytes); bacc[(size_t)b * RB_STRIDE + RB_OFF] = carry + pre + v - bytes; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) poff[nblocks] = (i64)s_carry;
}

__global__ void k_zero_words(u32* __restrict__ out, const i64* __restrict__ total_bytes, size_t cap_words) {
    size_t n = ((size_t)*total_bytes + 3) / 4 + 1;
    if (n > cap_words) n = cap_words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// KF pack (KF.py:662-684)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KOLM_THREADS, 4) k_rice_kf_pack(const u8* __restrict__ mtf, const TileDesc* __restrict__ tiles,
                                                               const BlockInfo* __restrict__ binfo, const u64* __restrict__ tacc, const u64* __restrict__ bacc,
                                                               u32* __restrict__ out, const i64* __restrict__ cap_total, u64 cap) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_last[KOLM_THREADS / 32];
    if ((u64)*cap_total + 8 > cap) return;                  // exact total known before any bit is packed (words are written whole: 8 bytes of slack)
    const u32 tid = threadIdx.x;
    const u32 tile = blockIdx.x;                            // no ticket, no look-back: entry state and bit offset were recorded per tile
    const u64* trec = tacc + (size_t)tile * 32;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u64* a = bacc + (size_t)td.block * RB_STRIDE;
    const u32 prm = (u32)a[RB_PARAM];
    const u32 k0 = prm & 0xff, k1 = (prm >> 8) & 0xff; const bool urz = (prm >> 16) & 1, urn = (prm >> 17) & 1;
    const u64 bitbase = a[RB_OFF] * 8;
    const u32 t0 = td.start - bi.pbase;
    u32 v[KOLM_IPT + 1];
    load_items(mtf + bi.ioff + t0, t0, td.count, bi.len, v);
    u32 lastnz = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { u32 r = tid * KOLM_IPT + i; if (r < td.count && v[i]) lastnz = t0 + r + 1; }
    const u32 ln0 = scan_last_nonzero_known(lastnz, (u32)trec[RT_LN0], s_warp, s_last);
    // token of item i (KF.py:670-684): tag bit, then Rice(k) of x or gamma of x.  non-zero v -> tag 1, x = v-1 (Rice) / v (gamma);
    // a zero that ends a run -> tag 0, x = run length.  The parameters are uniform over the block, so every token with a
    // value below 256 (all non-zeros, almost all runs) comes from one of two 256-entry tables built per CTA:
    // s_tok[class][value] = code | length << 32; length 0 = "does not fit 32 bits" -> the out-of-line path.
    __shared__ u64 s_tok[2][256];
    {
        const u32 t = tid;
#pragma unroll
        for (int cls = 0; cls < 2; ++cls) {
            const bool rice = cls ? urn : urz; const u32 k = cls ? k1 : k0, tag = (u32)cls;
            const u32 x = cls ? t - (urn ? 1u : 0u) : t;       // class 1 is indexed by the symbol v, class 0 by the run length
            u32 n = 0, code = 0;
            if (t) {
                if (rice) { const u32 q = x >> k; n = q + 2 + k; if (n <= 32) code = (tag << (n - 1)) | ((((u32)1 << q) - 1u) << (k + 1)) | (x & ((1u << k) - 1u)); }
                else { n = 2 * bitlen32(x); code = (tag << ((n - 1) & 31)) | x; }
                if (n > 32) n = 0;
            }
            s_tok[cls][t] = (u64)code | ((u64)n << 32);
        }
    }
    __syncthreads();
    auto exact_len = [&](bool nz, u32 idx) -> u32 {            // tokens outside the tables
        const bool rice = nz ? urn : urz; const u32 k = nz ? k1 : k0;
        const u32 x = nz ? idx - (urn ? 1u : 0u) : idx;
        return rice ? (x >> k) + 2 + k : 2 * bitlen32(x);
    };
    u64 mybits = 0; u32 ln = ln0;
    u32 tokmask = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 r = tid * KOLM_IPT + i;
        if (r < td.count) {
            u32 pos1 = t0 + r + 1;
            const bool nz = v[i] != 0, tok = nz || v[i + 1] != 0;
            const u32 idx = nz ? v[i] : pos1 - ln;             // the symbol, or the length of the zero run ending here
            if (nz) ln = pos1;
            if (tok) {
                u32 n = idx < 256u ? (u32)(s_tok[nz ? 1 : 0][idx] >> 32) : 0u;
                if (n == 0) n = exact_len(nz, idx);
                mybits += n;
                tokmask |= 1u << i;
                v[i] = idx | (nz ? 0x80000000u : 0u);
            }
        }
    }
    u64 btot;
    u64 bincl = block_scan_incl(mybits, 0ull, OpAdd(), s_warp, &btot);
    const u64 texcl = trec[RT_BITOFF];
    const u64 bp0 = bitbase + 10 + texcl + (bincl - mybits);
    __shared__ u32 s_stage[STAGE_WORDS];
    BitStage st;
    {   // this tile's bit range (the first tile of a block also owns the 10 header bits)
        u64 tb0 = bitbase + 10 + texcl, tbn = btot;
        if (td.flags & 1u) { tb0 = bitbase; tbn += 10; }
        st.begin(s_stage, out, tb0, tbn);
    }
    if ((td.flags & 1u) && tid == 0) st.bits(bitbase, (((urn ? 2u : 0u) | (urz ? 1u : 0u)) << 8) | (k0 << 4) | k1, 10);   // KF.py:664-668
    {   // a tile whose bit range is too long for the stage (st.use false) sends every token down the out-of-line path
        BitAcc ba; ba.init(st, bp0);
#pragma unroll
        for (int i = 0; i < KOLM_IPT; ++i) {
            if ((tokmask >> i) & 1u) {
                const bool nz = v[i] >> 31; const u32 idx = v[i] & 0x7fffffffu;
                const u64 e = idx < 256u ? s_tok[nz ? 1 : 0][idx] : 0ull;
                const u32 n = (u32)(e >> 32);
                if (n && st.use) ba.push(st, (u32)e, n);
                else {
                    const bool rice = nz ? urn : urz; const u32 k = nz ? k1 : k0, tag = nz ? 1u : 0u;
                    const u32 x = nz ? idx - (urn ? 1u : 0u) : idx;
                    ba.finish(st);
                    u64 bp = ba.bitpos(st);
                    if (rice) long_rice_tok(&st, bp, 1, tag, x, k); else long_gamma_tok(&st, bp, tag, x);
                    ba.init(st, bp + exact_len(nz, idx));
                }
            }
        }
        ba.finish(st);
    }
    st.flush();
}

// ---------------------------------------------------------------------------------------------
// V22 Rice(k=2) pack of T_flags(mtf)  (V22.py:1413-1421)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KOLM_THREADS, 4) k_rice_k2_pack(const u8* __restrict__ mtf, const TileDesc* __restrict__ tiles,
                                                               const BlockInfo* __restrict__ binfo, const u64* __restrict__ tacc, const u64* __restrict__ bacc,
                                                               u32* __restrict__ out, int flags, const i64* __restrict__ cap_total, u64 cap) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    if ((u64)*cap_total + 8 > cap) return;                  // exact total known before any bit is packed (words are written whole: 8 bytes of slack)
    const u32 tid = threadIdx.x;
    const u32 tile = blockIdx.x;                            // bit offset of the tile from k_rice_tile_offsets: no look-back
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u64 bitbase = bacc[(size_t)td.block * RB_STRIDE + RB_OFF] * 8;
    const u32 t0 = td.start - bi.pbase;
    u32 v[KOLM_IPT + 1];
    load_items(mtf + bi.ioff + t0, t0, td.count, bi.len, v);
    u32 sym[KOLM_IPT]; u32 nsym_mask = 0;
    u64 mybits = 0;
#pragma unroll
    for (int gi = 0; gi < KOLM_IPT / 8; ++gi) {
        u32 r0 = tid * KOLM_IPT + gi * 8;
        u64 g = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (r0 + i < td.count) g |= (u64)v[gi * 8 + i] << (8 * i);
        if (flags & 1) g = bitplane8(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bool coded = (flags & 1) ? (r0 < td.count) : (r0 + i < td.count);
            u32 t = v22_xform((u32)(g >> (8 * i)) & 0xFF, flags);
            sym[gi * 8 + i] = t;
            if (coded) { nsym_mask |= 1u << (gi * 8 + i); mybits += (t >> 2) + 3; }
        }
    }
    u64 btot;
    u64 bincl = block_scan_incl(mybits, 0ull, OpAdd(), s_warp, &btot);
    const u64 texcl = tacc[(size_t)tile * 32 + RT_BITOFF];
    const u64 bp0 = bitbase + texcl + (bincl - mybits);
    __shared__ u32 s_stage[STAGE_WORDS];
    BitStage st;
    st.begin(s_stage, out, bitbase + texcl, btot);
    BitAcc ba; ba.init(st, bp0);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if ((nsym_mask >> i) & 1u) {
            const u32 t = sym[i], q = t >> 2, n = q + 3;     // q ones, the terminating 0 and the 2 remainder bits
            if (n <= 32 && st.use) ba.push(st, ((((u32)1 << q) - 1u) << 3) | (t & 3u), n);
            else {
                ba.finish(st);
                u64 bp = ba.bitpos(st);
                long_rice_tok(&st, bp, 0, 0, t, 2);
                ba.init(st, bp + n);
            }
        }
    }
    ba.finish(st);
    st.flush();
}

// ---------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------
static int rice_finish(kolm_ctx* c, i64* out_off, int* params, i64* sizes, size_t out_cap, cudaStream_t s) {
    const int nb = c->nblocks;
    CUDA_TRY(cudaMemcpyAsync(c->h_poff, c->d_poff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (params) CUDA_TRY(cudaMemcpyAsync(c->h_params, c->d_params, (size_t)nb * 16, cudaMemcpyDeviceToHost, s));
    if (sizes) CUDA_TRY(cudaMemcpyAsync(c->h_sizes, c->d_sizes, (size_t)nb * 40, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_off, c->h_poff, (size_t)(nb + 1) * 8);
    if (params) memcpy(params, c->h_params, (size_t)nb * 16);
    if (sizes) memcpy(sizes, c->h_sizes, (size_t)nb * 40);
    if ((size_t)out_off[nb] + 8 > out_cap) return KOLM_E_CAPACITY;   // the pack kernels did not run (device-side guard)
    c->algbytes[KC_RICE_PACK] += out_off[nb];                 // payload bytes written
    return KOLM_OK;
}

int kolm_rice_kf_enc_impl(kolm_ctx* c, const u8* mtf, u8* out, size_t out_cap, i64* out_off, int* params, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    if (((uintptr_t)out & 3) != 0) return KOLM_E_ARG;
    if (!nb) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * RB_STRIDE * 8, s));
    if (nt) {
        KL(c, KC_RICE_COST, c->total_bytes, s, k_rice_cost<true, false><<<nt, KOLM_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, c->d_lb, (u64*)c->d_thist));
        KL(c, KC_RICE_COST, (i64)nt * 64, s, k_rice_kf_fixup<<<(nb + 3) / 4, 128, 0, s>>>((u64*)c->d_thist, c->d_tiles, c->d_binfo, c->d_btile0, c->d_btilen, nb));
        KL(c, KC_RICE_COST, (i64)nt * 256, s, k_tile_reduce<<<nb, 256, 0, s>>>((const u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 0, 20));
    }
    KL(c, KC_RICE_PLAN, (i64)nb * 256, s, k_rice_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_binfo, c->d_poff, c->d_params, c->d_sizes, nb, 1, 0));
    KL(c, KC_ZERO, 0, s, k_zero_words<<<4 * c->sm_count, 256, 0, s>>>((u32*)out, c->d_poff + nb, out_cap / 4));
    if (nt) {
        KL(c, KC_RICE_PLAN, (i64)nt * 64, s, k_rice_tile_offsets<<<nb, 256, 0, s>>>((u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 1, 0));
        KL(c, KC_RICE_PACK, c->total_bytes, s, k_rice_kf_pack<<<nt, KOLM_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, (const u64*)c->d_thist, c->d_bacc, (u32*)out, c->d_poff + nb, (u64)out_cap));
    }
    CUDA_TRY(cudaGetLastError());
    return rice_finish(c, out_off, params, nullptr, out_cap, s);
}

static int k2_slot(int flags) { switch (flags) { case 0: return 0; case 1: return 1; case 4: return 2; case 8: return 3; case 16: return 4; } return -1; }

int kolm_rice_k2_enc_impl(kolm_ctx* c, const u8* mtf, int flags, u8* out, size_t out_cap, i64* out_off, i64* sizes, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    int slot = k2_slot(flags);
    if (slot < 0 || ((uintptr_t)out & 3) != 0) return KOLM_E_ARG;
    if (!nb) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * RB_STRIDE * 8, s));
    if (nt) {
        KL(c, KC_RICE_COST, c->total_bytes, s, k_rice_cost<false, true><<<nt, KOLM_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, c->d_lb, (u64*)c->d_thist));
        KL(c, KC_RICE_COST, (i64)nt * 256, s, k_tile_reduce<<<nb, 256, 0, s>>>((const u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 20, 5));
    }
    KL(c, KC_RICE_PLAN, (i64)nb * 256, s, k_rice_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_binfo, c->d_poff, c->d_params, c->d_sizes, nb, 2, slot));
    KL(c, KC_ZERO, 0, s, k_zero_words<<<4 * c->sm_count, 256, 0, s>>>((u32*)out, c->d_poff + nb, out_cap / 4));
    if (nt) {
        KL(c, KC_RICE_PLAN, (i64)nt * 64, s, k_rice_tile_offsets<<<nb, 256, 0, s>>>((u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 2, slot));
        KL(c, KC_RICE_PACK, c->total_bytes, s, k_rice_k2_pack<<<nt, KOLM_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, (const u64*)c->d_thist, c->d_bacc, (u32*)out, flags, c->d_poff + nb, (u64)out_cap));
    }
    CUDA_TRY(cudaGetLastError());
    return rice_finish(c, out_off, nullptr, sizes, out_cap, s);
}

// ---------------------------------------------------------------------------------------------
// decode (v0): one thread walks one block's bitstream; zero runs only advance the cursor because
// the output is pre-zeroed.  Mirrors BitReader (KF.py:455-497) / rice_decode (V22.py:1423-1452).
// ---------------------------------------------------------------------------------------------
struct DBits {
    const u8* p; u64 n; u64 pos; bool eof;
    __device__ __forceinline__ u32 bit() {
        if (pos >= n) { eof = true; return 0; }
        u32 b = (p[pos >> 3] >> (7 - (pos & 7))) & 1u; ++pos; return b;
    }
    __device__ __forceinline__ u32 bits(int k) { u32 v = 0; for (int i = 0; i < k; ++i) v = (v << 1) | bit(); return v; }
    __device__ __forceinline__ u64 run_of(u32 want) {            // consecutive `want` bits, terminator consumed
        u64 q = 0;
        const u8 full = want ? 0xFF : 0x00;
        for (;;) {
            if (pos >= n) { eof = true; return q; }
            if ((pos & 7) == 0 && pos + 8 <= n && p[pos >> 3] == full) { q += 8; pos += 8; continue; }
            if (bit() == want) ++q; else return q;
        }
    }
};

__global__ void k_rice_kf_dec(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo,
                              u8* __restrict__ out, int* __restrict__ err, int nblocks) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    BlockInfo bi = binfo[b];
    DBits br; br.p = pay + pay_off[b]; br.n = (u64)(pay_off[b + 1] - pay_off[b]) * 8; br.pos = 0; br.eof = false;
    u8* dst = out + bi.ioff;
    int e = KOLM_OK;
    if (bi.len) {
        u32 flags = br.bits(2), k0 = br.bits(4), k1 = br.bits(4);
        bool urz = flags & 1u, urn = (flags >> 1) != 0;
        u64 o = 0;
        while (o < bi.len) {
            u32 tag = br.bit();
            if (br.eof) { e = KOLM_E_TRUNCATED; break; }
            if (tag == 0) {
                u64 run;
                if (urz) { u64 q = br.run_of(1); u32 r = k0 ? br.bits(k0) : 0; run = (q << k0) | r; }
                else { u64 z = br.run_of(0); run = z ? ((1ull << z) | br.bits((int)(z > 31 ? 31 : z))) : 1; if (z > 31) e = KOLM_E_CORRUPT; }
                if (br.eof) { e = KOLM_E_TRUNCATED; break; }
                o += run;
            } else {
                u64 v;
                if (urn) { u64 q = br.run_of(1); u32 r = k1 ? br.bits(k1) : 0; v = (q << k1) | r; }
                else { u64 z = br.run_of(0); v = (z ? ((1ull << z) | br.bits((int)(z > 31 ? 31 : z))) : 1) - 1; }
                if (br.eof) { e = KOLM_E_TRUNCATED; break; }
                if (v + 1 > 255) { e = KOLM_E_INDEX; break; }
                dst[o++] = (u8)(v + 1);
            }
            if (e) break;
        }
    }
    err[b] = e;
}

__global__ void k_rice_k2_dec(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo,
                              u8* __restrict__ out, int* __restrict__ err, int nblocks, int flags) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    BlockInfo bi = binfo[b];
    DBits br; br.p = pay + pay_off[b]; br.n = (u64)(pay_off[b + 1] - pay_off[b]) * 8; br.pos = 0; br.eof = false;
    u8* dst = out + bi.ioff;
    int e = KOLM_OK;
    if ((flags & 1) && (bi.len % 8)) e = KOLM_E_INDEX;           // reference: bitplane_deinterleave IndexError
    for (u32 o = 0; o < bi.len && !e; ++o) {
        u64 q = br.run_of(1);
        if (br.eof) { e = KOLM_E_TRUNCATED; break; }
        if (br.pos + 2 > br.n) { e = KOLM_E_TRUNCATED; break; }
        u32 r = br.bits(2);
        u64 v = q * 4 + r;
        if (v > 255) { e = KOLM_E_CORRUPT; break; }               // reference: bytes(seq) ValueError
        u32 t = (u32)v;
        if (flags & 16) { t ^= t >> 1; t ^= t >> 2; t ^= t >> 4; t &= 0xFF; }
        if (flags & 8) t = dev_bitrev8(t);
        if (flags & 4) t = ((t & 0x0F) << 4) | ((t & 0xF0) >> 4);
        dst[o] = (u8)t;
    }
    if (!e && (flags & 1)) {                                       // inverse 8x8 transpose per group (self-inverse)
        for (u32 g0 = 0; g0 < bi.len; g0 += 8) {
            u64 g = 0;
            for (int i = 0; i < 8; ++i) g |= (u64)dst[g0 + i] << (8 * i);
            u64 t = bitplane8(g);
            for (int i = 0; i < 8; ++i) dst[g0 + i] = (u8)(t >> (8 * i));
        }
    }
    err[b] = e;
}

static int rice_dec_finish(kolm_ctx* c, cudaStream_t s) {
    CUDA_TRY(cudaMemcpyAsync(c->h_err, c->d_err, (size_t)c->nblocks * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < c->nblocks; ++b) if (c->h_err[b]) return c->h_err[b];
    return KOLM_OK;
}

