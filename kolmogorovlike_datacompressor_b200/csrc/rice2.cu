// rice2.cu — entropy-coder stage, warp-per-tile form (same streams as rice.cu's first version).
//
//   KF model 2  (kolm_final.py:499-529 cost_gamma / cost_rice / choose_rice_grid, :636-691 pack)
//   V22 models 2-6 (kolm_final_researched_v2-2.py:1100-1120, 1413-1421, 1650-1680, 2044-2065)
//
// The first version ran one 256-thread CTA per 4 KiB tile with four block-wide barriers and two 2 KB tables built per CTA:
// ncu showed both passes latency bound (a few µs of dependent set-up per tile, DRAM idle).  Here ONE WARP owns a tile and walks it
// in eight steps of 512 bytes (16 per lane, one 128-bit load each), so every scan is a handful of shuffles, there is no block
// barrier, and the per-tile state (last non-zero, bit offset) is carried in registers:
//   k_rice2_cost   all seven Rice parameters + gamma for both KF token classes AND the five V22 variants from ONE read; the V22
//                  sums are SIMD-in-register: sum (T(b) >> 2) per 32-bit word with DP4A; the bit-plane variant needs no transpose
//                  for its COST: sum_j (t_j >> 2) over a transposed 8-byte group = sum_{i<6} popc(in_i) << (5 - i).
//   k_rice2_fixup  zero runs that cross tile boundaries (one warp per block over the tile records).
//   k_rice2_kf_pack / k_rice2_k2_pack   token codes by arithmetic (no tables), bit offsets by warp scans, each 512-byte step staged
//                  in a per-warp shared-memory window and written back as whole words.
// Zero-run tokens are attributed "look-behind": a run is costed and emitted at the non-zero that follows it (or at the end of the
// block), so no thread needs the byte after its own.
#include "common.cuh"

#define R2_WARPS 8
#define R2_THREADS (R2_WARPS * 32)
#define R2_STEP 512u               // bytes per warp step
#define R2_STAGE 1088u             // words of the per-warp staging window: 512 bytes x 66 bits (Rice(k=2) of 255) + slack

// 16 bytes at p (any alignment) as four little-endian words; only the first `valid` bytes are defined, the rest read as zero.
// `after` = bytes of the same tile behind my 16 (the unaligned path reads up to four of them).
__device__ __forceinline__ void r2_load16(const u8* __restrict__ p, u32 valid, u32 after, u32 (&w)[4]) {
    if (valid == 16) {
        const uintptr_t a = (uintptr_t)p;
        if ((a & 15) == 0) { const uint4 q = __ldg(reinterpret_cast<const uint4*>(p)); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; return; }
        const u32 sh = (u32)(a & 3) * 8;
        if (sh == 0) {
            const u32* q = reinterpret_cast<const u32*>(p);
            w[0] = __ldg(q); w[1] = __ldg(q + 1); w[2] = __ldg(q + 2); w[3] = __ldg(q + 3);
            return;
        }
        if (after >= 4) {                                   // five aligned words around my bytes, funnel-shifted into place
            const u32* q = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
            const u32 x0 = __ldg(q), x1 = __ldg(q + 1), x2 = __ldg(q + 2), x3 = __ldg(q + 3), x4 = __ldg(q + 4);
            w[0] = __funnelshift_r(x0, x1, sh); w[1] = __funnelshift_r(x1, x2, sh); w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
            return;
        }
    }
    w[0] = w[1] = w[2] = w[3] = 0;
    for (u32 i = 0; i < valid; ++i) w[i >> 2] |= (u32)p[i] << (8 * (i & 3));
}
// bit i of the result = byte i of the four words is non-zero
__device__ __forceinline__ u32 r2_nzmask(const u32 (&w)[4]) {
    u32 m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const u32 t = __vcmpne4(w[j], 0u) & 0x08040201u;     // 0xFF per non-zero byte -> one distinct bit per byte
        m |= ((t * 0x01010101u) >> 24) << (4 * j);           // the four bits meet in the top byte
    }
    return m;
}
__device__ __forceinline__ u32 r2_byte(const u32 (&w)[4], u32 i) {   // byte i (0..15) without dynamic register indexing
    const u32 lo = (i & 8) ? w[2] : w[0], hi = (i & 8) ? w[3] : w[1];
    const u32 x = (i & 4) ? hi : lo;
    return (x >> (8 * (i & 3))) & 0xFFu;
}
__device__ __forceinline__ u32 r2_warp_incl_max(u32 v) {
    const u32 lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v = max(v, n); }
    return v;
}
__device__ __forceinline__ u64 r2_warp_incl_add(u64 v) {
    const u32 lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v += n; }
    return v;
}

// ---------------------------------------------------------------------------------------------
// cost pass
// ---------------------------------------------------------------------------------------------
template <bool KF, bool K2>
__global__ void __launch_bounds__(R2_THREADS) k_rice2_cost(const u8* __restrict__ mtf, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u64* __restrict__ tacc, int ntiles) {
    // s_ptab[x] = (x>>0) | (x>>1)<<12 | (x>>2)<<23 | (x>>3)<<33 | (x>>4)<<42 | (x>>5)<<50 | (x>>6)<<57: field k holds the sum of a
    // lane's <= 16 tokens of one step, so one 64-bit add per token replaces seven shift/add pairs
    __shared__ u64 s_ptab[256];
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (KF) {
        const u64 x = tid;
        s_ptab[tid] = x | ((x >> 1) << 12) | ((x >> 2) << 23) | ((x >> 3) << 33) | ((x >> 4) << 42) | ((x >> 5) << 50) | ((x >> 6) << 57);
        __syncthreads();
    }
    const int tile = blockIdx.x * R2_WARPS + w;
    if (tile >= ntiles) return;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase, count = td.count;
    const u8* src = mtf + bi.ioff + t0;
    const bool last_tile = (td.flags & 2u) != 0;
    u32 accz[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accn[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nzt = 0, nnt = 0;
    u32 k2[5] = {0, 0, 0, 0, 0};
    u32 carry = 0;                                            // 1-based in-tile index of the last non-zero before this step (warp uniform)
    u32 myfirst = 0xffffffffu;                                // in-tile index of my first non-zero
    const int off7[7] = {0, 12, 23, 33, 42, 50, 57}, wid7[7] = {12, 11, 10, 9, 8, 7, 6};
    u32 wcur[4];
    r2_load16(src + lane * 16, min(16u, count > lane * 16 ? count - lane * 16 : 0u), count > lane * 16 + 16 ? count - lane * 16 - 16 : 0u, wcur);
    for (u32 base = 0; base < count; base += R2_STEP) {
        const u32 o = base + lane * 16;
        u32 wd[4] = {wcur[0], wcur[1], wcur[2], wcur[3]};
        {   // next step's bytes are requested before this step's arithmetic
            const u32 on = o + R2_STEP;
            if (base + R2_STEP < count) r2_load16(src + on, min(16u, count > on ? count - on : 0u), count > on + 16 ? count - on - 16 : 0u, wcur);
        }
        const u32 valid = count > o ? min(16u, count - o) : 0u;
        if (K2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const u32 x = wd[j];
                k2[0] = __dp4a((x >> 2) & 0x3F3F3F3Fu, 0x01010101u, k2[0]);
                k2[2] = __dp4a(((x & 0x0F0F0F0Fu) << 2) | ((x >> 6) & 0x03030303u), 0x01010101u, k2[2]);
                k2[3] = __dp4a((__brev(x) >> 2) & 0x3F3F3F3Fu, 0x01010101u, k2[3]);
                const u32 g = x ^ ((x >> 1) & 0x7F7F7F7Fu);
                k2[4] = __dp4a((g >> 2) & 0x3F3F3F3Fu, 0x01010101u, k2[4]);
                u32 p = x - ((x >> 1) & 0x55555555u);
                p = (p & 0x33333333u) + ((p >> 2) & 0x33333333u);
                p = (p + (p >> 4)) & 0x0F0F0F0Fu;            // popcount of every byte
                k2[1] = __dp4a(p, (j & 1) ? 0x00000102u : 0x04081020u, k2[1]);   // group byte i weighs 2^(5-i), i < 6
            }
        }
        if (KF) {
            u32 nzm = r2_nzmask(wd);
            const u32 mylast1 = nzm ? o + (31u - __clz(nzm)) + 1u : 0u;
            const u32 incl = r2_warp_incl_max(mylast1);
            u32 ln = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) ln = 0;
            ln = max(ln, carry);
            carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
            if (nzm && myfirst == 0xffffffffu) myfirst = o + (__ffs(nzm) - 1);
            u64 packn = 0, packz = 0;
            while (nzm) {
                const u32 i = __ffs(nzm) - 1; nzm &= nzm - 1;
                const u32 b = r2_byte(wd, i), pos = o + i;
                if (ln) {                                     // the run before the tile's FIRST non-zero belongs to the fix-up
                    const u32 run = pos - ln;
                    if (run) {
                        if (run < 256u) packz += s_ptab[run];
                        else {
#pragma unroll
                            for (int k = 0; k < 7; ++k) accz[k] += run >> k;
                        }
                        accz[7] += 2 * (32u - __clz(run)) - 1; ++nzt;
                    }
                }
                packn += s_ptab[b - 1];
                accn[7] += 2 * (32u - __clz(b)) - 1; ++nnt;
                ln = pos + 1;
            }
            if (last_tile && valid && o + valid == count && ln && count > ln) {   // zeros up to the end of the block: a run that ends here
                const u32 run = count - ln;
                if (run < 256u) packz += s_ptab[run];
                else {
#pragma unroll
                    for (int k = 0; k < 7; ++k) accz[k] += run >> k;
                }
                accz[7] += 2 * (32u - __clz(run)) - 1; ++nzt;
            }
            if (packn | packz) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    accn[k] += (u32)(packn >> off7[k]) & ((1u << wid7[k]) - 1u);
                    accz[k] += (u32)(packz >> off7[k]) & ((1u << wid7[k]) - 1u);
                }
            }
        }
    }
    // warp sums -> the tile's record, slot s written by lane s
    u64 mine = 0;
    if (KF) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const u32 cz = k < 7 ? nzt * (1 + k) : 0u, cn = k < 7 ? nnt * (1 + k) : 0u;   // per-token constants 1 + k of Rice(k)
            const u32 z = __reduce_add_sync(0xffffffffu, accz[k] + cz), n = __reduce_add_sync(0xffffffffu, accn[k] + cn);
            if (lane == (u32)(RB_KF_Z + k)) mine = z;
            if (lane == (u32)(RB_KF_N + k)) mine = n;
        }
        const u32 tz = __reduce_add_sync(0xffffffffu, nzt), tn = __reduce_add_sync(0xffffffffu, nnt);
        if (lane == RB_KF_NZ) mine = tz;
        if (lane == RB_KF_NN) mine = tn;
        const u32 first = __reduce_min_sync(0xffffffffu, myfirst);
        const bool any = carry != 0;
        const u32 lead = any ? first : (last_tile ? count : 0u);
        const u32 trail = last_tile ? 0u : (any ? count - carry : count);
        if (lane == RT_LEAD) mine = lead;
        if (lane == RT_TRAIL) mine = trail;
        if (lane == RT_ANYNZ) mine = any ? 1u : 0u;
    }
    if (K2) {
        const u32 padded = (count + 7u) & ~7u;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const u32 v = __reduce_add_sync(0xffffffffu, k2[k]) + 3u * (k == 1 ? padded : count);
            if (lane == (u32)(RB_K2 + k)) mine = v;
        }
    }
    tacc[(size_t)tile * 32 + lane] = mine;
}

// KF: zero runs that cross tile boundaries.  One warp per block walks the tile records in order with the number of pending zeros
// (carry): a tile that holds a non-zero ends the pending run at its first non-zero — run = carry + lead, costed in THAT tile, also
// when lead is 0 (the run ended on the tile boundary); the block's last tile without a non-zero ends it at the end of the block.
// Every tile learns the position of the last non-zero before it (RT_LN0, what the pack pass starts from).
__global__ void k_rice2_fixup(u64* __restrict__ tacc, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                              const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks) {
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
            carry = an ? tr : (l ? 0u : carry + c);
        }
        if (t < nt) {
            a[RT_LN0] = (u64)(start - my_carry);
            const u32 run = (any || lead) ? my_carry + lead : 0u;
            if (run) {
                for (int k = 0; k < 7; ++k) a[RB_KF_Z + k] += (u64)(run >> k) + 1 + k;
                a[RB_KF_Z + 7] += 2 * bitlen32(run) - 1;
                a[RB_KF_NZ] += 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per-warp staging window of one step's bit range (cf. BitStage in rice.cu, which works per CTA)
// ---------------------------------------------------------------------------------------------
struct WStage {
    u32* sm; u32* out; u64 first_word; u32 nwords; bool use;
    __device__ __forceinline__ void begin(u32* smem, u32* out_, u64 bit0, u64 nbits) {
        sm = smem; out = out_;
        first_word = bit0 >> 5;
        const u64 nw = nbits ? ((bit0 + nbits - 1) >> 5) - first_word + 1 : 0;
        use = nw <= R2_STAGE; nwords = use ? (u32)nw : 0;
        for (u32 i = lane_id(); i < nwords; i += 32) sm[i] = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void word(u64 w, u32 be_bits) const {
        if (!be_bits) return;
        if (use) atomicOr(sm + (u32)(w - first_word), be_bits);
        else atomicOr(out + w, __byte_perm(be_bits, 0, 0x0123));
    }
    __device__ __forceinline__ void bits(u64 bitpos, u32 value, u32 n) const {      // n in 1..32, value < 2^n
        const u64 w = bitpos >> 5; const u32 o = (u32)bitpos & 31;
        if (o + n <= 32) { word(w, value << (32 - n - o)); return; }
        const u64 v = (u64)value << (64 - n - o);
        word(w, (u32)(v >> 32)); word(w + 1, (u32)v);
    }
    __device__ __forceinline__ void ones(u64 bitpos, u64 q) const {
        if (!use && q >= 96) {                              // long unary part straight to global memory: whole words are plain stores
            const u32 o = (u32)bitpos & 31;
            if (o) { const u32 h = 32 - o; bits(bitpos, (1u << h) - 1u, h); bitpos += h; q -= h; }
            u64 w = bitpos >> 5;
            while (q >= 32) { out[w++] = 0xffffffffu; q -= 32; }
            bitpos = w << 5;
        }
        while (q >= 32) { bits(bitpos, 0xffffffffu, 32); bitpos += 32; q -= 32; }
        if (q) bits(bitpos, (1u << q) - 1u, (u32)q);
    }
    __device__ __forceinline__ void rice(u64 bitpos, u64 x, u32 k) const {
        const u64 q = x >> k; ones(bitpos, q);
        if (k) bits(bitpos + q + 1, (u32)(x & ((1u << k) - 1u)), k);
    }
    __device__ __forceinline__ void gamma(u64 bitpos, u32 x) const { const u32 b = bitlen32(x); bits(bitpos + (b - 1), x, b); }
    __device__ __forceinline__ void flush() const {
        __syncwarp();
        for (u32 i = lane_id(); i < nwords; i += 32) {
            u32 v = sm[i];
            if (!v) continue;
            v = __byte_perm(v, 0, 0x0123);
            if (i == 0 || i == nwords - 1) atomicOr(out + first_word + i, v); else out[first_word + i] = v;
        }
        __syncwarp();
    }
};
// A lane's tokens are contiguous in the stream: they are shifted into a 64-bit register and leave as whole words; only the first
// word a lane touches and its trailing partial word can be shared with a neighbouring lane and use an atomic OR.
struct WAcc {
    u64 acc; u32 fill, w; bool first;
    __device__ __forceinline__ void init(const WStage& st, u64 bp) { w = (u32)((bp >> 5) - st.first_word); fill = (u32)bp & 31u; acc = 0; first = true; }
    __device__ __forceinline__ u64 bitpos(const WStage& st) const { return ((st.first_word + w) << 5) + fill; }
    __device__ __forceinline__ void push(const WStage& st, u32 code, u32 n) {        // n in 1..32, code < 2^n, st.use
        acc |= (u64)code << (64u - fill - n); fill += n;
        if (fill >= 32u) {
            const u32 word = (u32)(acc >> 32);
            if (first) { if (word) atomicOr(st.sm + w, word); first = false; } else st.sm[w] = word;
            ++w; acc <<= 32; fill -= 32u;
        }
    }
    __device__ __forceinline__ void finish(const WStage& st) {
        const u32 word = (u32)(acc >> 32);
        if (word) atomicOr(st.sm + w, word);
        acc = 0;
    }
};
// The same without the first-word bookkeeping: every completed word goes out with a shared-memory atomic OR (one instruction, like the
// plain store it replaces) and the cursor is a shared-memory pointer, so a push is a 64-bit shift, two ORs and a predicated flush.
struct WFast {
    u32 hi, lo, fill; u32* wp;
    __device__ __forceinline__ void init(const WStage& st, u32 rel) { wp = st.sm + (rel >> 5); fill = rel & 31u; hi = lo = 0; }   // rel: bits from the window's first word
    __device__ __forceinline__ void push(u32 code, u32 n) {                          // n in 1..32, code < 2^n
        const u64 v = (u64)code << (64u - fill - n);
        hi |= (u32)(v >> 32); lo |= (u32)v; fill += n;
        if (fill >= 32u) { atomicOr(wp, hi); ++wp; hi = lo; lo = 0; fill -= 32u; }
    }
    __device__ __forceinline__ u32 rel(const WStage& st) const { return (u32)(wp - st.sm) * 32u + fill; }
    __device__ __forceinline__ void finish() { if (hi) atomicOr(wp, hi); hi = 0; }
};
// tokens longer than 32 bits, or any token of a step that does not fit the window: rare, kept out of line
// (the window travels BY VALUE so that the callers' copy never has its address taken and stays in registers)
__device__ __noinline__ void r2_long_rice(WStage st, u64 bp, u32 tagbits, u32 tag, u32 x, u32 k) {
    if (tagbits && tag) st.bits(bp, 1, 1);
    st.rice(bp + tagbits, x, k);
}
__device__ __noinline__ void r2_long_gamma(WStage st, u64 bp, u32 tag, u32 x) {
    if (tag) st.bits(bp, 1, 1);
    st.gamma(bp + 1, x);
}

__device__ __forceinline__ u32 r2_warp_incl_add32(u32 v) {
    const u32 lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v += n; }
    return v;
}
// one KF token of more than 32 bits at absolute bit position bp (tag, then Rice(k) of val - tag or gamma of val)
__device__ __forceinline__ void r2_long_tok(const WStage& st, u64 bp, bool rice, u32 k, u32 tag, u32 val) {
    if (rice) r2_long_rice(st, bp, 1, tag, val - tag, k); else r2_long_gamma(st, bp, tag, val);
}
// KF token (KF.py:670-684): tag bit, then Rice(k) of x or gamma of g.  Zero run of length r: tag 0, x = g = r.  Non-zero v: tag 1,
// x = v - 1, g = v.  Returns the token's length in bits; code is valid when the length is <= 32.
__device__ __forceinline__ u32 r2_kf_token(bool rice, u32 k, u32 tag, u32 val /* r or v */, u32& code) {
    if (rice) {
        const u32 x = val - tag, q = x >> k, n = q + 2 + k;
        if (n <= 32) code = (tag << (n - 1)) | ((((u32)1 << q) - 1u) << (k + 1)) | (x & ((1u << k) - 1u));
        return n;
    }
    const u32 b = 32u - __clz(val), n = 2 * b;
    code = (tag << ((n - 1) & 31)) | val;
    return n;
}

// emits one KF token at the lane's running position
__device__ __forceinline__ void r2_kf_emit(const WStage& st, WAcc& ba, bool rice, u32 k, u32 tag, u32 val) {
    u32 code = 0;
    const u32 n = r2_kf_token(rice, k, tag, val, code);
    if (n <= 32 && st.use) ba.push(st, code, n);
    else {
        ba.finish(st);
        const u64 bp = ba.bitpos(st);
        if (rice) r2_long_rice(st, bp, 1, tag, val - tag, k); else r2_long_gamma(st, bp, tag, val);
        ba.init(st, bp + n);
    }
}

__global__ void __launch_bounds__(R2_THREADS, 4) k_rice2_kf_pack(const u8* __restrict__ mtf, const TileDesc* __restrict__ tiles,
                                                              const BlockInfo* __restrict__ binfo, const u64* __restrict__ tacc, const u64* __restrict__ bacc,
                                                              u32* __restrict__ out, int ntiles, const i64* __restrict__ cap_total, u64 cap,
                                                              const int* __restrict__ method, int want) {
    __shared__ u32 s_stage[R2_WARPS][R2_STAGE];
    if ((u64)*cap_total + 8 > cap) return;                  // exact total known before any bit is packed (words are written whole: 8 bytes of slack)
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * R2_WARPS + w;
    if (tile >= ntiles) return;
    const u64* trec = tacc + (size_t)tile * 32;
    const TileDesc td = tiles[tile];
    if (method && method[td.block] != want) return;         // kolm_encode_blocks: only the blocks this coder won are packed
    const BlockInfo bi = binfo[td.block];
    const u64* a = bacc + (size_t)td.block * RB_STRIDE;
    const u32 prm = (u32)a[RB_PARAM];
    const u32 k0 = prm & 0xff, k1 = (prm >> 8) & 0xff; const bool urz = (prm >> 16) & 1, urn = (prm >> 17) & 1;
    const u64 bitbase = a[RB_OFF] * 8;
    const u32 t0 = td.start - bi.pbase, count = td.count;
    const u8* src = mtf + bi.ioff + t0;
    const bool first_tile = (td.flags & 1u) != 0, last_tile = (td.flags & 2u) != 0;
    // block-local 1-based position of the last non-zero before the current step (warp uniform); RT_LN0 is the tile's entry state
    u32 carry = (u32)trec[RT_LN0];
    u64 bitpos = first_tile ? bitbase : bitbase + 10 + trec[RT_BITOFF];   // the first tile also owns the 10 header bits
    u32 wcur[4];
    r2_load16(src + lane * 16, min(16u, count > lane * 16 ? count - lane * 16 : 0u), count > lane * 16 + 16 ? count - lane * 16 - 16 : 0u, wcur);
    for (u32 base = 0; base < count; base += R2_STEP) {
        const u32 o = base + lane * 16;
        u32 wd[4] = {wcur[0], wcur[1], wcur[2], wcur[3]};
        {
            const u32 on = o + R2_STEP;
            if (base + R2_STEP < count) r2_load16(src + on, min(16u, count > on ? count - on : 0u), count > on + 16 ? count - on - 16 : 0u, wcur);
        }
        const u32 valid = count > o ? min(16u, count - o) : 0u;
        const u32 nzm0 = r2_nzmask(wd);
        const u32 mylast1 = nzm0 ? t0 + o + (31u - __clz(nzm0)) + 1u : 0u;
        const u32 incl = r2_warp_incl_max(mylast1);
        u32 ln0 = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) ln0 = 0;
        ln0 = max(ln0, carry);
        carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
        const bool hdr = first_tile && base == 0 && lane == 0;
        const bool tail = last_tile && valid && o + valid == count;     // I hold the block's last byte
        const u32 p0 = t0 + o;                                           // block-local position of my first byte
        // run tokens end at the non-zeros that follow a zero: bit i of rz = a run ends before my byte i
        const u32 rz = nzm0 & ~(nzm0 << 1) & ~((ln0 == p0) ? 1u : 0u);
        // pass 1: bits of my tokens.  Non-zero tokens as SIMD-in-register sums (no loop), run tokens one by one (few)
        u32 mybits = hdr ? 10u : 0u;
        {
            const u32 cnt = __popc(nzm0);
            if (urn) {                                       // Rice(k1) of x = v - 1: sum q + cnt * (2 + k1)
                const u32 mk = (0xFFu >> k1) * 0x01010101u;
                u32 q = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) { const u32 x = wd[j] - (__vcmpne4(wd[j], 0u) & 0x01010101u); q = __dp4a((x >> k1) & mk, 0x01010101u, q); }
                mybits += q + cnt * (2u + k1);
            } else {                                         // gamma of v: 2 * bitlen(v) per non-zero (tag included)
                u32 bl = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u32 x = wd[j];
                    x |= (x >> 1) & 0x7F7F7F7Fu; x |= (x >> 2) & 0x3F3F3F3Fu; x |= (x >> 4) & 0x0F0F0F0Fu;   // smear: bitlen = popcount
                    x = x - ((x >> 1) & 0x55555555u); x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u); x = (x + (x >> 4)) & 0x0F0F0F0Fu;
                    bl = __dp4a(x, 0x01010101u, bl);
                }
                mybits += 2u * bl;
            }
            u32 m = rz, code;
            while (m) {
                const u32 i = __ffs(m) - 1; m &= m - 1;
                const u32 below = nzm0 & ((1u << i) - 1u);
                const u32 ln = below ? p0 + (32u - __clz(below)) : ln0;      // 1-based position of the previous non-zero
                mybits += r2_kf_token(urz, k0, 0, p0 + i - ln, code);
            }
            if (tail) {
                const u32 ln = nzm0 ? p0 + (32u - __clz(nzm0)) : ln0;
                if (bi.len > ln) mybits += r2_kf_token(urz, k0, 0, bi.len - ln, code);
            }
        }
        const u32 incb = r2_warp_incl_add32(mybits);
        const u32 total = __shfl_sync(0xffffffffu, incb, 31);
        WStage st;
        st.begin(s_stage[w], out, bitpos, total);
        if (st.use) {
            // the step fits the window: lean accumulator, positions relative to the window
            WFast fa; fa.init(st, ((u32)bitpos & 31u) + (incb - mybits));
            if (hdr) fa.push((((urn ? 2u : 0u) | (urz ? 1u : 0u)) << 8) | (k0 << 4) | k1, 10);   // KF.py:664-668: flags, k0, k1
            u32 nzm = nzm0, ln = ln0;
            while (nzm) {
                const u32 i = __ffs(nzm) - 1; nzm &= nzm - 1;
                const u32 b = r2_byte(wd, i), pos = p0 + i;
                u32 code;
                if (pos > ln) {
                    const u32 n = r2_kf_token(urz, k0, 0, pos - ln, code);
                    if (n <= 32) fa.push(code, n);
                    else { fa.finish(); const u32 r = fa.rel(st); r2_long_tok(st, (st.first_word << 5) + r, urz, k0, 0, pos - ln); fa.init(st, r + n); }
                }
                const u32 n = r2_kf_token(urn, k1, 1, b, code);
                if (n <= 32) fa.push(code, n);
                else { fa.finish(); const u32 r = fa.rel(st); r2_long_tok(st, (st.first_word << 5) + r, urn, k1, 1, b); fa.init(st, r + n); }
                ln = pos + 1;
            }
            if (tail && bi.len > ln) {
                u32 code; const u32 n = r2_kf_token(urz, k0, 0, bi.len - ln, code);
                if (n <= 32) fa.push(code, n);
                else { fa.finish(); const u32 r = fa.rel(st); r2_long_tok(st, (st.first_word << 5) + r, urz, k0, 0, bi.len - ln); fa.init(st, r + n); }
            }
            fa.finish();
        } else {
            // a step whose bits exceed the window (long unary parts): every token straight to global memory
            WAcc ba; ba.init(st, bitpos + (incb - mybits));
            if (hdr) { st.bits(bitpos, (((urn ? 2u : 0u) | (urz ? 1u : 0u)) << 8) | (k0 << 4) | k1, 10); ba.init(st, bitpos + 10); }
            u32 nzm = nzm0, ln = ln0;
            while (nzm) {
                const u32 i = __ffs(nzm) - 1; nzm &= nzm - 1;
                const u32 b = r2_byte(wd, i), pos = p0 + i;
                if (pos > ln) r2_kf_emit(st, ba, urz, k0, 0, pos - ln);
                r2_kf_emit(st, ba, urn, k1, 1, b);
                ln = pos + 1;
            }
            if (tail && bi.len > ln) r2_kf_emit(st, ba, urz, k0, 0, bi.len - ln);
            ba.finish(st);
        }
        st.flush();
        bitpos += total;
    }
}

// V22 byte transforms of one word (four bytes at once): nibble swap, bit reverse, Gray (V22.py:1650-1680)
__device__ __forceinline__ u32 r2_xform4(u32 x, int flags) {
    if (flags & 4) x = ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
    if (flags & 8) x = __byte_perm(__brev(x), 0, 0x0123);    // reverse the bits of every byte, keep the byte order
    if (flags & 16) x = x ^ ((x >> 1) & 0x7F7F7F7Fu);
    return x;
}

__global__ void __launch_bounds__(R2_THREADS) k_rice2_k2_pack(const u8* __restrict__ mtf, const TileDesc* __restrict__ tiles,
                                                              const BlockInfo* __restrict__ binfo, const u64* __restrict__ tacc, const u64* __restrict__ bacc,
                                                              u32* __restrict__ out, int flags, int ntiles, const i64* __restrict__ cap_total, u64 cap,
                                                              const int* __restrict__ method, int want) {
    __shared__ u32 s_stage[R2_WARPS][R2_STAGE];
    if ((u64)*cap_total + 8 > cap) return;
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * R2_WARPS + w;
    if (tile >= ntiles) return;
    const TileDesc td = tiles[tile];
    if (method && method[td.block] != want) return;
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase, count = td.count;
    const u8* src = mtf + bi.ioff + t0;
    u64 bitpos = bacc[(size_t)td.block * RB_STRIDE + RB_OFF] * 8 + tacc[(size_t)tile * 32 + RT_BITOFF];
    u32 wcur[4];
    r2_load16(src + lane * 16, min(16u, count > lane * 16 ? count - lane * 16 : 0u), count > lane * 16 + 16 ? count - lane * 16 - 16 : 0u, wcur);
    for (u32 base = 0; base < count; base += R2_STEP) {
        const u32 o = base + lane * 16;
        u32 wd[4] = {wcur[0], wcur[1], wcur[2], wcur[3]};
        {
            const u32 on = o + R2_STEP;
            if (base + R2_STEP < count) r2_load16(src + on, min(16u, count > on ? count - on : 0u), count > on + 16 ? count - on - 16 : 0u, wcur);
        }
        u32 valid = count > o ? min(16u, count - o) : 0u;
        if (flags & 1) {                                     // 8x8 bit transpose of my two (zero padded) groups; a started group is coded whole
            const u64 g0 = bitplane8((u64)wd[0] | ((u64)wd[1] << 32)), g1 = bitplane8((u64)wd[2] | ((u64)wd[3] << 32));
            wd[0] = (u32)g0; wd[1] = (u32)(g0 >> 32); wd[2] = (u32)g1; wd[3] = (u32)(g1 >> 32);
            valid = (valid + 7u) & ~7u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) wd[j] = r2_xform4(wd[j], flags);
        u32 mybits = 0;                                      // sum over my coded symbols of (t >> 2) + 3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u32 nb = valid > 4u * j ? min(4u, valid - 4u * j) : 0u;
            const u32 keep = nb == 4 ? 0xffffffffu : ((1u << (8 * nb)) - 1u);
            mybits = __dp4a((wd[j] >> 2) & 0x3F3F3F3Fu & keep, 0x01010101u, mybits) + 3u * nb;
        }
        const u32 incb = r2_warp_incl_add32(mybits);
        const u32 total = __shfl_sync(0xffffffffu, incb, 31);
        WStage st;
        st.begin(s_stage[w], out, bitpos, total);
        // q ones, the terminating 0, 2 remainder bits per symbol; a symbol below 120 codes in at most 32 bits
        const bool allshort = ((__vcmpgeu4(wd[0], 0x78787878u) | __vcmpgeu4(wd[1], 0x78787878u) | __vcmpgeu4(wd[2], 0x78787878u) | __vcmpgeu4(wd[3], 0x78787878u)) == 0u);
        if (st.use && allshort) {
            WFast fa; fa.init(st, ((u32)bitpos & 31u) + (incb - mybits));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((u32)i < valid) {
                    const u32 t = (wd[i >> 2] >> (8 * (i & 3))) & 0xFFu, q = t >> 2;
                    fa.push(((((u32)1 << q) - 1u) << 3) | (t & 3u), q + 3);
                }
            }
            fa.finish();
        } else {
            WAcc ba; ba.init(st, bitpos + (incb - mybits));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if ((u32)i < valid) {
                    const u32 t = (wd[i >> 2] >> (8 * (i & 3))) & 0xFFu, q = t >> 2, n = q + 3;
                    if (n <= 32 && st.use) ba.push(st, ((((u32)1 << q) - 1u) << 3) | (t & 3u), n);
                    else {
                        ba.finish(st);
                        const u64 bp = ba.bitpos(st);
                        r2_long_rice(st, bp, 0, 0, t, 2);
                        ba.init(st, bp + n);
                    }
                }
            }
            ba.finish(st);
        }
        st.flush();
        bitpos += total;
    }
}

// ---------------------------------------------------------------------------------------------
// host drivers (the plan / offset / reduce kernels are rice.cu's)
// ---------------------------------------------------------------------------------------------
static int rice2_costs(kolm_ctx* c, const u8* mtf, bool kf, bool k2, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * RB_STRIDE * 8, s));
    if (!nt) return KOLM_OK;
    const int g = (nt + R2_WARPS - 1) / R2_WARPS;
    u64* tacc = (u64*)c->d_thist;
    if (kf && k2) KL(c, KC_RICE_COST, c->total_bytes, s, k_rice2_cost<true, true><<<g, R2_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, tacc, nt));
    else if (kf) KL(c, KC_RICE_COST, c->total_bytes, s, k_rice2_cost<true, false><<<g, R2_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, tacc, nt));
    else KL(c, KC_RICE_COST, c->total_bytes, s, k_rice2_cost<false, true><<<g, R2_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, tacc, nt));
    if (kf) KL(c, KC_RICE_COST, (i64)nt * 64, s, k_rice2_fixup<<<(nb + 3) / 4, 128, 0, s>>>(tacc, c->d_tiles, c->d_binfo, c->d_btile0, c->d_btilen, nb));
    KL(c, KC_RICE_COST, (i64)nt * 256, s, k_tile_reduce<<<nb, 256, 0, s>>>((const u64*)tacc, c->d_btile0, c->d_btilen, c->d_bacc, kf ? 0 : 20, (kf ? 20 : 0) + (k2 ? (kf ? 5 : 5) : 0)));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}
// plan + offsets + pack of one coder from the sums already in bacc / tacc.  mode 1 = KF, 2 = K2 (slot, flags)
static int rice2_pack(kolm_ctx* c, const u8* mtf, int mode, int slot, int flags, u8* out, size_t out_cap, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    KL(c, KC_RICE_PLAN, (i64)nb * 256, s, k_rice_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_binfo, c->d_poff, c->d_params, c->d_sizes, nb, mode, slot));
    KL(c, KC_ZERO, 0, s, k_zero_words<<<4 * c->sm_count, 256, 0, s>>>((u32*)out, c->d_poff + nb, out_cap / 4));
    if (nt) {
        const int g = (nt + R2_WARPS - 1) / R2_WARPS;
        KL(c, KC_RICE_PLAN, (i64)nt * 64, s, k_rice_tile_offsets<<<nb, 256, 0, s>>>((u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, mode, slot));
        if (mode == 1) KL(c, KC_RICE_PACK, c->total_bytes, s, k_rice2_kf_pack<<<g, R2_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, (const u64*)c->d_thist, c->d_bacc, (u32*)out, nt, c->d_poff + nb, (u64)out_cap, nullptr, 0));
        else KL(c, KC_RICE_PACK, c->total_bytes, s, k_rice2_k2_pack<<<g, R2_THREADS, 0, s>>>(mtf, c->d_tiles, c->d_binfo, (const u64*)c->d_thist, c->d_bacc, (u32*)out, flags, nt, c->d_poff + nb, (u64)out_cap, nullptr, 0));
    }
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}

int kolm_rice2_kf_enc_impl(kolm_ctx* c, const u8* mtf, u8* out, size_t out_cap, i64* out_off, int* params, cudaStream_t s) {
    if (((uintptr_t)out & 3) != 0) return KOLM_E_ARG;
    if (!c->nblocks) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    KOLM_TRY(rice2_costs(c, mtf, true, false, s));
    KOLM_TRY(rice2_pack(c, mtf, 1, 0, 0, out, out_cap, s));
    return rice_finish(c, out_off, params, nullptr, out_cap, s);
}

int kolm_rice2_k2_enc_impl(kolm_ctx* c, const u8* mtf, int flags, u8* out, size_t out_cap, i64* out_off, i64* sizes, cudaStream_t s) {
    const int slot = k2_slot(flags);
    if (slot < 0 || ((uintptr_t)out & 3) != 0) return KOLM_E_ARG;
    if (!c->nblocks) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    KOLM_TRY(rice2_costs(c, mtf, false, true, s));
    KOLM_TRY(rice2_pack(c, mtf, 2, slot, flags, out, out_cap, s));
    return rice_finish(c, out_off, nullptr, sizes, out_cap, s);
}

// both coders of one MTF batch from ONE cost read: KF model 2 (payload + parameters) and the five V22 variants' sizes + one packed
// variant.  The KF offsets travel home while the V22 pack runs.
int kolm_rice2_dual_enc_impl(kolm_ctx* c, const u8* mtf, int k2_flags, u8* kf_out, size_t kf_cap, i64* kf_off, int* kf_params,
                             u8* k2_out, size_t k2_cap, i64* k2_off, i64* k2_sizes, cudaStream_t s) {
    const int slot = k2_slot(k2_flags);
    const int nb = c->nblocks;
    if (slot < 0 || (((uintptr_t)kf_out | (uintptr_t)k2_out) & 3) != 0) return KOLM_E_ARG;
    if (!nb) { if (kf_off) kf_off[0] = 0; if (k2_off) k2_off[0] = 0; return KOLM_OK; }
    KOLM_TRY(rice2_costs(c, mtf, true, true, s));
    KOLM_TRY(rice2_pack(c, mtf, 1, 0, 0, kf_out, kf_cap, s));
    // the KF plan's results leave d_poff / d_params before the V22 plan reuses them: second set of pinned mirrors = h_bacc
    i64* h_kfoff = (i64*)c->h_bacc; int* h_kfprm = (int*)(c->h_bacc + (size_t)nb + 1);
    CUDA_TRY(cudaMemcpyAsync(h_kfoff, c->d_poff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h_kfprm, c->d_params, (size_t)nb * 16, cudaMemcpyDeviceToHost, s));
    KOLM_TRY(rice2_pack(c, mtf, 2, slot, k2_flags, k2_out, k2_cap, s));
    int rc = rice_finish(c, k2_off, nullptr, k2_sizes, k2_cap, s);          // synchronises the stream
    memcpy(kf_off, h_kfoff, (size_t)(nb + 1) * 8);
    if (kf_params) memcpy(kf_params, h_kfprm, (size_t)nb * 16);
    if ((size_t)kf_off[nb] + 8 > kf_cap) return KOLM_E_CAPACITY;
    c->algbytes[KC_RICE_PACK] += kf_off[nb];
    return rc;
}
