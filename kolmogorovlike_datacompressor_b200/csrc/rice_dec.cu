// rice_dec.cu — parallel decoding of the two Rice/gamma bitstreams (decode side).
//
//   KF  decode_model_bbwt_mtf bit parser   kolm_final.py:771-794  (BitReader :455-497)
//   V22 rice_decode (k = 2)                kolm_final_researched_v2-2.py:1423-1452, transforms undone as in :2080-2086
//
// A prefix-code stream has no random access, but it re-synchronises: a decoder started at a wrong bit falls into step with
// the true token sequence after a few tokens.  Each payload is cut into chunks of RD_CHUNK bits, one thread per chunk:
//   k_rdec_iter : decode from the current entry guess up to the end of the chunk -> exit (= entry of the next chunk) and the
//                 number of symbols produced by tokens starting in the chunk.  Iterated (Jacobi) until no entry changes; the
//                 first chunk of a block starts at the true position, so a stable state is the true parse.
//   k_rdec_scan : per block exclusive scan of the symbol counts.
//   k_rdec_write: decode once more from the true entries and store the symbols (the output is pre-zeroed, zero runs only skip).
// Streams that do not settle within RD_MAX_ITERS sweeps fall back to the one-thread-per-block parser in rice.cu.
#include "common.cuh"

#define RD_CHUNK 1024u            // bits per chunk
#define RD_MAX_ITERS 64

struct RdecArgs {
    const u8* pay; const i64* pay_off; const BlockInfo* binfo;
    const u32* cblock; const u32* cfirst;      // chunk -> block, block -> first chunk
    u32* entry; u32* exitp; u32* count;        // per chunk (bit positions relative to the block's payload)
    u32* changed; int* err;
    u8* out; u32 nchunks; int flags;           // flags: V22 transform flags (K2 only)
};

struct BitWin {                                 // MSB-first bit window over one payload
    const u8* p; u64 nbits;
    // 32 bits starting at pos (zero beyond the end).  Two ALIGNED 32-bit loads and a funnel shift: an aligned word that holds at
    // least one byte of the payload lies in the same page as that byte, so the loads cannot fault even when the payload itself
    // is not 4-byte aligned; bits that belong to the neighbouring payload are masked off.
    __device__ __forceinline__ u32 peek32(u64 pos) const {
        if (pos >= nbits) return 0;
        const u64 byte = pos >> 3; const u32 sh = (u32)pos & 7u;
        const uintptr_t a = (uintptr_t)(p + byte); const u32 off = (u32)a & 3u;
        const u32* w = reinterpret_cast<const u32*>(a - off);
        const u32 w0 = __byte_perm(w[0], 0, 0x0123);
        const u32 w1 = ((byte + 4 - off) * 8 < nbits) ? __byte_perm(w[1], 0, 0x0123) : 0u;
        u32 v = __funnelshift_l(w1, w0, off * 8 + sh);
        const u64 valid = nbits - pos;
        if (valid < 32) v &= 0xffffffffu << (32 - (u32)valid);
        return v;
    }
    // length of the run of `want` bits starting at pos (stops at the end of the payload)
    __device__ __forceinline__ u64 run(u64 pos, u32 want) const {
        u64 q = 0;
        for (;;) {
            if (pos + q >= nbits) return q;
            u32 w = peek32(pos + q);
            if (want) w = ~w;
            if (w) { u32 lz = __clz(w); u64 lim = nbits - (pos + q); return q + (lz < lim ? lz : lim); }
            q += 32;
        }
    }
};

// one KF token starting at pos: returns false at end of stream.  nsym = symbols it produces, val = non-zero symbol (tag 1)
__device__ __forceinline__ bool kf_token(const BitWin& bw, u64 pos, u32 k0, u32 k1, bool urz, bool urn, u64& npos, u64& nsym, u32& val, bool& bad) {
    if (pos >= bw.nbits) return false;
    u32 tag = bw.peek32(pos) >> 31;
    u64 p = pos + 1;
    bool rice = tag ? urn : urz; u32 k = tag ? k1 : k0;
    u64 v;
    if (rice) {
        u64 q = bw.run(p, 1);
        p += q + 1;                               // the terminating 0
        if (p + k > bw.nbits) return false;
        u32 r = k ? (bw.peek32(p) >> (32 - k)) : 0;
        p += k; v = (q << k) | r;
    } else {
        u64 z = bw.run(p, 0);
        if (p + z >= bw.nbits) return false;      // no terminating 1
        if (z > 31) { bad = true; z = 31; }
        p += z + 1;
        if (p + z > bw.nbits) return false;
        u32 r = z ? (bw.peek32(p) >> (32 - z)) : 0;
        p += z; v = (1ull << z) | r;
    }
    npos = p;
    if (tag) { val = (u32)(v - (rice ? 0 : 1)) + 1; nsym = 1; if (v - (rice ? 0 : 1) + 1 > 255) bad = true; }
    else { val = 0; nsym = v; }
    return true;
}

__device__ __forceinline__ bool k2_token(const BitWin& bw, u64 pos, u64& npos, u32& val, bool& bad) {
    if (pos >= bw.nbits) return false;
    u64 q = bw.run(pos, 1);
    u64 p = pos + q + 1;
    if (p + 2 > bw.nbits || pos + q >= bw.nbits) return false;
    u32 r = bw.peek32(p) >> 30;
    npos = p + 2;
    u64 v = q * 4 + r;
    if (v > 255) bad = true;
    val = (u32)v & 0xFF;
    return true;
}

struct KfHdr { u32 k0, k1; bool urz, urn; };
__device__ __forceinline__ KfHdr kf_header(const BitWin& bw) {
    u32 h = bw.peek32(0) >> 22;                  // 10 bits
    KfHdr r; r.urz = (h >> 8) & 1u; r.urn = (h >> 9) & 1u; r.k0 = (h >> 4) & 15u; r.k1 = h & 15u;
    return r;
}

// Register bit window for the common case.  peek32 costs two dependent loads per token (~4 bits of payload); the cursor keeps
// 33..64 upcoming bits in a 64-bit register and refills with one aligned word per 32 bits consumed, so the token-to-token
// dependency is arithmetic only.  It is used while at least 64 bits of payload lie ahead (no end-of-stream cases inside) and for
// tokens that fit 33 bits (unary part up to 16, gamma width up to 15); everything else goes through kf_token / k2_token above,
// whose results define the format, and the cursor is re-seated behind that token.
struct BitCur {
    const u32* wbase; u64 buf; u32 avail, widx;
    __device__ __forceinline__ void seat(const BitWin& bw, u64 pos) {
        const uintptr_t a = (uintptr_t)(bw.p + (pos >> 3));
        wbase = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
        const u32 skip = ((u32)a & 3u) * 8u + ((u32)pos & 7u);
        buf = (u64)__byte_perm(wbase[0], 0, 0x0123) << (32 + skip);
        avail = 32 - skip; widx = 1;
    }
    __device__ __forceinline__ void refill() {
        if (avail <= 32) { buf |= (u64)__byte_perm(wbase[widx++], 0, 0x0123) << (32 - avail); avail += 32; }
    }
    __device__ __forceinline__ u32 top() const { return (u32)(buf >> 32); }
    __device__ __forceinline__ void skip(u32 n) { buf <<= n; avail -= n; }
};

// fast KF token from the cursor: returns the number of bits consumed, 0 if the token needs the general path
__device__ __forceinline__ u32 kf_token_fast(BitCur& cu, u32 k0, u32 k1, bool urz, bool urn, u64& nsym, u32& val) {
    cu.refill();
    const u32 t = cu.top();
    const u32 tag = t >> 31;
    const bool rice = tag ? urn : urz; const u32 k = tag ? k1 : k0;
    const u32 body = t << 1;                               // the 31 bits after the tag
    u32 n; u64 v;
    if (rice) {
        const u32 q = __clz(~body);                        // leading ones
        if (q > 16) return 0;
        n = 1 + q + 1 + k;                                 // <= 33
        const u64 rest = cu.buf << (q + 2);                // bits after the terminating 0
        const u32 r = k ? (u32)(rest >> (64 - k)) : 0u;
        v = ((u64)q << k) | r;
    } else {
        const u32 z = __clz(body);                         // leading zeros before the 1 (body == 0 -> 32)
        if (z > 15) return 0;
        n = 1 + z + 1 + z;                                 // <= 32
        const u64 rest = cu.buf << (z + 2);
        const u32 r = z ? (u32)(rest >> (64 - z)) : 0u;
        v = (1ull << z) | r;
    }
    if (tag) { const u64 x = v - (rice ? 0 : 1) + 1; if (x > 255) return 0; val = (u32)x; nsym = 1; }
    else { val = 0; nsym = v; }
    cu.skip(n);
    return n;
}

__device__ __forceinline__ u32 k2_token_fast(BitCur& cu, u32& val) {
    cu.refill();
    const u32 q = __clz(~cu.top());
    if (q > 24) return 0;
    const u32 r = (u32)((cu.buf << (q + 1)) >> 62);
    const u32 v = q * 4 + r;
    if (v > 255) return 0;
    val = v;
    cu.skip(q + 3);
    return q + 3;
}

// one token at pos through the cursor when it applies, else through the general parser (cursor re-seated afterwards)
template <bool KF>
__device__ __forceinline__ bool rd_token(const BitWin& bw, BitCur& cu, bool& seated, u64 pos, const KfHdr& h, u64& npos, u64& nsym, u32& val, bool& bad) {
    if (pos + 64 <= bw.nbits) {
        if (!seated) { cu.seat(bw, pos); seated = true; }
        nsym = 1;
        const u32 n = KF ? kf_token_fast(cu, h.k0, h.k1, h.urz, h.urn, nsym, val) : k2_token_fast(cu, val);
        if (n) { npos = pos + n; return true; }
    }
    seated = false;
    nsym = 1;
    return KF ? kf_token(bw, pos, h.k0, h.k1, h.urz, h.urn, npos, nsym, val, bad) : k2_token(bw, pos, npos, val, bad);
}

template <bool KF>
__global__ void __launch_bounds__(128) k_rdec_iter(RdecArgs a, int first_iter) {
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nchunks) return;
    const u32 b = a.cblock[c];
    const u32 c0 = a.cfirst[b];
    BitWin bw; bw.p = a.pay + a.pay_off[b]; bw.nbits = (u64)(a.pay_off[b + 1] - a.pay_off[b]) * 8;
    const u32 nc = (u32)((bw.nbits + RD_CHUNK - 1) / RD_CHUNK);
    u32 lc = c - c0;
    u64 e;
    if (lc == 0) e = KF ? 10 : 0;
    else e = first_iter ? (u64)lc * RD_CHUNK : a.exitp[c - 1];
    if (!first_iter && a.entry[c] == (u32)e) return;
    KfHdr h; if (KF) h = kf_header(bw);
    // Decode my chunk from the new entry; if that moves my exit, keep walking into the following chunks until the parse
    // meets an entry that is already recorded there (re-synchronised) — periodic streams can otherwise hold a shifted,
    // self-consistent parse for many chunks and would need one sweep per chunk.
    for (u32 walked = 0;; ++walked) {
        const u64 cend = (u64)(lc + 1) * RD_CHUNK;
        u64 pos = e, cnt = 0; bool bad = false;
        BitCur cu; bool seated = false;
        while (pos < cend) {
            u64 np, ns = 1; u32 val;
            bool ok = rd_token<KF>(bw, cu, seated, pos, h, np, ns, val, bad);
            if (!ok) { pos = bw.nbits > cend ? bw.nbits : cend; break; }      // truncated token: nothing more starts in this block
            cnt += ns; pos = np;
        }
        if (cnt > 0xffffffffull) cnt = 0xffffffffull;
        a.entry[c0 + lc] = (u32)e; a.exitp[c0 + lc] = (u32)pos; a.count[c0 + lc] = (u32)cnt;
        if (first_iter) break;
        ++lc;
        if (lc >= nc || walked >= 4096u) break;
        if (a.entry[c0 + lc] == (u32)pos) break;                               // the next chunk already starts there
        e = pos;
    }
    *a.changed = 1;
}

// per block: exclusive scan of count[] over its chunks (in place, saturating at 2^32-1)
__global__ void __launch_bounds__(256) k_rdec_scan(u32* __restrict__ count, const u32* __restrict__ cfirst, const i64* __restrict__ pay_off, int nblocks) {
    __shared__ u64 s_w[8];
    __shared__ u64 s_carry;
    const u32 b = blockIdx.x, tid = threadIdx.x;
    const u32 c0 = cfirst[b];
    const u64 nbits = (u64)(pay_off[b + 1] - pay_off[b]) * 8;
    const u32 nc = (u32)((nbits + RD_CHUNK - 1) / RD_CHUNK);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (u32 base = 0; base < nc; base += 256) {
        u32 i = base + tid;
        u64 v = i < nc ? count[c0 + i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) { u64 n = __shfl_up_sync(0xffffffffu, x, o); if ((tid & 31) >= (u32)o) x += n; }
        if ((tid & 31) == 31) s_w[tid >> 5] = x;
        __syncthreads();
        u64 pre = 0;
        for (u32 k = 0; k < (tid >> 5); ++k) pre += s_w[k];
        u64 carry = s_carry;
        u64 ex = carry + pre + x - v;
        if (i < nc) count[c0 + i] = ex > 0xffffffffull ? 0xffffffffu : (u32)ex;
        __syncthreads();
        if (tid == 255) s_carry = carry + pre + x;
        __syncthreads();
    }
}

template <bool KF>
__global__ void __launch_bounds__(128) k_rdec_write(RdecArgs a) {
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nchunks) return;
    const u32 b = a.cblock[c];
    const u32 lc = c - a.cfirst[b];
    const BlockInfo bi = a.binfo[b];
    BitWin bw; bw.p = a.pay + a.pay_off[b]; bw.nbits = (u64)(a.pay_off[b + 1] - a.pay_off[b]) * 8;
    const u64 cend = (u64)(lc + 1) * RD_CHUNK;
    KfHdr h; if (KF) h = kf_header(bw);
    u64 pos = a.entry[c], o = a.count[c];
    u8* dst = a.out + bi.ioff;
    bool bad = false;
    const bool last = cend >= bw.nbits;
    BitCur cu; bool seated = false;
    while (pos < cend && o < bi.len) {
        u64 np, ns = 1; u32 val;
        bool ok = rd_token<KF>(bw, cu, seated, pos, h, np, ns, val, bad);
        if (!ok) { atomicMax((unsigned int*)(a.err + b), 1u); break; }    // stream ends inside a token before orig_len symbols
        if (KF) { if (val) { if (bad) break; dst[o] = (u8)val; } }
        else {
            if (bad) break;
            u32 t = val;
            if (a.flags & 16) { t ^= t >> 1; t ^= t >> 2; t ^= t >> 4; t &= 0xFF; }
            if (a.flags & 8) t = __brev(t) >> 24;
            if (a.flags & 4) t = ((t & 0x0F) << 4) | ((t & 0xF0) >> 4);
            dst[o] = (u8)t;
        }
        o += ns; pos = np;
    }
    if (bad && o < bi.len) atomicMax((unsigned int*)(a.err + b), 2u);
    if (last && o < bi.len && pos >= bw.nbits) atomicMax((unsigned int*)(a.err + b), 1u);   // ran out of bits: truncated
}

// inverse 8x8 bit-plane transpose per group of 8 symbols (V22 flag 1); blocks with len % 8 != 0 raise IndexError in the reference
__global__ void k_rdec_bitplane(u8* __restrict__ out, const BlockInfo* __restrict__ binfo, const TileDesc* __restrict__ tiles, int* __restrict__ err) {
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    if (bi.len % 8) { if (threadIdx.x == 0) atomicMax((unsigned int*)(err + td.block), 3u); return; }
    u8* dst = out + bi.ioff + (td.start - bi.pbase);
    for (u32 g0 = threadIdx.x * 8; g0 < td.count; g0 += KOLM_THREADS * 8) {
        u64 g = 0;
        for (int i = 0; i < 8; ++i) g |= (u64)dst[g0 + i] << (8 * i);
        u64 t = bitplane8(g);
        for (int i = 0; i < 8; ++i) dst[g0 + i] = (u8)(t >> (8 * i));
    }
}

__global__ void k_rdec_chunkmap(const i64* __restrict__ pay_off, const u32* __restrict__ cfirst, u32* __restrict__ cblock, int nblocks) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    u64 nbits = (u64)(pay_off[b + 1] - pay_off[b]) * 8;
    u32 n = (u32)((nbits + RD_CHUNK - 1) / RD_CHUNK), f = cfirst[b];
    for (u32 k = 0; k < n; ++k) cblock[f + k] = b;
}

template <bool KF>
static int rice_dec_parallel(kolm_ctx* c, const u8* pay, const i64* pay_off, int flags, u8* out, cudaStream_t s, bool* fell_back) {
    const int nb = c->nblocks;
    *fell_back = false;
    // chunk tables
    u32* cfirst = c->h_u32 + 2 * (size_t)nb;
    u64 total_chunks = 0;
    for (int b = 0; b < nb; ++b) {
        u64 nbits = (u64)(pay_off[b + 1] - pay_off[b]) * 8;
        if (nbits >= (1ull << 32)) { *fell_back = true; return KOLM_OK; }          // 32-bit bit positions
        cfirst[b] = (u32)total_chunks;
        total_chunks += (nbits + RD_CHUNK - 1) / RD_CHUNK;
    }
    if (total_chunks * 4 > (u64)c->max_elems || total_chunks >= (1ull << 31)) { *fell_back = true; return KOLM_OK; }
    const u32 nchunks = (u32)total_chunks;
    memcpy(c->h_poff, pay_off, (size_t)(nb + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->d_atile0, cfirst, (size_t)nb * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(c->d_err, 0, (size_t)nb * 4, s));
    if (c->total_bytes) CUDA_TRY(cudaMemsetAsync(out + c->h_binfo[0].ioff, 0, (size_t)c->total_bytes, s));
    if (!nchunks) return KOLM_OK;
    RdecArgs a;
    a.pay = pay; a.pay_off = c->d_poff; a.binfo = c->d_binfo; a.cfirst = c->d_atile0;
    a.cblock = c->d_k0; a.entry = c->d_v0; a.exitp = c->d_k1; a.count = c->d_v1; a.changed = c->d_stats + 12; a.err = c->d_err;
    a.out = out; a.nchunks = nchunks; a.flags = flags;
    KL(c, KC_MISC, (i64)nchunks * 4, s, k_rdec_chunkmap<<<(nb + 127) / 128, 128, 0, s>>>(c->d_poff, c->d_atile0, c->d_k0, nb));
    const int grid = (int)((nchunks + 127) / 128);
    const i64 pbytes = pay_off[nb] - pay_off[0];
    int it = 0;
    for (; it < RD_MAX_ITERS; ++it) {
        CUDA_TRY(cudaMemsetAsync(c->d_stats + 12, 0, 4, s));
        KL(c, KC_RICE_PACK, pbytes, s, k_rdec_iter<KF><<<grid, 128, 0, s>>>(a, it == 0 ? 1 : 0));
        CUDA_TRY(cudaMemcpyAsync(c->h_stats + 12, c->d_stats + 12, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (!c->h_stats[12]) break;
    }
    CUDA_TRY(cudaGetLastError());
    if (it == RD_MAX_ITERS) { *fell_back = true; return KOLM_OK; }
    KL(c, KC_RICE_PLAN, (i64)nchunks * 8, s, k_rdec_scan<<<nb, 256, 0, s>>>(c->d_v1, c->d_atile0, c->d_poff, nb));
    KL(c, KC_RICE_PACK, pbytes + c->total_bytes, s, k_rdec_write<KF><<<grid, 128, 0, s>>>(a));
    if (!KF && (flags & 1) && c->ntiles) KL(c, KC_MISC, c->total_bytes * 2, s, k_rdec_bitplane<<<c->ntiles, KOLM_THREADS, 0, s>>>(out, c->d_binfo, c->d_tiles, c->d_err));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_err, c->d_err, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < nb; ++b) {
        if (c->h_err[b] == 3) return KOLM_E_INDEX;
        if (c->h_err[b] == 2) return KF ? KOLM_E_INDEX : KOLM_E_CORRUPT;
        if (c->h_err[b] == 1) return KOLM_E_TRUNCATED;
    }
    return KOLM_OK;
}


// ---- entry points: parallel parser first, one-thread-per-block parser (rice.cu) as the fallback
int kolm_rice_kf_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, u8* mtf_out, cudaStream_t s) {
    const int nb = c->nblocks;
    if (!nb) return KOLM_OK;
    static int par = -1;
    if (par < 0) { const char* e = getenv("KOLM_RICE_DEC_PARALLEL"); par = e ? atoi(e) : 1; }
    if (par) {
        bool fb = false;
        int r = rice_dec_parallel<true>(c, pay, pay_off, 0, mtf_out, s, &fb);
        if (!fb) return r;
    }
    memcpy(c->h_poff, pay_off, (size_t)(nb + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    if (c->total_bytes) CUDA_TRY(cudaMemsetAsync(mtf_out + c->h_binfo[0].ioff, 0, (size_t)c->total_bytes, s));
    KL(c, KC_RICE_PACK, c->total_bytes + pay_off[nb] - pay_off[0], s, k_rice_kf_dec<<<(nb + 63) / 64, 64, 0, s>>>(pay, c->d_poff, c->d_binfo, mtf_out, c->d_err, nb));
    CUDA_TRY(cudaGetLastError());
    return rice_dec_finish(c, s);
}

int kolm_rice_k2_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, int flags, u8* mtf_out, cudaStream_t s) {
    const int nb = c->nblocks;
    if (k2_slot(flags) < 0) return KOLM_E_ARG;
    if (!nb) return KOLM_OK;
    static int par = -1;
    if (par < 0) { const char* e = getenv("KOLM_RICE_DEC_PARALLEL"); par = e ? atoi(e) : 1; }
    if (par) {
        bool fb = false;
        int r = rice_dec_parallel<false>(c, pay, pay_off, flags, mtf_out, s, &fb);
        if (!fb) return r;
    }
    memcpy(c->h_poff, pay_off, (size_t)(nb + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    KL(c, KC_RICE_PACK, c->total_bytes + pay_off[nb] - pay_off[0], s, k_rice_k2_dec<<<(nb + 63) / 64, 64, 0, s>>>(pay, c->d_poff, c->d_binfo, mtf_out, c->d_err, nb, flags));
    CUDA_TRY(cudaGetLastError());
    return rice_dec_finish(c, s);
}
