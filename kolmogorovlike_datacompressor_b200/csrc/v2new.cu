// v2new.cu — the V2 bit-plane pipeline, method 10 of the KOLR container.
//
//   decode_new_pipeline            kolm_final_researched_v2-2.py:1578-1648
//   _rice_decode_until_len         :1454-1486      unrle_binary :1184-1189      unpack_bits_from_bytes :1197-1201
//   bitplanes_to_bytes             :1156-1175      circuit_map_automaton_inverse :1054-1092 (models :664-901)
//
// Payload: header0 (mode<<5 | param_len), param (LE, param_len bytes), raw_mask, b1_mask, one k per ENCODED plane, then
// the eight planes MSB first, each byte aligned: RAW = ceil(L/8) packed bytes; ENCODED = Rice(k) run lengths of the
// BBWT of the plane (first bit in b1_mask), decoded until the runs sum to L.
//
// Decode = three steps over a batch of blocks:
//   k_v2_parse    one warp per block: lane 0 walks the bit stream (planes are chained: a plane starts where the previous one
//                 ended), all lanes fill the runs -> eight 0/1 byte planes per block;
//   inverse BBWT  of all 8*nblocks planes as one batch of the ordinary inverse transform (bbwt_inv.cu) — RAW planes are
//                 transformed too and that result is simply not used;
//   k_v2_combine  planes -> bytes (all lanes), then the model's inverse recurrence (serial in the byte index: raw[i] needs
//                 raw[i-1..i-3]) by lane 0, 16 bytes per load/store.
// The shipped reference never EMITS method 10; its decoder is live, and this is its GPU replacement.
#include "common.cuh"

struct V2Bits {                       // MSB-first reader over data[0, nbits/8)
    const u8* d; u64 nbits, pos;
    __device__ __forceinline__ bool ones_then_zero(u64& q) {        // unary part: q ones and the terminating zero
        q = 0;
        for (;;) {
            if (pos >= nbits) return false;
            const u32 sh = (u32)pos & 7u, rem = 8u - sh;
            const u32 w = (u32)d[pos >> 3] << (24 + sh);
            const u32 ones = (u32)__clz((int)~w);                   // <= rem: the padding below the byte reads as zeros
            q += ones; pos += ones;
            if (ones < rem) { ++pos; return true; }
        }
    }
    __device__ __forceinline__ bool bit(u32& b) { if (pos >= nbits) return false; b = (d[pos >> 3] >> (7 - ((u32)pos & 7u))) & 1u; ++pos; return true; }
};

#define V2_SAT (1ull << 62)

// params[4*b] = mode, param, raw_mask, 0
__global__ void __launch_bounds__(128) k_v2_parse(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo,
                                                  i64 off0, u8* __restrict__ planes, int* __restrict__ params, int* __restrict__ err, int nblocks) {
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const BlockInfo bi = binfo[b];
    const u64 L = bi.len;
    if (L == 0) { if (lane == 0) { err[b] = KOLM_OK; params[4 * b] = 0; params[4 * b + 1] = 0; params[4 * b + 2] = 0xff; } return; }
    const u8* d = pay + pay_off[b];
    const i64 n = pay_off[b + 1] - pay_off[b];
    u8* P0 = planes + 8 * (bi.ioff - off0);
    int e = KOLM_OK;
    // ---- header (every lane reads the same few bytes)
    u32 mode = 0, plen = 0, mp = 0, raw_mask = 0, b1_mask = 0; i64 pos = 0;
    if (n < 3) e = KOLM_E_CORRUPT;
    if (!e) { u32 h0 = d[pos++]; mode = (h0 >> 5) & 7u; plen = h0 & 7u; if (plen > 4) e = KOLM_E_CORRUPT; }
    if (!e && n < 1 + (i64)plen + 2) e = KOLM_E_CORRUPT;
    if (!e) {
        for (u32 i = 0; i < plen; ++i) mp |= (u32)d[pos++] << (8 * i);
        raw_mask = d[pos++]; b1_mask = d[pos++];
        if (pos + (8 - __popc(raw_mask)) > n) e = KOLM_E_CORRUPT;
    }
    if (e) { if (lane == 0) err[b] = e; return; }
    const u8* kl = d + pos; pos += 8 - __popc(raw_mask);
    const u8* data = d + pos; const i64 dlen = n - pos; i64 dpos = 0; u32 ki = 0;
    for (int j = 0; j < 8 && !e; ++j) {
        u8* P = P0 + (u64)j * L;
        if ((raw_mask >> j) & 1u) {
            const i64 need = (i64)((L + 7) / 8);
            if (dpos + need > dlen) { e = KOLM_E_CORRUPT; break; }
            const u8* src = data + dpos;
            for (u64 t = lane; t < L; t += 32) P[t] = (src[t >> 3] >> (7 - ((u32)t & 7u))) & 1u;
            dpos += need;
        } else {
            const u32 k = kl[ki++]; u32 bit = (b1_mask >> j) & 1u;
            V2Bits br; br.d = data; br.nbits = (u64)dlen * 8; br.pos = (u64)dpos * 8;
            u64 total = 0;
            while (total < L) {
                u64 val = 0; int te = KOLM_OK;
                if (lane == 0) {
                    u64 q, r = 0;
                    if (!br.ones_then_zero(q)) te = KOLM_E_CORRUPT;                    // BitReader: out of data (ValueError)
                    for (u32 i = 0; i < k && !te; ++i) { u32 x; if (!br.bit(x)) te = KOLM_E_CORRUPT; else r = r >= V2_SAT ? V2_SAT : ((r << 1) | x); }
                    // q * 2^k + r is an unbounded integer in the reference: saturate, anything above L is an overrun anyway
                    val = (k >= 62) ? (q ? V2_SAT : r) : ((q > (V2_SAT >> k)) ? V2_SAT : (q << k) + r);
                    if (!te && (val == 0 || val > L - total)) te = KOLM_E_CORRUPT;     // non-positive run / RLE overrun
                }
                te = __shfl_sync(0xffffffffu, te, 0);
                if (te) { e = te; break; }
                val = __shfl_sync(0xffffffffu, val, 0);
                for (u64 i = lane; i < val; i += 32) P[total + i] = (u8)bit;
                total += val; bit ^= 1u;
            }
            if (e) break;
            const u64 endbits = __shfl_sync(0xffffffffu, br.pos, 0);
            dpos = (i64)((endbits + 7) >> 3);                                           // align_next_byte
        }
    }
    if (lane == 0) { err[b] = e; params[4 * b] = (int)mode; params[4 * b + 1] = (int)mp; params[4 * b + 2] = (int)raw_mask; params[4 * b + 3] = 0; }
}

// predictor of raw[i] for i >= 3 (and i >= k for the delta model) from the last raw bytes: hist = raw[i-1] | raw[i-2]<<8 | raw[i-3]<<16 | raw[i-4]<<24
__device__ __forceinline__ u32 v2_dilate(u32 x) { return (x | (x << 1) | (x >> 1)) & 0xFFu; }
__device__ __forceinline__ u32 v2_erode(u32 x) { return (~v2_dilate(~x & 0xFFu)) & 0xFFu; }
__device__ __forceinline__ u32 v2_gray(u32 v) { return (v ^ (v >> 1)) & 0xFFu; }
__device__ __forceinline__ u32 v2_pred(u32 mode, u32 param, u32 hist) {
    const u32 p1 = hist & 0xFFu, p2 = (hist >> 8) & 0xFFu, p3 = (hist >> 16) & 0xFFu;
    switch (mode) {
        case 1: return (hist >> (8 * (param - 1))) & 0xFFu;                             // Delta-k, k = 1..4          V22.py:677-688
        case 2: switch (param & 3u) { case 0: return v2_gray(p1); case 1: return v2_gray(p2); case 2: return v2_gray(p1 ^ p2); default: return v2_gray(p1 | p2); }
        case 3: return (p1 & 0xF0u) | (p2 & 0x0Fu);                                     // Nibble-MUX: the mux always yields a's high, b's low nibble (V22.py:803-821)
        case 4: return (p1 & p2) | (p1 & p3) | (p2 & p3);                               // majority of three          V22.py:848-863
        default: {                                                                      // Morpho-Predict             V22.py:886-901
            const u32 dil = v2_dilate(p1), ero = v2_erode(p1), edge = dil ^ ero;
            const u32 mor = (param & 1u) == 0 ? v2_erode(dil) : v2_dilate(ero);
            return (mor & edge) | (p1 & ~edge & 0xFFu);
        }
    }
}
// the first bytes of a block, where the models fall back to shorter histories (V22.py:664-901, `backward` methods)
__device__ __forceinline__ bool v2_pred_head(u32 mode, u32 param, const u8* raw, u64 i, u32& p) {
    switch (mode) {
        case 1: if (param == 0 || i < param) return false; p = raw[i - param]; return true;
        case 2: case 3: if (i == 0) return false; if (i == 1) { p = raw[0]; return true; } break;
        case 4: if (i == 0) return false; if (i < 3) { p = raw[i - 1]; return true; } break;
        case 5: if (i == 0) return false; break;
        default: return false;
    }
    u32 hist = raw[i - 1];
    if (i >= 2) hist |= (u32)raw[i - 2] << 8;
    if (i >= 3) hist |= (u32)raw[i - 3] << 16;
    p = v2_pred(mode, param, hist);
    return true;
}

// binfo describes the 8*nblocks PLANES (block b = planes 8b..8b+7, all of length L); out block b at off0 + binfo[8b].ioff / 8
__global__ void __launch_bounds__(128) k_v2_combine(const u8* __restrict__ planes_raw, const u8* __restrict__ planes_inv, const BlockInfo* __restrict__ binfo,
                                                    const int* __restrict__ params, u8* __restrict__ out, i64 off0, int nblocks) {
    const u32 lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const BlockInfo bi = binfo[8 * b];
    const u64 L = bi.len;
    if (!L) return;
    const u32 mode = (u32)params[4 * b], param = (u32)params[4 * b + 1], raw_mask = (u32)params[4 * b + 2];
    u8* dst = out + off0 + bi.ioff / 8;
    for (u64 t = lane; t < L; t += 32) {
        u32 v = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const u8* src = ((raw_mask >> j) & 1u) ? planes_raw : planes_inv;
            v |= ((u32)src[bi.ioff + (u64)j * L + t] & 1u) << (7 - j);
        }
        dst[t] = (u8)v;
    }
    __syncwarp();
    if (lane != 0 || mode < 1 || mode > 5 || (mode == 1 && param == 0)) return;
    if (mode == 1 && param > 4) {                             // never produced by the reference encoder (k = 1..4): plain loop
        for (u64 i = param; i < L; ++i) dst[i] ^= dst[i - param];
        return;
    }
    // head: until the history is full and dst + i is 16-byte aligned
    u64 i = 0;
    for (; i < L && (i < 4 || ((uintptr_t)(dst + i) & 15)); ++i) { u32 p; if (v2_pred_head(mode, param, dst, i, p)) dst[i] ^= (u8)p; }
    if (i >= L) return;
    u32 hist = (u32)dst[i - 1] | ((u32)dst[i - 2] << 8) | ((u32)dst[i - 3] << 16) | ((u32)dst[i - 4] << 24);
    for (; i + 16 <= L; i += 16) {
        uint4 q = *reinterpret_cast<const uint4*>(dst + i);
        u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const u32 r = ((w[k >> 2] >> (8 * (k & 3))) & 0xFFu) ^ v2_pred(mode, param, hist);
            hist = (hist << 8) | r;
            w[k >> 2] = (w[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (r << (8 * (k & 3)));
        }
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (; i < L; ++i) { const u32 r = (u32)dst[i] ^ v2_pred(mode, param, hist); hist = (hist << 8) | r; dst[i] = (u8)r; }
}

int kolm_set_batch(kolm_ctx* c, const i64* off, int nblocks, cudaStream_t s);
int kolm_bbwt_inv_impl(kolm_ctx* c, const u8* in, u8* out, cudaStream_t s);

// payload / out: device.  pay_off, off: host.  Needs a context of >= 8x the batch bytes and 8x the blocks.
int kolm_v2new_dec_impl(kolm_ctx* c, const u8* payload, const i64* pay_off, const i64* off, int nblocks, u8* out, cudaStream_t s) {
    if (nblocks <= 0) return KOLM_OK;
    const i64 off0 = off[0], total = off[nblocks] - off0;
    if (8 * (i64)nblocks > c->max_blocks || (size_t)(8 * total) + (size_t)KOLM_PAD * 8 * nblocks + 2 * KOLM_PAD > c->max_elems) return KOLM_E_CAPACITY;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    memcpy(c->h_poff, pay_off, (size_t)(nblocks + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nblocks + 1) * 8, cudaMemcpyHostToDevice, s));
    KL(c, KC_MISC, total * 9, s, k_v2_parse<<<(nblocks + 3) / 4, 128, 0, s>>>(payload, c->d_poff, c->d_binfo, off0, c->d_tmp8a, c->d_params, c->d_err, nblocks));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_err, c->d_err, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < nblocks; ++b) if (c->h_err[b]) return c->h_err[b];
    // the 8*nblocks planes as one batch of the inverse transform
    std::vector<i64> poffs((size_t)8 * nblocks + 1);
    for (int b = 0; b < nblocks; ++b) {
        const i64 L = off[b + 1] - off[b], base = 8 * (off[b] - off0);
        for (int j = 0; j < 8; ++j) poffs[(size_t)8 * b + j] = base + j * L;
    }
    poffs[(size_t)8 * nblocks] = 8 * total;
    KOLM_TRY(kolm_set_batch(c, poffs.data(), 8 * nblocks, s));
    KOLM_TRY(kolm_bbwt_inv_impl(c, c->d_tmp8a, c->d_tmp8b, s));
    KL(c, KC_MISC, total * 10, s, k_v2_combine<<<(nblocks + 3) / 4, 128, 0, s>>>(c->d_tmp8a, c->d_tmp8b, c->d_binfo, c->d_params, out, off0, nblocks));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s));                      // poffs (host) was read by set_batch's staging copy; keep the call self-contained
    return KOLM_OK;
}
