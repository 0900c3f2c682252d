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

// =================================================================================================
// Encoder: encode_new_pipeline (V22.py:1498-1576) with circuit_map_automaton_forward(parallel=False) (V22.py:1013-1052).
// The shipped reference cannot reach this code (its default parallel=True path uses names the file never imports, SURVEY
// fact 4), so the KOLR drop-in only offers it as an opt-in candidate; the bytes equal what the reference's function
// returns once that path is taken (pinned by tests/golden/v2new.json).
//   k_v2_hist       residual histograms of the 13 candidate models in one read  -> host: H0 in fp64 (glibc log2, the
//                   reference's own arithmetic and summation order) and the _pick_better fold (V22.py:936-946)
//   k_v2_planes     chosen model's residuals -> eight 0/1 byte planes + their packed form
//   BBWT            of all 8*nblocks planes as one batch of the ordinary forward transform
//   k_v2_runs_cost  run lengths of every transformed plane, Rice cost for k = 0..15 (run starts by a look-back max-scan)
//   host            k per plane (first minimum of the byte sizes), RAW when ceil(L/8) <= rice + 1, payload layout
//   k_v2_pack       RAW planes copied, ENCODED planes Rice packed (bit offsets by a look-back add-scan)
// =================================================================================================
__device__ __forceinline__ u32 v2_forward_byte(u32 mode, u32 param, const u8* src, u64 i, u32 hist, u32 cur) {
    if (mode == 0) return cur;
    if (i >= 4) return cur ^ v2_pred(mode, param, hist);       // encoder models use k <= 4, so i >= 4 has a full history
    u32 p;
    return v2_pred_head(mode, param, src, i, p) ? (cur ^ p) : cur;
}

#define V2_NCAND 13
__global__ void __launch_bounds__(KOLM_THREADS) k_v2_hist(const u8* __restrict__ in, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                          u32* __restrict__ gh) {
    constexpr u32 KM[V2_NCAND] = {0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 4, 5, 5}, KP[V2_NCAND] = {0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 0, 0, 1};   // V22.py:1024-1028
    __shared__ u32 h[V2_NCAND * 256];
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    for (u32 e = threadIdx.x; e < V2_NCAND * 256; e += KOLM_THREADS) h[e] = 0;
    __syncthreads();
    const u8* src = in + bi.ioff;
    const u32 t0 = td.start - bi.pbase;
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        const u64 i = (u64)t0 + x;
        const u32 cur = src[i];
        u32 hist = 0;
        if (i >= 4) hist = (u32)src[i - 1] | ((u32)src[i - 2] << 8) | ((u32)src[i - 3] << 16) | ((u32)src[i - 4] << 24);
#pragma unroll
        for (int c = 0; c < V2_NCAND; ++c) atomicAdd(&h[c * 256 + v2_forward_byte(KM[c], KP[c], src, i, hist, cur)], 1u);
    }
    __syncthreads();
    u32* g = gh + (size_t)td.block * (V2_NCAND * 256);
    for (u32 e = threadIdx.x; e < V2_NCAND * 256; e += KOLM_THREADS) if (h[e]) atomicAdd(g + e, h[e]);
}

// params[4*block] = mode, param.  planes: [block][plane][L] 0/1 bytes at 8*(ioff-off0); rawoff[8*block+j]: packed plane j in rawpack
__global__ void __launch_bounds__(KOLM_THREADS) k_v2_planes(const u8* __restrict__ in, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                            const int* __restrict__ params, i64 off0, u8* __restrict__ planes, u8* __restrict__ rawpack,
                                                            const i64* __restrict__ rawoff) {
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    const u32 mode = (u32)params[4 * td.block], param = (u32)params[4 * td.block + 1];
    const u8* src = in + bi.ioff;
    const u32 t0 = td.start - bi.pbase, lane = threadIdx.x & 31;
    const u64 L = bi.len;
    u8* P0 = planes + 8 * (bi.ioff - off0);
#pragma unroll 1
    for (u32 it = 0; it < KOLM_IPT; ++it) {
        const u32 x = it * KOLM_THREADS + threadIdx.x;
        const bool valid = x < td.count;
        const u64 i = (u64)t0 + x;
        u32 m = 0;
        if (valid) {
            u32 hist = 0;
            if (i >= 4) hist = (u32)src[i - 1] | ((u32)src[i - 2] << 8) | ((u32)src[i - 3] << 16) | ((u32)src[i - 4] << 24);
            m = v2_forward_byte(mode, param, src, i, hist, src[i]);
        }
        const u64 iw = i - lane;                                   // first position of this warp's 32 (a multiple of 32 inside the block)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const u32 bit = (m >> (7 - j)) & 1u;
            if (valid) P0[(u64)j * L + i] = (u8)bit;
            const u32 w = __brev(__ballot_sync(0xffffffffu, valid && bit));   // lane l -> bit 31-l: byte 3-k holds lanes 8k..8k+7, MSB first
            if (lane < 4 && iw + 8 * lane < L && iw < (u64)t0 + td.count) rawpack[rawoff[8 * td.block + j] + (iw >> 3) + lane] = (u8)(w >> (8 * (3 - lane)));
        }
    }
}

// Planes batch (block = one transformed plane): tacc[tile*32 + k] = sum over the runs ending in the tile of (len >> k) + 1 + k,
// k < 16; slot 16 = runs; slot 17 = the plane's first bit (first tile only).
__global__ void __launch_bounds__(KOLM_THREADS, 4) k_v2_runs_cost(const u8* __restrict__ b, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                                 u64* lb, u64* __restrict__ tacc) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_last[KOLM_THREADS / 32];
    __shared__ u64 s_excl;
    __shared__ unsigned long long s_acc[18];
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase;
    const u8* src = b + bi.ioff + t0;
    if (tid < 18) s_acc[tid] = 0;
    u32 v[KOLM_IPT + 1];
    load_items(src, t0, td.count, bi.len, v);
    const u32 r0 = tid * KOLM_IPT;
    const u32 prev = (r0 < td.count && t0 + r0 > 0) ? (u32)src[(int)r0 - 1] : 0x100u;
    u32 laststart = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { const u32 r = r0 + i; if (r < td.count && v[i] != (i ? v[i - 1] : prev)) laststart = t0 + r + 1; }
    u32 ln = scan_last_nonzero(laststart, lb, tile, (td.flags & 1u) != 0, s_warp, s_last, &s_excl);
    u32 acc[16]; u32 nruns = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        const u32 r = r0 + i;
        if (r < td.count) {
            const u32 pos1 = t0 + r + 1;
            if (v[i] != (i ? v[i - 1] : prev)) ln = pos1;
            if (v[i + 1] != v[i]) {                                 // the run ends here (V_END differs from every symbol)
                const u32 len = pos1 - ln + 1;
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[k] += len >> k;
                ++nruns;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const u32 x = __reduce_add_sync(0xffffffffu, acc[k] + nruns * (1 + k));   // runs of one plane are disjoint: sums stay below 2^32
        if ((tid & 31) == 0 && x) atomicAdd(&s_acc[k], (unsigned long long)x);
    }
    { const u32 x = __reduce_add_sync(0xffffffffu, nruns); if ((tid & 31) == 0 && x) atomicAdd(&s_acc[16], (unsigned long long)x); }
    if (tid == 0 && t0 == 0) s_acc[17] = v[0] & 1u;
    __syncthreads();
    if (tid < 32) tacc[(size_t)tile * 32 + tid] = tid < 18 ? s_acc[tid] : 0ull;
}

struct V2Hdr { i64 dst; u32 n; u8 bytes[16]; u32 pad; };
__global__ void k_v2_headers(const V2Hdr* __restrict__ h, u8* __restrict__ out, int nblocks) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    for (u32 i = 0; i < h[b].n; ++i) out[h[b].dst + i] = h[b].bytes[i];
}

// Planes batch.  params[4*p] = k, raw flag; poff[p] = byte offset of the plane's chunk in out; rawoff[p] = its packed bytes in rawpack.
__global__ void __launch_bounds__(KOLM_THREADS, 4) k_v2_pack(const u8* __restrict__ b, const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                            u64* lb, const int* __restrict__ params, const i64* __restrict__ poff,
                                                            const u8* __restrict__ rawpack, const i64* __restrict__ rawoff, u32* __restrict__ out) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_last[KOLM_THREADS / 32];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    u64* lb2 = lb + gridDim.x;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    const u32 t0 = td.start - bi.pbase;
    const u32 k = (u32)params[4 * td.block];
    if (params[4 * td.block + 1]) {                              // RAW plane: this tile's packed bytes (tiles start on multiples of 4096 bits)
        const u8* s8 = rawpack + rawoff[td.block] + (t0 >> 3);
        u8* d8 = reinterpret_cast<u8*>(out) + poff[td.block] + (t0 >> 3);
        for (u32 x = tid; x < (td.count + 7) / 8; x += KOLM_THREADS) d8[x] = s8[x];
        return;
    }
    const u8* src = b + bi.ioff + t0;
    u32 v[KOLM_IPT + 1];
    load_items(src, t0, td.count, bi.len, v);
    const u32 r0 = tid * KOLM_IPT;
    const u32 prev = (r0 < td.count && t0 + r0 > 0) ? (u32)src[(int)r0 - 1] : 0x100u;
    u32 laststart = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) { const u32 r = r0 + i; if (r < td.count && v[i] != (i ? v[i - 1] : prev)) laststart = t0 + r + 1; }
    u32 ln = scan_last_nonzero(laststart, lb, tile, (td.flags & 1u) != 0, s_warp, s_last, &s_excl);
    u64 mybits = 0; u32 tokmask = 0;
    u32 lenv[KOLM_IPT];
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        const u32 r = r0 + i;
        lenv[i] = 0;
        if (r < td.count) {
            const u32 pos1 = t0 + r + 1;
            if (v[i] != (i ? v[i - 1] : prev)) ln = pos1;
            lenv[i] = pos1 - ln + 1;
            if (v[i + 1] != v[i]) { mybits += (u64)(lenv[i] >> k) + 1 + k; tokmask |= 1u << i; }
        }
    }
    u64 btot;
    u64 bincl = block_scan_incl(mybits, 0ull, OpAdd(), s_warp, &btot);
    if (tid < 32) {
        u64 e = lb_exclusive(lb2, tile, (td.flags & 1u) != 0, btot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    const u64 bitbase = (u64)poff[td.block] * 8;
    const u64 bp0 = bitbase + s_excl + (bincl - mybits);
    __shared__ u32 s_stage[STAGE_WORDS];
    BitStage st;
    st.begin(s_stage, out, bitbase + s_excl, btot);
    BitAcc ba; ba.init(st, bp0);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if ((tokmask >> i) & 1u) {
            const u32 len = lenv[i], q = len >> k; const u64 n = (u64)q + 1 + k;
            if (n <= 32 && st.use) ba.push(st, ((((u32)1 << q) - 1u) << (k + 1)) | (len & ((1u << k) - 1u)), (u32)n);
            else {
                ba.finish(st);
                u64 bp = ba.bitpos(st);
                long_rice_tok(&st, bp, 0, 0, len, k);
                ba.init(st, bp + n);
            }
        }
    }
    ba.finish(st);
    st.flush();
}

int kolm_bbwt_fwd_impl(kolm_ctx* c, const u8* in, u8* out, int* rounds_plain, int* rounds_cyclic, cudaStream_t s);
int kolm_lb_reset_mode(kolm_ctx* c, bool active, int ntiles, int* grid, int mode, cudaStream_t s);

// H0 of one residual histogram exactly as zero_order_entropy_bits_per_byte (V22.py:631-643): fp64, index order, p = f / n,
// H -= p * log2(p) with the product rounded before the subtraction (no fused multiply-add).
static double v2_h0_hist(const u32* f, i64 n) {
    if (!n) return 0.0;
    const double nn = (double)n;
    double H = 0.0;
    for (int v = 0; v < 256; ++v) if (f[v]) { const double p = (double)f[v] / nn; volatile double t = p * log2(p); H -= t; }
    return H;
}

// in / out: device.  off / out_off: host.  Same context requirements as kolm_v2new_dec (8x bytes, 8x blocks).
int kolm_v2new_enc_impl(kolm_ctx* c, const u8* in, const i64* off, int nblocks, u8* out, size_t out_cap, i64* out_off, cudaStream_t s) {
    if (nblocks <= 0) { if (out_off) out_off[0] = 0; return KOLM_OK; }
    if (((uintptr_t)out & 3) != 0) return KOLM_E_ARG;
    static const int KM[V2_NCAND] = {0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 4, 5, 5}, KP[V2_NCAND] = {0, 1, 2, 3, 4, 0, 1, 2, 3, 0, 0, 0, 1};
    const i64 off0 = off[0], total = off[nblocks] - off0;
    const int np = 8 * nblocks;
    if (np > c->max_blocks || (size_t)(8 * total) + (size_t)KOLM_PAD * np + 2 * KOLM_PAD > c->max_elems) return KOLM_E_CAPACITY;
    // scratch in the (lazily allocated) inverse-transform node buffer, idle during encodes: histograms | packed planes | headers
    const size_t jump_bytes = 2 * c->max_elems * 16;
    const size_t hist_bytes = ((size_t)nblocks * V2_NCAND * 256 * 4 + 255) & ~(size_t)255;
    const size_t raw_bytes = ((size_t)total + 8 * (size_t)np + 255) & ~(size_t)255;
    if (hist_bytes + raw_bytes > jump_bytes || (size_t)nblocks * sizeof(V2Hdr) > hist_bytes) return KOLM_E_CAPACITY;
    if (!c->d_jump) CUDA_TRY(cudaMalloc((void**)&c->d_jump, jump_bytes));
    u32* gh = (u32*)c->d_jump;
    u8* rawpack = (u8*)c->d_jump + hist_bytes;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    const int nt = c->ntiles;
    // ---- model choice
    std::vector<u32> hist((size_t)nblocks * V2_NCAND * 256, 0u);
    if (nt) {
        CUDA_TRY(cudaMemsetAsync(gh, 0, hist_bytes, s));
        KL(c, KC_MISC, total * 14, s, k_v2_hist<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, gh));
        CUDA_TRY(cudaMemcpyAsync(hist.data(), gh, hist.size() * 4, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<int> bmode(nblocks), bparam(nblocks);
    std::vector<i64> poffs((size_t)np + 1);
    i64 rp = 0;
    for (int b = 0; b < nblocks; ++b) {
        const i64 L = off[b + 1] - off[b];
        int bm = 0, bp = 0; double bh = 0.0;
        for (int cnd = 0; cnd < V2_NCAND && L; ++cnd) {                                   // _best_choice / _pick_better, V22.py:936-946, 1004-1009
            const double h = v2_h0_hist(hist.data() + ((size_t)b * V2_NCAND + cnd) * 256, L);
            bool take = cnd == 0;
            if (!take) {
                if (h < bh - 1e-12) take = true;
                else if (fabs(h - bh) <= 1e-12) { if (KM[cnd] < bm) take = true; else if (KM[cnd] == bm && KP[cnd] < bp) take = true; }
            }
            if (take) { bm = KM[cnd]; bp = KP[cnd]; bh = h; }
        }
        bmode[b] = bm; bparam[b] = bp;
        c->h_params[4 * b] = bm; c->h_params[4 * b + 1] = bp; c->h_params[4 * b + 2] = 0; c->h_params[4 * b + 3] = 0;
        const i64 rl = (L + 7) / 8, base = 8 * (off[b] - off0);
        for (int j = 0; j < 8; ++j) { c->h_sizes[(size_t)8 * b + j] = rp; rp += rl; poffs[(size_t)8 * b + j] = base + j * L; }
    }
    poffs[(size_t)np] = 8 * total;
    if ((size_t)np > (size_t)c->max_blocks * 5) return KOLM_E_CAPACITY;
    CUDA_TRY(cudaMemcpyAsync(c->d_params, c->h_params, (size_t)nblocks * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->d_sizes, c->h_sizes, (size_t)np * 8, cudaMemcpyHostToDevice, s));
    if (nt) KL(c, KC_MISC, total * 10, s, k_v2_planes<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_params, off0, c->d_tmp8a, rawpack, c->d_sizes));
    CUDA_TRY(cudaGetLastError());
    // ---- BBWT of every plane, run-length costs
    KOLM_TRY(kolm_set_batch(c, poffs.data(), np, s));
    const int pt = c->ntiles;
    int rpl = 0, rcy = 0;
    KOLM_TRY(kolm_bbwt_fwd_impl(c, c->d_tmp8a, c->d_tmp8b, &rpl, &rcy, s));
    c->counters[0] = rpl; c->counters[1] = rcy;
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)np * 64 * 8, s));
    int lgrid = pt;
    if (pt) {
        KOLM_TRY(kolm_lb_reset_mode(c, false, pt, &lgrid, 1, s));
        KL(c, KC_RICE_COST, 8 * total, s, k_v2_runs_cost<<<lgrid, KOLM_THREADS, 0, s>>>(c->d_tmp8b, c->d_tiles, c->d_binfo, c->d_lb, (u64*)c->d_thist));
        KL(c, KC_RICE_COST, (i64)pt * 256, s, k_tile_reduce<<<np, 256, 0, s>>>((const u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 0, 18));
    }
    CUDA_TRY(cudaMemcpyAsync(c->h_bacc, c->d_bacc, (size_t)np * 64 * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    // ---- per-plane decision and payload layout (V22.py:1529-1574)
    std::vector<V2Hdr> hdr(nblocks);
    i64 o = 0;
    for (int b = 0; b < nblocks; ++b) {
        const i64 L = off[b + 1] - off[b];
        out_off[b] = o;
        V2Hdr& h = hdr[b]; h.dst = o; h.n = 0;
        if (!L) { for (int j = 0; j < 8; ++j) { c->h_params[4 * (8 * b + j)] = 0; c->h_params[4 * (8 * b + j) + 1] = 1; c->h_poff[8 * b + j] = o; } continue; }
        const u32 mp = (u32)bparam[b];
        const int plen = mp == 0 ? 0 : mp <= 0xFF ? 1 : mp <= 0xFFFF ? 2 : mp <= 0xFFFFFF ? 3 : 4;
        u32 raw_mask = 0, b1_mask = 0; int ks[8]; i64 psize[8];
        const i64 rl = (L + 7) / 8;
        for (int j = 0; j < 8; ++j) {
            const u64* a = c->h_bacc + (size_t)(8 * b + j) * 64;
            int kb = 0; i64 best = -1;
            for (int k = 0; k < 16; ++k) { const i64 bytes = (i64)((a[k] + 7) >> 3); if (best < 0 || bytes < best) { best = bytes; kb = k; } }   // _choose_best_rice: strict <
            if (rl <= best + 1) { raw_mask |= 1u << j; psize[j] = rl; ks[j] = 0; }
            else { if (a[17] & 1) b1_mask |= 1u << j; psize[j] = best; ks[j] = kb; }
        }
        h.bytes[h.n++] = (u8)(((u32)bmode[b] & 7u) << 5 | (u32)plen);
        for (int i = 0; i < plen; ++i) h.bytes[h.n++] = (u8)(mp >> (8 * i));
        h.bytes[h.n++] = (u8)raw_mask; h.bytes[h.n++] = (u8)b1_mask;
        for (int j = 0; j < 8; ++j) if (!((raw_mask >> j) & 1u)) h.bytes[h.n++] = (u8)ks[j];
        o += h.n;
        for (int j = 0; j < 8; ++j) {
            c->h_params[4 * (8 * b + j)] = ks[j]; c->h_params[4 * (8 * b + j) + 1] = (raw_mask >> j) & 1; c->h_params[4 * (8 * b + j) + 2] = 0; c->h_params[4 * (8 * b + j) + 3] = 0;
            c->h_poff[8 * b + j] = o; o += psize[j];
        }
    }
    out_off[nblocks] = o; c->h_poff[np] = o;
    if ((size_t)o > out_cap) return KOLM_E_CAPACITY;
    V2Hdr* d_hdr = (V2Hdr*)c->d_jump;                         // the histograms are dead
    CUDA_TRY(cudaMemcpyAsync(d_hdr, hdr.data(), (size_t)nblocks * sizeof(V2Hdr), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->d_params, c->h_params, (size_t)np * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, s));
    KL(c, KC_ZERO, 0, s, k_zero_words<<<4 * c->sm_count, 256, 0, s>>>((u32*)out, c->d_poff + np, out_cap / 4));
    KL(c, KC_MISC, (i64)nblocks * 16, s, k_v2_headers<<<(nblocks + 127) / 128, 128, 0, s>>>(d_hdr, out, nblocks));
    if (pt) {
        KOLM_TRY(kolm_lb_reset_mode(c, false, pt, &lgrid, 1, s));
        KL(c, KC_RICE_PACK, 8 * total + o, s, k_v2_pack<<<lgrid, KOLM_THREADS, 0, s>>>(c->d_tmp8b, c->d_tiles, c->d_binfo, c->d_lb, c->d_params, c->d_poff, rawpack, c->d_sizes, (u32*)out));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s));                       // hdr / poffs (host vectors) were sources of async copies
    return KOLM_OK;
}
