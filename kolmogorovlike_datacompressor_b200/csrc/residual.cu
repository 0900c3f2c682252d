// residual.cu — the cheap byte-predictor candidates.
//
//   kind 0  XOR    KF  encode_model_xor / decode_model_xor    kolm_final.py:545-565, 702-721     r = b[i] ^ b[i-1]
//   kind 1  DELTA  V22 encode_xor / decode_xor (a subtraction) v2-2.py:2105-2122                  r = b[i] - b[i-1]
//   kind 2  LFSR   V22 encode_lfsr_predict / decode            v2-2.py:1984-2019                  r = b[i] - s[i]
// Every residual is one ULEB128 value: 1 byte if r < 128, else [r, 0x01].  Payload size = n + #{r >= 128}.
// s[i] is data independent (s0 = 1, s' = (s<<1 & 0xFF) | parity(s & 0x96)): a period table in constant memory.
#include "common.cuh"

__constant__ u8 c_lfsr[256];
__constant__ u32 c_lfsr_period;
static bool g_lfsr_ready[64];

static int lfsr_init(kolm_ctx* c) {
    if (c->device < 64 && g_lfsr_ready[c->device]) return KOLM_OK;
    u8 tab[256]; u32 s = 1, n = 0; bool seen[256] = {false};
    while (!seen[s]) { seen[s] = true; tab[n++] = (u8)s; s = ((s << 1) & 0xFF) | (u32)(__builtin_popcount(s & 0x96) & 1); }
    // the orbit of 1 is purely periodic iff it returns to 1; otherwise keep the pre-period too (handled below)
    u32 period = n;
    if (s != 1) return KOLM_E_UNSUPPORTED;
    CUDA_TRY(cudaMemcpyToSymbol(c_lfsr, tab, 256));
    CUDA_TRY(cudaMemcpyToSymbol(c_lfsr_period, &period, 4));
    if (c->device < 64) g_lfsr_ready[c->device] = true;
    return KOLM_OK;
}

__device__ __forceinline__ u32 residual_of(const u8* __restrict__ src, u32 lp, int kind) {
    u32 b = src[lp];
    if (kind == 2) return (b - c_lfsr[lp % c_lfsr_period]) & 0xFF;
    u32 p = lp ? src[lp - 1] : 0;
    return kind == 0 ? (b ^ p) : ((b - p) & 0xFF);
}

// sizes of all three kinds in one read: bacc[b*64 + 40 + kind] += #{r >= 128}
__global__ void __launch_bounds__(KOLM_THREADS) k_res_cost(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u64* __restrict__ bacc) {
    __shared__ u32 s_cnt[3];
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    u32 c0 = 0, c1 = 0, c2 = 0;
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 lp = td.start + x - bi.pbase;
        u32 b = src[lp], p = lp ? src[lp - 1] : 0;
        c0 += ((b ^ p) >= 128u); c1 += (((b - p) & 0xFF) >= 128u); c2 += (((b - c_lfsr[lp % c_lfsr_period]) & 0xFF) >= 128u);
    }
    for (int o = 16; o > 0; o >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], c0); atomicAdd(&s_cnt[1], c1); atomicAdd(&s_cnt[2], c2); }
    __syncthreads();
    if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd((unsigned long long*)(bacc + (size_t)td.block * 64 + 40 + threadIdx.x), (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void k_res_plan_sizes(u64* __restrict__ bacc, const BlockInfo* __restrict__ binfo, int nblocks, int kind) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblocks) bacc[(size_t)b * 64 + 32] = binfo[b].len + bacc[(size_t)b * 64 + 40 + kind];
}

__global__ void __launch_bounds__(KOLM_THREADS) k_res_emit(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u64* lb, const u64* __restrict__ bacc,
                                                           u8* __restrict__ out, int kind, const i64* __restrict__ cap_total, u64 cap,
                                                           const int* __restrict__ method, int want) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_excl;
    if ((u64)*cap_total > cap) return;                      // exact total known before any byte is emitted: never write past the caller's buffer
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = tiles[tile];
    if (method && method[td.block] != want) return;         // whole blocks drop out together: no tile of another block looks back at them
    const BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff;
    const u32 t0 = td.start - bi.pbase;
    u32 r[KOLM_IPT]; u32 big = 0;
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        u32 x = tid * KOLM_IPT + i;
        r[i] = x < td.count ? residual_of(src, t0 + x, kind) : 0;
        big += (r[i] >= 128u);
    }
    u64 tot;
    u64 incl = block_scan_incl((u64)big, 0ull, OpAdd(), s_warp, &tot);
    if (tid < 32) {
        u64 e = lb_exclusive(lb, tile, (td.flags & 1u) != 0, tot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u8* dst = out + bacc[(size_t)td.block * 64 + 33] + t0 + tid * KOLM_IPT + s_excl + (incl - big);
#pragma unroll
    for (int i = 0; i < KOLM_IPT; ++i) {
        if (tid * KOLM_IPT + i < td.count) { *dst++ = (u8)r[i]; if (r[i] >= 128u) *dst++ = 1; }
    }
}

// decode (v0): one thread per block
__global__ void k_res_dec(const u8* __restrict__ pay, const i64* __restrict__ pay_off, const BlockInfo* __restrict__ binfo, u8* __restrict__ out,
                          int* __restrict__ err, int nblocks, int kind) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    BlockInfo bi = binfo[b];
    const u8* d = pay + pay_off[b];
    i64 n = pay_off[b + 1] - pay_off[b], p = 0;
    u8* dst = out + bi.ioff;
    u32 prev = 0; int e = KOLM_OK;
    for (u32 i = 0; i < bi.len; ++i) {
        u64 r = 0; int sh = 0;
        for (;;) { if (p >= n) { e = KOLM_E_TRUNCATED; break; } u8 x = d[p++]; if (sh < 64) r |= (u64)(x & 0x7F) << sh; if (!(x & 0x80)) break; sh += 7; }
        if (e) break;
        u32 v;
        if (kind == 0) { if (r > 255) { e = KOLM_E_CORRUPT; break; } v = (u32)r ^ prev; }      // bytearray.append(>255) raises ValueError
        else if (kind == 1) v = (prev + (u32)r) & 0xFF;
        else v = ((u32)r + c_lfsr[i % c_lfsr_period]) & 0xFF;
        dst[i] = (u8)v; prev = v;
    }
    err[b] = e;
}

int kolm_residual_sizes_impl(kolm_ctx* c, const u8* in, i64* sizes3, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    KOLM_TRY(lfsr_init(c));
    if (!nb) return KOLM_OK;
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * 64 * 8, s));
    if (nt) KL(c, KC_MISC, c->total_bytes, s, k_res_cost<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_bacc));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_bacc, c->d_bacc, (size_t)nb * 64 * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < nb; ++b) for (int k = 0; k < 3; ++k) sizes3[3 * (size_t)b + k] = (i64)c->h_binfo[b].len + (i64)c->h_bacc[(size_t)b * 64 + 40 + k];
    return KOLM_OK;
}

int kolm_residual_enc_impl(kolm_ctx* c, const u8* in, int kind, u8* out, size_t out_cap, i64* out_off, cudaStream_t s) {
    const int nb = c->nblocks, nt = c->ntiles;
    if (kind < 0 || kind > 2) return KOLM_E_ARG;
    KOLM_TRY(lfsr_init(c));
    if (!nb) { out_off[0] = 0; return KOLM_OK; }
    CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * 64 * 8, s));
    if (nt) KL(c, KC_MISC, c->total_bytes, s, k_res_cost<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_bacc));
    KL(c, KC_MISC, (i64)nb * 16, s, k_res_plan_sizes<<<(nb + 255) / 256, 256, 0, s>>>(c->d_bacc, c->d_binfo, nb, kind));
    KL(c, KC_RICE_PLAN, (i64)nb * 16, s, k_lz_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_poff, nb));
    if (nt) {
        int lgrid = nt;
        KOLM_TRY(kolm_lb_reset_mode(c, false, nt, &lgrid, 1, s));
        KL(c, KC_MISC, c->total_bytes * 2, s, k_res_emit<<<lgrid, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_lb, c->d_bacc, out, kind, c->d_poff + nb, (u64)out_cap, nullptr, 0));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->h_poff, c->d_poff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    memcpy(out_off, c->h_poff, (size_t)(nb + 1) * 8);
    if ((size_t)out_off[nb] > out_cap) return KOLM_E_CAPACITY;
    return KOLM_OK;
}

int kolm_residual_dec_impl(kolm_ctx* c, const u8* pay, const i64* pay_off, int kind, u8* out, cudaStream_t s) {
    const int nb = c->nblocks;
    if (kind < 0 || kind > 2) return KOLM_E_ARG;
    KOLM_TRY(lfsr_init(c));
    if (!nb) return KOLM_OK;
    memcpy(c->h_poff, pay_off, (size_t)(nb + 1) * 8);
    CUDA_TRY(cudaMemcpyAsync(c->d_poff, c->h_poff, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, s));
    KL(c, KC_MISC, c->total_bytes * 2, s, k_res_dec<<<(nb + 63) / 64, 64, 0, s>>>(pay, c->d_poff, c->d_binfo, out, c->d_err, nb, kind));
    CUDA_TRY(cudaGetLastError());
    return rice_dec_finish(c, s);
}
