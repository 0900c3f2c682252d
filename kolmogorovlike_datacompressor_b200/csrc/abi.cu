// abi.cu — extern "C" surface of libkolm_b200.so (include/kolm_abi.h), context and batch setup.
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

static thread_local char g_cuda_err[512] = "";
void kolm_set_cuda_error(cudaError_t e, const char* file, int line) {
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}

// single translation unit: the stage files are included here (no -rdc needed)
#include "bbwt_fwd.cu"
#include "bbwt_inv.cu"
#include "mtf.cu"
#include "rice.cu"
#include "rice2.cu"
#include "rice_dec.cu"
#include "lz77.cu"
#include "residual.cu"
#include "repair.cu"
#include "repair_big.cu"
#include "v2new.cu"

static size_t padded_capacity(size_t max_batch_bytes, int max_blocks) {
    return max_batch_bytes + (size_t)KOLM_PAD * (size_t)max_blocks + 4 * KOLM_PAD;
}

template <class T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

extern "C" {

int kolm_abi_version(void) { return 1; }

const char* kolm_strerror(int code) {
    switch (code) {
        case 0: return "ok";
        case KOLM_E_CUDA: return "CUDA error";
        case KOLM_E_ARG: return "bad argument";
        case KOLM_E_CAPACITY: return "batch exceeds context capacity";
        case KOLM_E_TRUNCATED: return "truncated payload";
        case KOLM_E_CORRUPT: return "corrupt payload";
        case KOLM_E_UNSUPPORTED: return "unsupported";
        case KOLM_E_INDEX: return "index error (reference decoder bug reproduced)";
    }
    return "unknown";
}
const char* kolm_last_cuda_error(void) { return g_cuda_err; }

size_t kolm_scratch_bytes(size_t max_batch_bytes, int max_blocks) {
    size_t e = padded_capacity(max_batch_bytes, max_blocks);
    size_t tiles = e / KOLM_TILE + max_blocks + 2;
    return e * (4 * 10 + 4 / 8 + 2) + tiles * (sizeof(TileDesc) * 2 + 16 + 1024) + (size_t)max_blocks * (sizeof(BlockInfo) + 4 * 9 + 64 * 8);
}

int kolm_create(int device, size_t max_batch_bytes, int max_blocks, kolm_ctx** out) { return kolm_create_ex(device, max_batch_bytes, max_blocks, 0u, out); }

// flags & KOLM_CTX_REPAIR_ONLY: a context for kolm_repair_enc alone — block tables and the two staging arrays it uses (8 instead of
// 38 bytes per element), so that one context can hold a whole container's worth of long blocks next to the slab pool
int kolm_create_ex(int device, size_t max_batch_bytes, int max_blocks, unsigned flags, kolm_ctx** out) {
    if (!out || max_blocks < 1 || max_batch_bytes < 1 || (flags & ~(unsigned)KOLM_CTX_REPAIR_ONLY)) return KOLM_E_ARG;
    const bool light = (flags & KOLM_CTX_REPAIR_ONLY) != 0;
    size_t e = padded_capacity(max_batch_bytes, max_blocks);
    if (e >= (1ull << 31)) return KOLM_E_ARG;     // 31-bit positions / ranks
    CUDA_TRY(cudaSetDevice(device));
    kolm_ctx* c = new kolm_ctx();
    memset(c, 0, sizeof *c);
    c->device = device; c->max_elems = e; c->max_blocks = max_blocks; c->light = light ? 1 : 0;
    c->max_tiles = (int)(e / KOLM_TILE) + max_blocks + 2;
    cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    size_t nb = (size_t)max_blocks, nt = (size_t)c->max_tiles;
    CUDA_TRY(dalloc(&c->d_binfo, nb)); CUDA_TRY(dalloc(&c->d_btile0, nb)); CUDA_TRY(dalloc(&c->d_btilen, nb));
    CUDA_TRY(dalloc(&c->d_atile0, nb)); CUDA_TRY(dalloc(&c->d_atilen, nb)); CUDA_TRY(dalloc(&c->d_active, nb));
    CUDA_TRY(dalloc(&c->d_newcls, nb)); CUDA_TRY(dalloc(&c->d_done, nb)); CUDA_TRY(dalloc(&c->d_nfac, nb));
    CUDA_TRY(dalloc(&c->d_stats, 16)); CUDA_TRY(dalloc(&c->d_bacc, nb * 64));
    CUDA_TRY(dalloc(&c->d_tiles, nt)); CUDA_TRY(dalloc(&c->d_atiles, nt));
    CUDA_TRY(dalloc(&c->d_lb, 32 + 8 * nt)); CUDA_TRY(dalloc(&c->d_thist, nt * 256));
    CUDA_TRY(dalloc(&c->d_k0, e)); CUDA_TRY(dalloc(&c->d_v0, e));
    if (!light) {
        CUDA_TRY(dalloc(&c->d_k1, e)); CUDA_TRY(dalloc(&c->d_v1, e));
        CUDA_TRY(dalloc(&c->d_sa, e)); CUDA_TRY(dalloc(&c->d_rank, e)); CUDA_TRY(dalloc(&c->d_nr, e)); CUDA_TRY(dalloc(&c->d_lo, e));
        CUDA_TRY(dalloc(&c->d_single, e / 32 + 8)); CUDA_TRY(dalloc(&c->d_fstart, e));
        CUDA_TRY(dalloc(&c->d_grp, e)); CUDA_TRY(dalloc(&c->d_live, nt)); CUDA_TRY(dalloc(&c->d_lact, nb));
        CUDA_TRY(dalloc(&c->d_tmp8a, e)); CUDA_TRY(dalloc(&c->d_tmp8b, e));
    }
    CUDA_TRY(cudaMallocHost((void**)&c->h_binfo, nb * sizeof(BlockInfo)));
    CUDA_TRY(cudaMallocHost((void**)&c->h_u32, nb * 4 * sizeof(u32)));
    CUDA_TRY(cudaMallocHost((void**)&c->h_stats, 16 * sizeof(u32)));
    CUDA_TRY(cudaMallocHost((void**)&c->h_bacc, nb * 64 * sizeof(u64)));
    CUDA_TRY(dalloc(&c->d_err, nb)); CUDA_TRY(cudaMallocHost((void**)&c->h_err, nb * sizeof(int)));
    CUDA_TRY(dalloc(&c->d_poff, nb + 1)); CUDA_TRY(dalloc(&c->d_params, nb * 4)); CUDA_TRY(dalloc(&c->d_sizes, nb * 5));
    CUDA_TRY(cudaMallocHost((void**)&c->h_poff, (nb + 1) * sizeof(i64)));
    CUDA_TRY(cudaMallocHost((void**)&c->h_params, nb * 4 * sizeof(int)));
    CUDA_TRY(cudaMallocHost((void**)&c->h_sizes, nb * 5 * sizeof(i64)));
    {   // KOLM_POISON=1 (set by tests/conftest.py): every scratch array starts as 0xA5 bytes instead of whatever the allocation held, so
        // a kernel that reads scratch nothing wrote in the same call fails reproducibly (compute-sanitizer's initcheck is closed on the pool)
        const char* pz = getenv("KOLM_POISON");
        if (pz && atoi(pz)) {
            struct { void* p; size_t n; } arr[] = {
                {c->d_binfo, nb * sizeof(BlockInfo)}, {c->d_btile0, nb * 4}, {c->d_btilen, nb * 4}, {c->d_atile0, nb * 4}, {c->d_atilen, nb * 4},
                {c->d_active, nb * 4}, {c->d_newcls, nb * 4}, {c->d_done, nb * 4}, {c->d_nfac, nb * 4}, {c->d_stats, 64}, {c->d_bacc, nb * 512},
                {c->d_tiles, nt * sizeof(TileDesc)}, {c->d_atiles, nt * sizeof(TileDesc)}, {c->d_lb, (32 + 8 * nt) * 8}, {c->d_thist, nt * 1024},
                {c->d_k0, e * 4}, {c->d_v0, e * 4}, {c->d_k1, e * 4}, {c->d_v1, e * 4}, {c->d_sa, e * 4}, {c->d_rank, e * 4}, {c->d_nr, e * 4},
                {c->d_lo, e * 4}, {c->d_grp, e * 4}, {c->d_live, nt}, {c->d_lact, nb * 4}, {c->d_single, (e / 32 + 8) * 4}, {c->d_fstart, e * 4}, {c->d_tmp8a, e}, {c->d_tmp8b, e}, {c->d_err, nb * 4},
                {c->d_poff, (nb + 1) * 8}, {c->d_params, nb * 16}, {c->d_sizes, nb * 40}};
            for (auto& a : arr) if (a.p) CUDA_TRY(cudaMemset(a.p, 0xA5, a.n));
        }
    }
    c->prof_ev = new cudaEvent_t[2 * KOLM_PROF_MAX];
    for (int i = 0; i < 2 * KOLM_PROF_MAX; ++i) CUDA_TRY(cudaEventCreate(&c->prof_ev[i]));
    *out = c;
    return KOLM_OK;
}

void kolm_destroy(kolm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    void* dptrs[] = {c->d_binfo, c->d_btile0, c->d_btilen, c->d_atile0, c->d_atilen, c->d_active, c->d_newcls, c->d_done, c->d_nfac,
                     c->d_stats, c->d_bacc, c->d_tiles, c->d_atiles, c->d_lb, c->d_thist, c->d_k0, c->d_v0, c->d_k1, c->d_v1,
                     c->d_sa, c->d_rank, c->d_nr, c->d_lo, c->d_grp, c->d_live, c->d_lact, c->d_single, c->d_fstart, c->d_tmp8a, c->d_tmp8b, c->d_poff, c->d_params, c->d_sizes, c->d_jump, c->d_err, c->d_rpb};
    for (void* p : dptrs) if (p) cudaFree(p);
    if (c->prof_ev) { for (int i = 0; i < 2 * KOLM_PROF_MAX; ++i) cudaEventDestroy(c->prof_ev[i]); delete[] c->prof_ev; }
    if (c->h_binfo) cudaFreeHost(c->h_binfo);
    if (c->h_u32) cudaFreeHost(c->h_u32);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->h_bacc) cudaFreeHost(c->h_bacc);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->h_poff) cudaFreeHost(c->h_poff);
    if (c->h_params) cudaFreeHost(c->h_params);
    if (c->h_sizes) cudaFreeHost(c->h_sizes);
    delete c;
}

}  // extern "C"

// Describe the batch: block table, padded index space, static tile map.
int kolm_set_batch(kolm_ctx* c, const i64* off, int nblocks, cudaStream_t s) {
    if (!c || !off || nblocks < 0) return KOLM_E_ARG;
    if (nblocks > c->max_blocks) return KOLM_E_CAPACITY;
    if (c->light && !c->light_ok) return KOLM_E_UNSUPPORTED;   // a Re-Pair-only context has no scratch for the other operators
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(s));           // pinned staging below may still be in flight from the previous call
    u64 p = 0; u32 t = 0, maxlen = 0, rows = 0;
    u32* bt0 = c->h_u32; u32* btn = c->h_u32 + nblocks;
    for (int b = 0; b < nblocks; ++b) {
        i64 len = off[b + 1] - off[b];
        if (len < 0 || len >= (1ll << 30)) return KOLM_E_ARG;
        c->h_binfo[b].ioff = off[b]; c->h_binfo[b].pbase = (u32)p; c->h_binfo[b].len = (u32)len;
        bt0[b] = t; btn[b] = (u32)((len + KOLM_TILE - 1) / KOLM_TILE); t += btn[b];
        if (btn[b] > rows) rows = btn[b];
        p += len ? ((u64)len + KOLM_PAD - 1) / KOLM_PAD * KOLM_PAD : KOLM_PAD;   // empty blocks still own a scratch slot
        if ((u32)len > maxlen) maxlen = (u32)len;
        if (p + 2 * KOLM_PAD > c->max_elems) return KOLM_E_CAPACITY;
    }
    if ((int)t > c->max_tiles) return KOLM_E_CAPACITY;
    c->nblocks = nblocks; c->total_elems = (u32)p; c->ntiles = (int)t; c->max_len = maxlen; c->static_rows = rows; c->active_rows = 0;
    c->total_bytes = nblocks ? off[nblocks] - off[0] : 0;
    if (!nblocks) return KOLM_OK;
    CUDA_TRY(cudaMemcpyAsync(c->d_binfo, c->h_binfo, (size_t)nblocks * sizeof(BlockInfo), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->d_btile0, bt0, (size_t)nblocks * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->d_btilen, btn, (size_t)nblocks * 4, cudaMemcpyHostToDevice, s));
    if (t) {
        int g = nblocks < 1024 ? nblocks : 1024;
        KL(c, KC_TILES, (i64)t * 16, s, k_build_tiles<<<g, 128, 0, s>>>(c->d_binfo, c->d_btile0, c->d_btilen, nullptr, c->d_tiles, nblocks));
        CUDA_TRY(cudaGetLastError());
    }
    return KOLM_OK;
}

static const char* KC_NAMES[KC_COUNT] = {"build_tiles", "boot_keys", "radix_hist", "radix_scan", "radix_scatter", "rerank", "apply_ranks",
                                        "gather", "plan", "lyndon_scan", "bbwt_emit", "mtf_pre", "mtf_scan", "mtf_main", "rice_cost",
                                        "rice_plan", "rice_pack", "zero_fill", "bbwt_inverse", "misc"};

extern "C" {

// KOLM_RICE_V1=1: the first (CTA-per-tile) Rice encoders of rice.cu instead of the warp-per-tile ones of rice2.cu (A/B runs)
static bool rice_v1() { static int v = -1; if (v < 0) { const char* e = getenv("KOLM_RICE_V1"); v = e ? atoi(e) : 0; } return v != 0; }

int kolm_lyndon(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* start_flags, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    int r = 0;
    KOLM_TRY(kolm_lyndon_impl(c, in, start_flags, &r, s));
    c->counters[0] = r; c->counters[1] = 0;
    return KOLM_OK;
}

int kolm_bbwt_fwd(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    int rp = 0, rc = 0;
    KOLM_TRY(kolm_bbwt_fwd_impl(c, in, out, &rp, &rc, s));
    c->counters[0] = rp; c->counters[1] = rc;
    return KOLM_OK;
}

int kolm_bbwt_inv(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_bbwt_inv_impl(c, in, out, s);
}

int kolm_mtf_enc(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_mtf_impl(c, in, out, false, s);
}

int kolm_mtf_dec(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_mtf_impl(c, in, out, true, s);
}

int kolm_rice_kf_enc(kolm_ctx* c, const uint8_t* mtf, const int64_t* off, int nblocks, uint8_t* out, size_t out_cap, int64_t* out_off,
                     int* params, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    if (rice_v1()) return kolm_rice_kf_enc_impl(c, mtf, out, out_cap, out_off, params, s);
    return kolm_rice2_kf_enc_impl(c, mtf, out, out_cap, out_off, params, s);
}

int kolm_rice_kf_dec(kolm_ctx* c, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint8_t* mtf_out,
                     kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_rice_kf_dec_impl(c, payload, pay_off, mtf_out, s);
}

int kolm_rice_k2_enc(kolm_ctx* c, const uint8_t* mtf, const int64_t* off, int nblocks, int flags, uint8_t* out, size_t out_cap,
                     int64_t* out_off, int64_t* sizes, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    if (rice_v1()) return kolm_rice_k2_enc_impl(c, mtf, flags, out, out_cap, out_off, sizes, s);
    return kolm_rice2_k2_enc_impl(c, mtf, flags, out, out_cap, out_off, sizes, s);
}

int kolm_rice_dual_enc(kolm_ctx* c, const uint8_t* mtf, const int64_t* off, int nblocks, int k2_flags, uint8_t* kf_out, size_t kf_cap,
                       int64_t* kf_off, int* kf_params, uint8_t* k2_out, size_t k2_cap, int64_t* k2_off, int64_t* k2_sizes, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!kf_off || !k2_off) return KOLM_E_ARG;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_rice2_dual_enc_impl(c, mtf, k2_flags, kf_out, kf_cap, kf_off, kf_params, k2_out, k2_cap, k2_off, k2_sizes, s);
}

int kolm_rice_k2_dec(kolm_ctx* c, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, int flags,
                     uint8_t* mtf_out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_rice_k2_dec_impl(c, payload, pay_off, flags, mtf_out, s);
}

int kolm_lz77_enc(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint32_t window, uint32_t max_len, uint8_t* out,
                  size_t out_cap, int64_t* out_off, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_lz77_enc_impl(c, in, window, max_len, out, out_cap, out_off, s);
}

int kolm_lz77_dec(kolm_ctx* c, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint32_t window_check,
                  uint8_t* out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_lz77_dec_impl(c, payload, pay_off, window_check, out, s);
}

int kolm_residual_sizes(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, int64_t* sizes3, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_residual_sizes_impl(c, in, sizes3, s);
}

int kolm_residual_enc(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, int kind, uint8_t* out, size_t out_cap,
                      int64_t* out_off, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_residual_enc_impl(c, in, kind, out, out_cap, out_off, s);
}

int kolm_residual_dec(kolm_ctx* c, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, int kind,
                      uint8_t* out, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_residual_dec_impl(c, payload, pay_off, kind, out, s);
}

int kolm_repair_enc(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, size_t out_cap, int64_t* out_off,
                    kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!c) return KOLM_E_ARG;
    c->light_ok = 1;
    const int rc = kolm_set_batch(c, off, nblocks, s);
    c->light_ok = 0;
    KOLM_TRY(rc);
    return kolm_repair_enc_impl(c, in, out, out_cap, out_off, s);
}

int kolm_repair_dec(kolm_ctx* c, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint8_t* out,
                    kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    return kolm_repair_dec_impl(c, payload, pay_off, out, s);
}

int kolm_repair_max_block(void) { return REPAIR_XL; }          // shared-memory kernel; longer blocks take the incremental kernel (repair_big.cu)

int kolm_v2new_enc(kolm_ctx* c, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, size_t out_cap, int64_t* out_off,
                   kolm_stream_t stream) {
    if (!c || !off || !out_off || nblocks < 0) return KOLM_E_ARG;
    return kolm_v2new_enc_impl(c, in, off, nblocks, out, out_cap, out_off, (cudaStream_t)stream);
}

int kolm_v2new_dec(kolm_ctx* c, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint8_t* out,
                   kolm_stream_t stream) {
    if (!c || !pay_off || !off || nblocks < 0) return KOLM_E_ARG;
    return kolm_v2new_dec_impl(c, payload, pay_off, off, nblocks, out, (cudaStream_t)stream);
}

int kolm_last_counters(kolm_ctx* c, int64_t* out4) {
    i64 total = 0;
    for (int i = 0; i < KC_COUNT; ++i) total += c->launches[i];
    out4[0] = c->counters[0]; out4[1] = c->counters[1]; out4[2] = total; out4[3] = c->counters[4];
    return KOLM_OK;
}

int kolm_profile_categories(void) { return KC_COUNT; }
const char* kolm_profile_name(int cat) { return (cat >= 0 && cat < KC_COUNT) ? KC_NAMES[cat] : ""; }
int kolm_profile_enable(kolm_ctx* c, int on) { c->prof_on = on; return KOLM_OK; }
int kolm_profile_reset(kolm_ctx* c) {
    memset(c->launches, 0, sizeof c->launches); memset(c->algbytes, 0, sizeof c->algbytes); memset(c->counters, 0, sizeof c->counters);
    c->prof_n = 0;
    return KOLM_OK;
}
int kolm_profile_read(kolm_ctx* c, double* ms, int64_t* launches, int64_t* algbytes) {
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < KC_COUNT; ++i) { ms[i] = 0; launches[i] = c->launches[i]; algbytes[i] = c->algbytes[i]; }
    for (int i = 0; i < c->prof_n; ++i) {
        float t = 0; CUDA_TRY(cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        ms[c->prof_cat[i]] += t;
    }
    return KOLM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// content-defined chunking (host side; SURVEY §8f row 1 — adjacent to the hot path, must be bit-identical
// to the reference or every later byte of the container differs)
//   KF  cdc_fast_boundaries        kolm_final.py:161-194, gear = random.Random(2025).getrandbits(32) x 256 (:148-159)
//   V22 cdc_fast_boundaries_strict kolm_final_researched_v2-2.py:210-309, gear = xorshift32(0x243F6A88) | 1 (:152-167)
// ------------------------------------------------------------------------------------------------
static void kf_gear_table(u32* g) {
    // MT19937 seeded the way CPython seeds random.Random(int): init_by_array([seed])
    const int N = 624, M = 397; static u32 mt[624];
    mt[0] = 19650218u; for (int i = 1; i < N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (u32)i;
    { u32 key = 2025u; int i = 1;
      for (int k = N; k; --k) { mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key; if (++i >= N) { mt[0] = mt[N - 1]; i = 1; } }
      for (int k = N - 1; k; --k) { mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (u32)i; if (++i >= N) { mt[0] = mt[N - 1]; i = 1; } }
      mt[0] = 0x80000000u; }
    int idx = N;
    for (int t = 0; t < 256; ++t) {
        if (idx >= N) {
            for (int k = 0; k < N; ++k) { u32 y = (mt[k] & 0x80000000u) | (mt[(k + 1) % N] & 0x7fffffffu); mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
            idx = 0;
        }
        u32 y = mt[idx++]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18; g[t] = y;
    }
}

extern "C" {

// ends[i] = exclusive end of chunk i; returns the number of chunks, or <0
int64_t kolm_cdc_kf(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, int64_t* ends, int64_t cap) {
    static u32 G[256]; static bool init = false;
    if (!init) { kf_gear_table(G); init = true; }
    if (n <= 0) return 0;
    int bl = 0; for (i64 a = avg_size; a > 0; a >>= 1) ++bl;
    int k = bl - 1; if (k > 20) k = 20; if (k < 6) k = 6;
    const u32 mask = (1u << k) - 1u;
    i64 i = 0, cnt = 0;
    while (i < n) {
        i64 start = i; u32 h = 0;
        i64 emin = start + min_size < n ? start + min_size : n, emax = start + max_size < n ? start + max_size : n;
        i = emin;
        while (i < emax) { h = (h << 1) + G[data[i]]; ++i; if ((h & mask) == 0) break; }
        if (cnt >= cap) return KOLM_E_CAPACITY;
        ends[cnt++] = i;
        if (i == start) return KOLM_E_ARG;            // min_size = max_size = 0 would never advance
    }
    return cnt;
}

int64_t kolm_cdc_v22(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, int64_t* ends, int64_t cap) {
    static u32 G[256]; static bool init = false;
    if (!init) { u32 x = 0x243F6A88u; for (int i = 0; i < 256; ++i) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; G[i] = x | 1u; } init = true; }
    if (n <= 0) return 0;
    if (!(min_size > 0 && min_size <= avg_size && avg_size <= max_size) || avg_size < 64) return KOLM_E_ARG;
    int bl = 0; for (i64 a = avg_size; a > 0; a >>= 1) ++bl;
    int k = bl - 1; if (k < 6) k = 6; if (k > 20) k = 20;
    const int ks = (k + 2 <= 20) ? k + 2 : 20, kl = (k > 2) ? k - 2 : 1;
    const u32 ms = (1u << ks) - 1u, ml = (1u << kl) - 1u;
    i64 i = 0, cnt = 0, last_start = 0;
    while (i < n) {
        i64 start = i, rem = n - start; last_start = start;
        if (cnt >= cap) return KOLM_E_CAPACITY;
        if (rem <= min_size) { ends[cnt++] = n; break; }
        i64 lmax = rem < max_size ? rem : max_size, normal = avg_size < lmax ? avg_size : lmax;
        i64 pos = start + min_size, en = start + normal, el = start + lmax; u32 fp = 0; bool found = false;
        while (pos < en && pos < el) { fp = (fp << 1) + G[data[pos]]; ++pos; if ((fp & ms) == 0) { found = true; break; } }
        if (!found) while (pos < el) { fp = (fp << 1) + G[data[pos]]; ++pos; if ((fp & ml) == 0) { found = true; break; } }
        if (!found) pos = el;
        ends[cnt++] = pos; i = pos;
    }
    if (cnt >= 2 && ends[cnt - 1] - last_start < min_size) { ends[cnt - 2] = ends[cnt - 1]; --cnt; }   // orphan tail merged (V22.py:301-306)
    return cnt;
}

}  // extern "C"

#include "cdc.cu"


// ------------------------------------------------------------------------------------------------
// payload gather (device part): after model selection the winners' payloads live in different
// per-model buffers; one kernel lays them out back to back in block order (= the container's payload area).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_payloads(const u64* __restrict__ src, const i64* __restrict__ len, const i64* __restrict__ dst_off,
                                                         u8* __restrict__ out, int nblocks) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const u8* s = reinterpret_cast<const u8*>(src[b]);
        u8* d = out + dst_off[b];
        const i64 n = len[b];
        // 16-byte body when source and destination are mutually aligned, bytes otherwise
        if ((((uintptr_t)s ^ (uintptr_t)d) & 15) == 0) {
            i64 head = (16 - ((uintptr_t)s & 15)) & 15; if (head > n) head = n;
            for (i64 i = threadIdx.x; i < head; i += blockDim.x) d[i] = s[i];
            const i64 body = (n - head) >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(s + head); uint4* d4 = reinterpret_cast<uint4*>(d + head);
            for (i64 i = threadIdx.x; i < body; i += blockDim.x) d4[i] = s4[i];
            for (i64 i = head + (body << 4) + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
        } else for (i64 i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
}

extern "C" int kolm_gather_payloads(kolm_ctx* c, const uint64_t* src_addr, const int64_t* len, int nblocks, uint8_t* out, int64_t* out_off,
                                    kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || nblocks < 0 || nblocks > c->max_blocks) return KOLM_E_ARG;
    if (!nblocks) { out_off[0] = 0; return KOLM_OK; }
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(s));                    // pinned staging reuse
    // staging in the pinned accumulator mirror: [src | len | off], each nblocks (+1) 8-byte words
    u64* hs = c->h_bacc; i64* hl = (i64*)(c->h_bacc + nblocks); i64* ho = (i64*)(c->h_bacc + 2 * (size_t)nblocks);
    i64 run = 0;
    for (int b = 0; b < nblocks; ++b) { hs[b] = src_addr[b]; hl[b] = len[b]; ho[b] = run; out_off[b] = run; run += len[b]; }
    out_off[nblocks] = run;
    u64* ds = c->d_bacc; i64* dl = (i64*)(c->d_bacc + nblocks); i64* dofs = (i64*)(c->d_bacc + 2 * (size_t)nblocks);
    CUDA_TRY(cudaMemcpyAsync(ds, hs, (size_t)nblocks * 24, cudaMemcpyHostToDevice, s));
    int grid = nblocks < 8 * c->sm_count ? nblocks : 8 * c->sm_count;
    KL(c, KC_MISC, run * 2, s, k_gather_payloads<<<grid, 256, 0, s>>>(ds, dl, dofs, out, nblocks));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}

// Per-block model selection (_encode_block kolm_final.py:821-864: ids in order, keep `plen < best`; the selection loops of
// kolm_final_researched_v2-2.py:2233-2252 / 2350-2369: strict '<'): the winner is the FIRST minimum of the candidates' payload
// sizes, i.e. the lowest id on ties.  One thread per block; sizes row-major [nblocks][ncand].
__global__ void __launch_bounds__(256) k_select_blocks(const i64* __restrict__ sizes, int nblocks, int ncand, i64* __restrict__ method, i64* __restrict__ best) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const i64* row = sizes + (size_t)b * ncand;
    i64 bs = row[0]; int bm = 0;
    for (int m = 1; m < ncand; ++m) { const i64 v = row[m]; if (v < bs) { bs = v; bm = m; } }
    method[b] = bm; best[b] = bs;
}

extern "C" int kolm_select_blocks(kolm_ctx* c, const int64_t* sizes, int nblocks, int ncand, int32_t* method, int64_t* best, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !sizes || !method || !best || nblocks < 0 || nblocks > c->max_blocks || ncand < 1 || ncand > 62) return KOLM_E_ARG;
    if (!nblocks) return KOLM_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(s));                    // pinned staging reuse
    const size_t nin = (size_t)nblocks * ncand;
    i64* hin = (i64*)c->h_bacc; i64* hout = hin + nin;     // [sizes | method | best] in the pinned accumulator mirror (64 words per block)
    memcpy(hin, sizes, nin * 8);
    i64* din = (i64*)c->d_bacc; i64* dout = din + nin;
    CUDA_TRY(cudaMemcpyAsync(din, hin, nin * 8, cudaMemcpyHostToDevice, s));
    KL(c, KC_PLAN, (i64)nin * 8, s, k_select_blocks<<<(nblocks + 255) / 256, 256, 0, s>>>(din, nblocks, ncand, dout, dout + nblocks));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(hout, dout, (size_t)nblocks * 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int b = 0; b < nblocks; ++b) { method[b] = (int32_t)hout[b]; best[b] = hout[nblocks + b]; }
    return KOLM_OK;
}

// general batched device-to-device copy: block b = len[b] bytes from src_addr[b] to dst_addr[b] (absolute device addresses, HOST arrays)
__global__ void __launch_bounds__(256) k_copy_blocks(const u64* __restrict__ src, const u64* __restrict__ dst, const i64* __restrict__ len, int nblocks) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const u8* s = reinterpret_cast<const u8*>(src[b]);
        u8* d = reinterpret_cast<u8*>(dst[b]);
        const i64 n = len[b];
        if ((((uintptr_t)s ^ (uintptr_t)d) & 15) == 0) {
            i64 head = (16 - ((uintptr_t)s & 15)) & 15; if (head > n) head = n;
            for (i64 i = threadIdx.x; i < head; i += blockDim.x) d[i] = s[i];
            const i64 body = (n - head) >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(s + head); uint4* d4 = reinterpret_cast<uint4*>(d + head);
            for (i64 i = threadIdx.x; i < body; i += blockDim.x) d4[i] = s4[i];
            for (i64 i = head + (body << 4) + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
        } else for (i64 i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
}

extern "C" int kolm_copy_blocks(kolm_ctx* c, const uint64_t* src_addr, const uint64_t* dst_addr, const int64_t* len, int nblocks, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || nblocks < 0 || nblocks > c->max_blocks) return KOLM_E_ARG;
    if (!nblocks) return KOLM_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(s));
    u64* hs = c->h_bacc; u64* hd = c->h_bacc + nblocks; i64* hl = (i64*)(c->h_bacc + 2 * (size_t)nblocks);
    i64 total = 0;
    for (int b = 0; b < nblocks; ++b) { hs[b] = src_addr[b]; hd[b] = dst_addr[b]; hl[b] = len[b]; total += len[b]; }
    CUDA_TRY(cudaMemcpyAsync(c->d_bacc, c->h_bacc, (size_t)nblocks * 24, cudaMemcpyHostToDevice, s));
    int grid = nblocks < 8 * c->sm_count ? nblocks : 8 * c->sm_count;
    KL(c, KC_MISC, total * 2, s, k_copy_blocks<<<grid, 256, 0, s>>>(c->d_bacc, c->d_bacc + nblocks, (const i64*)(c->d_bacc + 2 * (size_t)nblocks), nblocks));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}

#include "fused.cu"
