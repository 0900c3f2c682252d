// fused.cu — kolm_encode_blocks / kolm_decode_blocks: what compress() / decompress() call per batch.
//
//   KOLM  _encode_block   kolm_final.py:821-864 (ids 0..3 in order, keep `plen < best`)   / decompress loop :925-949
//   KOLR  selection loops kolm_final_researched_v2-2.py:2233-2252, 2350-2369 (strict '<') / decompress loop :2530-2540
//
// Encode: every candidate of the profile is evaluated on the device and its exact per-block size lands in ONE device table
// cs[candidate][block]; k_fe_select takes the first minimum per block (= the lowest id on ties) and scans the winning sizes into the
// blocks' final offsets; then only the WINNERS are emitted, straight into their final place in payload_out: the Rice packers, the
// residual coder and the raw copy run with a per-block filter (method[b] == id), LZ77 and Re-Pair — whose state does not survive the
// BBWT sort — were materialised in scratch and are copied.  No size, offset or method id visits the host before the single copy at
// the end (the BBWT sort keeps its own per-round loop control, and the incremental Re-Pair its slab-pool set-up).
// Decode: blocks are grouped by method; each group's payloads are compacted, decoded as one batch and copied to their final offsets.
#include "common.cuh"

#define FE_MAXC 12
#define FE_BIG 0x3fffffffffffffffll                          // size of a candidate that is not evaluated

__global__ void k_fe_fill(i64* __restrict__ row, int nb, i64 v) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) row[b] = v;
}
__global__ void k_fe_raw(const BlockInfo* __restrict__ binfo, i64* __restrict__ row, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) row[b] = binfo[b].len;
}
__global__ void k_fe_res(const u64* __restrict__ bacc, const BlockInfo* __restrict__ binfo, i64* __restrict__ row, int nb, int kind) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) row[b] = (i64)binfo[b].len + (i64)bacc[(size_t)b * 64 + 40 + kind];
}
__global__ void k_fe_diff(const i64* __restrict__ poff, i64* __restrict__ row, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) row[b] = poff[b + 1] - poff[b];
}
// Re-Pair row: blocks whose candidate stopped early (accumulator slot 34, see k_repair_enc) count as not evaluated
__global__ void k_fe_diff_rp(const i64* __restrict__ poff, const u64* __restrict__ acc, i64* __restrict__ row, int nb, u32* __restrict__ nstopped) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const bool st = acc[(size_t)b * 64 + 34] != 0;
    row[b] = st ? FE_BIG : poff[b + 1] - poff[b];
    if (st) atomicAdd(nstopped, 1u);
}
// limit[b] = the smallest size among the candidates other than `skip` (what Re-Pair has to beat)
__global__ void k_fe_limit(const i64* __restrict__ cs, int nb, int ncand, int skip, i64* __restrict__ limit) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    i64 v = FE_BIG;
    for (int m = 0; m < ncand; ++m) if (m != skip) { const i64 x = cs[(size_t)m * nb + b]; if (x < v) v = x; }
    limit[b] = v;
}
__global__ void k_fe_kf(const u64* __restrict__ bacc, i64* __restrict__ row, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) row[b] = (i64)bacc[(size_t)b * RB_STRIDE + RB_BYTES];
}
__global__ void k_fe_k2(const i64* __restrict__ sizes5, i64* __restrict__ cs, int nb, int row0, u32 mask) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) for (int k = 0; k < 5; ++k) if ((mask >> (row0 + k)) & 1u) cs[(size_t)(row0 + k) * nb + b] = sizes5[5 * (size_t)b + k];
}
// single CTA: winner per block = first minimum over the candidate rows in id order (strict '<'), then the exclusive scan of the
// winning sizes = the blocks' final byte offsets in the payload area (foff[nb] = its length)
__global__ void __launch_bounds__(1024) k_fe_select(const i64* __restrict__ cs, int nb, int ncand, int* __restrict__ method, i64* __restrict__ foff) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += blockDim.x) {
        const int b = base + threadIdx.x;
        u64 bytes = 0;
        if (b < nb) {
            i64 bs = cs[b]; int bm = 0;
            for (int m = 1; m < ncand; ++m) { const i64 v = cs[(size_t)m * nb + b]; if (v < bs) { bs = v; bm = m; } }
            method[b] = bm; bytes = (u64)bs;
        }
        u64 v = bytes;
        for (int o = 1; o < 32; o <<= 1) { const u64 n = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (u32)o) v += n; }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = v;
        __syncthreads();
        u64 pre = 0;
        for (u32 i = 0; i < (threadIdx.x >> 5); ++i) pre += s_w[i];
        const u64 carry = s_carry;
        if (b < nb) foff[b] = (i64)(carry + pre + v - bytes);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) foff[nb] = (i64)s_carry;
}
__global__ void k_fe_setoff(u64* __restrict__ bacc, const i64* __restrict__ foff, int nb, int slot) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) bacc[(size_t)b * 64 + slot] = (u64)foff[b];
}
// winners of candidate `want`: their materialised payload (src = base + soff[b], or an absolute address) -> out + foff[b]
__global__ void __launch_bounds__(256) k_fe_copy(const u8* __restrict__ base, const i64* __restrict__ soff, const u64* __restrict__ saddr,
                                                 const BlockInfo* __restrict__ binfo /* raw: src = base + ioff */, const i64* __restrict__ foff,
                                                 const int* __restrict__ method, int want, u8* __restrict__ out, int nb, u64 cap) {
    if ((u64)foff[nb] > cap) return;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        if (method[b] != want) continue;
        const u8* s = saddr ? reinterpret_cast<const u8*>(saddr[b]) : base + (binfo ? binfo[b].ioff : soff[b]);
        u8* d = out + foff[b];
        const i64 n = foff[b + 1] - foff[b];
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

static inline size_t fe_align(size_t x) { return (x + 255) & ~(size_t)255; }

struct FeLayout {
    size_t cs, foff, method, lzoff, rpoff, extaddr, L, M, lzpay, rppay, rpacc, rplim, misc, total, lzcap, rpcap;
};
static FeLayout fe_layout(int profile, size_t n, int nb, bool rp_inline) {
    FeLayout l; size_t p = 0;
    auto take = [&](size_t bytes) { size_t at = p; p = fe_align(p + bytes); return at; };
    l.cs = take((size_t)FE_MAXC * nb * 8); l.foff = take((size_t)(nb + 1) * 8); l.method = take((size_t)nb * 4);
    l.lzoff = take((size_t)(nb + 1) * 8); l.rpoff = take((size_t)(nb + 1) * 8); l.extaddr = take((size_t)nb * 8);
    l.L = take(n + 64); l.M = take(n + 64);
    l.lzcap = 2 * n + 16 * (size_t)nb + 64; l.lzpay = take(l.lzcap);
    l.rpcap = (profile == KOLM_PROFILE_KOLR && rp_inline) ? 5 * n + 16 * (size_t)nb + 64 : 0; l.rppay = take(l.rpcap);
    const bool rp = l.rpcap != 0;                            // Re-Pair runs last, beside the Rice state: its own accumulators and limits
    l.rpacc = take(rp ? (size_t)nb * 64 * 8 : 0); l.rplim = take(rp ? (size_t)nb * 8 : 0); l.misc = take(64);
    l.total = p;
    return l;
}

extern "C" size_t kolm_encode_blocks_scratch(int profile, size_t batch_bytes, int nblocks) {
    return fe_layout(profile, batch_bytes, nblocks < 1 ? 1 : nblocks, true).total;
}

extern "C" int kolm_encode_blocks(kolm_ctx* c, int profile, const uint8_t* in, const int64_t* off, int nblocks, uint32_t cand_mask,
                                  int ext_id, const int64_t* ext_sizes, const uint64_t* ext_addr, uint8_t* scratch, size_t scratch_bytes,
                                  uint8_t* payload_out, size_t cap, int64_t* payload_off, uint8_t* method_ids, int64_t* sizes_out,
                                  kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !off || !payload_off || !method_ids || (profile != KOLM_PROFILE_KOLM && profile != KOLM_PROFILE_KOLR)) return KOLM_E_ARG;
    if (((uintptr_t)payload_out & 3) != 0 || ((uintptr_t)scratch & 255) != 0) return KOLM_E_ARG;
    KOLM_TRY(kolm_set_batch(c, off, nblocks, s));
    const int nb = c->nblocks, nt = c->ntiles;
    if (!nb) { payload_off[0] = 0; return KOLM_OK; }
    if (off[0] != 0) return KOLM_E_ARG;                      // the staging buffers for the BBWT / MTF bytes are indexed like `in`
    const bool kolm = profile == KOLM_PROFILE_KOLM;
    const int ncand = kolm ? 4 : 10;
    const u32 all = (1u << ncand) - 1u;
    u32 mask = cand_mask ? (cand_mask & all) : all;
    mask |= 1u;                                              // raw is what the reference falls back to when nothing else is offered
    const int id_lz = kolm ? 3 : 7, id_rp = kolm ? -1 : 9;
    const bool ext = ext_id >= 0;
    if (ext && (ext_id >= ncand || !ext_sizes || !ext_addr)) return KOLM_E_ARG;
    if (ext) mask |= 1u << ext_id;
    const bool rp_inline = id_rp >= 0 && ((mask >> id_rp) & 1u) && !(ext && ext_id == id_rp);
    const size_t n = (size_t)c->total_bytes;
    const FeLayout L = fe_layout(profile, n, nb, rp_inline);
    if (L.total > scratch_bytes) return KOLM_E_CAPACITY;
    i64* cs = (i64*)(scratch + L.cs); i64* foff = (i64*)(scratch + L.foff); int* method = (int*)(scratch + L.method);
    i64* lzoff = (i64*)(scratch + L.lzoff); i64* rpoff = (i64*)(scratch + L.rpoff); u64* extaddr = (u64*)(scratch + L.extaddr);
    u8* dL = scratch + L.L; u8* dM = scratch + L.M; u8* lzpay = scratch + L.lzpay; u8* rppay = scratch + L.rppay;
    const int g1 = (nb + 255) / 256;
    auto row = [&](int id) { return cs + (size_t)id * nb; };
    for (int id = 0; id < ncand; ++id) if (!((mask >> id) & 1u)) k_fe_fill<<<g1, 256, 0, s>>>(row(id), nb, FE_BIG);
    k_fe_raw<<<g1, 256, 0, s>>>(c->d_binfo, row(0), nb);
    // ---- byte-predictor residual coders: sizes only (KOLM: id 1 = XOR; KOLR: id 1 = delta, id 8 = LFSR predictor)
    const bool any_res = kolm ? ((mask >> 1) & 1u) : (((mask >> 1) | (mask >> 8)) & 1u);
    if (any_res) {
        KOLM_TRY(lfsr_init(c));
        CUDA_TRY(cudaMemsetAsync(c->d_bacc, 0, (size_t)nb * 64 * 8, s));
        if (nt) KL(c, KC_MISC, c->total_bytes, s, k_res_cost<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_bacc));
        if (kolm) k_fe_res<<<g1, 256, 0, s>>>(c->d_bacc, c->d_binfo, row(1), nb, 0);
        else {
            if ((mask >> 1) & 1u) k_fe_res<<<g1, 256, 0, s>>>(c->d_bacc, c->d_binfo, row(1), nb, 1);
            if ((mask >> 8) & 1u) k_fe_res<<<g1, 256, 0, s>>>(c->d_bacc, c->d_binfo, row(8), nb, 2);
        }
    }
    // ---- external candidate (e.g. Re-Pair of long blocks run ahead on another context): sizes and payload addresses from the caller
    if (ext) {
        CUDA_TRY(cudaMemcpyAsync(row(ext_id), ext_sizes, (size_t)nb * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(extaddr, ext_addr, (size_t)nb * 8, cudaMemcpyHostToDevice, s));
    }
    // ---- LZ77 and Re-Pair: their working state lives in the sort buffers, so their payloads are materialised before the BBWT sort
    if ((mask >> id_lz) & 1u) {
        KOLM_TRY(kolm_lz77_enc_impl(c, in, kolm ? 255u : 4096u, kolm ? 127u : 0u, lzpay, L.lzcap, nullptr, s));
        CUDA_TRY(cudaMemcpyAsync(lzoff, c->d_poff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToDevice, s));
        k_fe_diff<<<g1, 256, 0, s>>>(lzoff, row(id_lz), nb);
    }
    // ---- BBWT -> MTF -> exact sizes of the Rice coders
    const u32 k2mask = kolm ? 0u : (mask & 0x7Cu);
    const bool kf = kolm && ((mask >> 2) & 1u);
    if (kf || k2mask) {
        int rp = 0, rc = 0;
        KOLM_TRY(kolm_bbwt_fwd_impl(c, in, dL, &rp, &rc, s));
        c->counters[0] = rp; c->counters[1] = rc;
        KOLM_TRY(kolm_mtf_impl(c, dL, dM, false, s));
        KOLM_TRY(rice2_costs(c, dM, kf, k2mask != 0, s));
        KL(c, KC_RICE_PLAN, (i64)nb * 256, s, k_rice_plan<<<1, 1024, 0, s>>>(c->d_bacc, c->d_binfo, c->d_poff, c->d_params, c->d_sizes, nb, kf ? 1 : 2, 0));
        if (kf) k_fe_kf<<<g1, 256, 0, s>>>(c->d_bacc, row(2), nb);
        else k_fe_k2<<<g1, 256, 0, s>>>(c->d_sizes, cs, nb, 2, k2mask);
    }
    // ---- Re-Pair last: every other size is in the table, so the shared-memory kernel can stop a block as soon as its lower
    //      bound reaches the best of them (k_repair_enc).  A caller that asks for the size table gets every candidate in full.
    //      The Rice state (d_bacc, d_poff) stays untouched: accumulators and offsets of this candidate live in the scratch.
    u32* nstopped = (u32*)(scratch + L.misc);
    CUDA_TRY(cudaMemsetAsync(nstopped, 0, 64, s));
    if (rp_inline) {
        static int stop_on = -1;
        if (stop_on < 0) { const char* e = getenv("KOLM_REPAIR_STOP"); stop_on = e ? atoi(e) : 1; }
        i64* rplim = nullptr;
        if (stop_on && !sizes_out) {
            rplim = (i64*)(scratch + L.rplim);
            k_fe_limit<<<g1, 256, 0, s>>>(cs, nb, ncand, id_rp, rplim);
        }
        u64* const bacc0 = c->d_bacc; i64* const poff0 = c->d_poff;
        c->d_bacc = (u64*)(scratch + L.rpacc); c->d_poff = rpoff;
        const int rc = kolm_repair_enc_impl(c, in, rppay, L.rpcap, nullptr, s, rplim);
        c->d_bacc = bacc0; c->d_poff = poff0;
        if (rc != KOLM_OK) return rc;
        k_fe_diff_rp<<<g1, 256, 0, s>>>(rpoff, (const u64*)(scratch + L.rpacc), row(id_rp), nb, nstopped);
    }
    // ---- selection and final offsets, on the device
    KL(c, KC_PLAN, (i64)nb * ncand * 8, s, k_fe_select<<<1, 1024, 0, s>>>(cs, nb, ncand, method, foff));
    // ---- winners into their final place
    KL(c, KC_ZERO, 0, s, k_zero_words<<<4 * c->sm_count, 256, 0, s>>>((u32*)payload_out, foff + nb, cap / 4));
    if ((kf || k2mask) && nt) {
        const int g = (nt + R2_WARPS - 1) / R2_WARPS;
        k_fe_setoff<<<g1, 256, 0, s>>>(c->d_bacc, foff, nb, RB_OFF);
        if (kf) {
            KL(c, KC_RICE_PLAN, (i64)nt * 64, s, k_rice_tile_offsets<<<nb, 256, 0, s>>>((u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 1, 0));
            KL(c, KC_RICE_PACK, c->total_bytes, s, k_rice2_kf_pack<<<g, R2_THREADS, 0, s>>>(dM, c->d_tiles, c->d_binfo, (const u64*)c->d_thist, c->d_bacc, (u32*)payload_out, nt,
                                                                                         foff + nb, (u64)cap, method, 2));
        } else {
            static const int FLAGS[5] = {0, 1, 4, 8, 16};
            for (int k = 0; k < 5; ++k) {
                if (!((k2mask >> (2 + k)) & 1u)) continue;
                KL(c, KC_RICE_PLAN, (i64)nt * 64, s, k_rice_tile_offsets<<<nb, 256, 0, s>>>((u64*)c->d_thist, c->d_btile0, c->d_btilen, c->d_bacc, 2, k));
                KL(c, KC_RICE_PACK, c->total_bytes, s, k_rice2_k2_pack<<<g, R2_THREADS, 0, s>>>(dM, c->d_tiles, c->d_binfo, (const u64*)c->d_thist, c->d_bacc, (u32*)payload_out, FLAGS[k],
                                                                                             nt, foff + nb, (u64)cap, method, 2 + k));
            }
        }
    }
    if (any_res && nt) {                                     // the residual coder recomputes its bytes: only the block offsets are needed
        k_fe_setoff<<<g1, 256, 0, s>>>(c->d_bacc, foff, nb, 33);
        const int ids[2] = {1, kolm ? -1 : 8}, kinds[2] = {kolm ? 0 : 1, 2};
        for (int q = 0; q < 2; ++q) {
            if (ids[q] < 0 || !((mask >> ids[q]) & 1u)) continue;
            int lgrid = nt;
            KOLM_TRY(kolm_lb_reset_mode(c, false, nt, &lgrid, 1, s));
            KL(c, KC_MISC, c->total_bytes * 2, s, k_res_emit<<<lgrid, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_lb, c->d_bacc, payload_out, kinds[q], foff + nb, (u64)cap,
                                                                                         method, ids[q]));
        }
    }
    const int cg = nb < 8 * c->sm_count ? nb : 8 * c->sm_count;
    KL(c, KC_MISC, c->total_bytes, s, k_fe_copy<<<cg, 256, 0, s>>>(in, nullptr, nullptr, c->d_binfo, foff, method, 0, payload_out, nb, (u64)cap));
    if ((mask >> id_lz) & 1u) KL(c, KC_MISC, 0, s, k_fe_copy<<<cg, 256, 0, s>>>(lzpay, lzoff, nullptr, nullptr, foff, method, id_lz, payload_out, nb, (u64)cap));
    if (rp_inline) KL(c, KC_MISC, 0, s, k_fe_copy<<<cg, 256, 0, s>>>(rppay, rpoff, nullptr, nullptr, foff, method, id_rp, payload_out, nb, (u64)cap));
    if (ext) KL(c, KC_MISC, 0, s, k_fe_copy<<<cg, 256, 0, s>>>(nullptr, nullptr, extaddr, nullptr, foff, method, ext_id, payload_out, nb, (u64)cap));
    CUDA_TRY(cudaGetLastError());
    // ---- one copy home: offsets, method ids and (optionally) the whole size table
    i64* h_foff = (i64*)c->h_bacc; int* h_m = (int*)(c->h_bacc + (size_t)nb + 1); i64* h_cs = (i64*)(c->h_bacc + (size_t)nb + 1 + ((size_t)nb + 1) / 2 + 1);
    CUDA_TRY(cudaMemcpyAsync(h_foff, foff, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h_m, method, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    if (sizes_out) CUDA_TRY(cudaMemcpyAsync(h_cs, cs, (size_t)ncand * nb * 8, cudaMemcpyDeviceToHost, s));
    u32* h_st = (u32*)(h_cs + (size_t)ncand * nb);
    CUDA_TRY(cudaMemcpyAsync(h_st, nstopped, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    c->counters[5] = (i64)*h_st;
    memcpy(payload_off, h_foff, (size_t)(nb + 1) * 8);
    for (int b = 0; b < nb; ++b) method_ids[b] = (uint8_t)h_m[b];
    if (sizes_out) for (int b = 0; b < nb; ++b) for (int m = 0; m < ncand; ++m) sizes_out[(size_t)b * ncand + m] = h_cs[(size_t)m * nb + b];
    if ((size_t)payload_off[nb] + 8 > cap) return KOLM_E_CAPACITY;
    return KOLM_OK;
}

extern "C" int kolm_encode_blocks_stats(kolm_ctx* c, int64_t* out2) {
    if (!c || !out2) return KOLM_E_ARG;
    out2[0] = c->counters[5]; out2[1] = 0;
    return KOLM_OK;
}

// ---------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------
extern "C" size_t kolm_decode_blocks_scratch(size_t payload_bytes, size_t out_bytes, int nblocks) {
    (void)nblocks;
    return fe_align(payload_bytes + 64) + 2 * fe_align(out_bytes + 64) + 1024;   // compacted payloads, decoded group, inverse-MTF staging
}

extern "C" int kolm_decode_blocks(kolm_ctx* c, int profile, const uint8_t* payload, const int64_t* payload_start, const int64_t* payload_len,
                                  const uint8_t* method_ids, const int64_t* out_off, int nblocks, uint8_t* scratch, size_t scratch_bytes,
                                  uint8_t* out, int* bad_block, kolm_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!c || !payload_start || !payload_len || !method_ids || !out_off || nblocks < 0 || (profile != KOLM_PROFILE_KOLM && profile != KOLM_PROFILE_KOLR)) return KOLM_E_ARG;
    if (bad_block) *bad_block = -1;
    if (!nblocks) return KOLM_OK;
    if (nblocks > c->max_blocks) return KOLM_E_CAPACITY;
    const int nmeth = profile == KOLM_PROFILE_KOLM ? 4 : 11;
    std::vector<std::vector<int>> groups(nmeth);
    for (int b = 0; b < nblocks; ++b) {
        if (method_ids[b] >= nmeth) { if (bad_block) *bad_block = b; return KOLM_E_CORRUPT; }   // "Unknown method id" (KF.py:929, V22.py:2533)
        groups[method_ids[b]].push_back(b);
    }
    if (profile == KOLM_PROFILE_KOLR && !groups[10].empty()) return KOLM_E_UNSUPPORTED;      // v2_new needs a context of 8x the batch: kolm_v2new_dec
    int first_err = KOLM_OK, first_bad = nblocks;
    auto note = [&](int code, int b) { if (code != KOLM_OK && b < first_bad) { first_err = code; first_bad = b; } };
    std::vector<u64> src, dst; std::vector<i64> len, goff, loff;
    for (int m = 0; m < nmeth; ++m) {
        const std::vector<int>& g = groups[m];
        if (g.empty()) continue;
        const int ng = (int)g.size();
        src.resize(ng); dst.resize(ng); len.resize(ng); goff.assign(ng + 1, 0); loff.assign(ng + 1, 0);
        i64 ptot = 0, otot = 0;
        for (int i = 0; i < ng; ++i) {
            const int b = g[i];
            const i64 pl = payload_len[b], ol = out_off[b + 1] - out_off[b];
            if (pl < 0 || ol < 0 || payload_start[b] < 0) return KOLM_E_ARG;
            src[i] = (u64)(uintptr_t)(payload + payload_start[b]); len[i] = pl; ptot += pl;
            loff[i + 1] = loff[i] + ol; otot += ol;
        }
        if (m == 0) {                                        // raw: the payload is the block (KF.py:697-700; V22.py:2101-2103)
            for (int i = 0; i < ng; ++i) {
                const int b = g[i];
                if (len[i] != loff[i + 1] - loff[i]) note(KOLM_E_CORRUPT, b);
                dst[i] = (u64)(uintptr_t)(out + out_off[b]);
                if (len[i] != loff[i + 1] - loff[i]) len[i] = 0;
            }
            KOLM_TRY(kolm_copy_blocks(c, src.data(), dst.data(), len.data(), ng, stream));
            continue;
        }
        const size_t gp = fe_align((size_t)ptot + 64);
        if (gp + (size_t)otot + 64 > scratch_bytes) return KOLM_E_CAPACITY;
        u8* gpay = scratch; u8* gout = scratch + gp;
        KOLM_TRY(kolm_gather_payloads(c, src.data(), len.data(), ng, gpay, goff.data(), stream));
        KOLM_TRY(kolm_set_batch(c, loff.data(), ng, s));
        int rc;
        if (profile == KOLM_PROFILE_KOLM) {
            if (m == 1) rc = kolm_residual_dec_impl(c, gpay, goff.data(), 0, gout, s);
            else if (m == 2) rc = kolm_rice_kf_dec_impl(c, gpay, goff.data(), gout, s);
            else rc = kolm_lz77_dec_impl(c, gpay, goff.data(), 0u, gout, s);
        } else {
            static const int FLAGS[5] = {0, 1, 4, 8, 16};
            if (m == 1) rc = kolm_residual_dec_impl(c, gpay, goff.data(), 1, gout, s);
            else if (m >= 2 && m <= 6) rc = kolm_rice_k2_dec_impl(c, gpay, goff.data(), FLAGS[m - 2], gout, s);
            else if (m == 7) rc = kolm_lz77_dec_impl(c, gpay, goff.data(), 4096u, gout, s);
            else if (m == 8) rc = kolm_residual_dec_impl(c, gpay, goff.data(), 2, gout, s);
            else rc = kolm_repair_dec_impl(c, gpay, goff.data(), gout, s);
        }
        const bool bbwt = profile == KOLM_PROFILE_KOLM ? m == 2 : (m >= 2 && m <= 6);
        if (rc == KOLM_E_CUDA || rc == KOLM_E_ARG || rc == KOLM_E_CAPACITY || rc == KOLM_E_UNSUPPORTED) return rc;
        if (rc != KOLM_OK) { for (int i = 0; i < ng; ++i) if (c->h_err[i]) { note(c->h_err[i], g[i]); break; } continue; }
        if (bbwt) {                                          // Rice/gamma parse -> inverse MTF -> inverse BBWT; the MTF bytes go through the (spent) payload area when it is large enough
            u8* tmp = ((size_t)otot + 64 <= gp) ? gpay : nullptr;
            if (!tmp) { if (gp + 2 * ((size_t)otot + 64) > scratch_bytes) return KOLM_E_CAPACITY; tmp = gout + fe_align((size_t)otot + 64); }
            KOLM_TRY(kolm_mtf_impl(c, gout, tmp, true, s));
            KOLM_TRY(kolm_bbwt_inv_impl(c, tmp, gout, s));
        }
        for (int i = 0; i < ng; ++i) { src[i] = (u64)(uintptr_t)(gout + loff[i]); dst[i] = (u64)(uintptr_t)(out + out_off[g[i]]); len[i] = loff[i + 1] - loff[i]; }
        KOLM_TRY(kolm_copy_blocks(c, src.data(), dst.data(), len.data(), ng, stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    if (first_err != KOLM_OK) { if (bad_block) *bad_block = first_bad; return first_err; }
    return KOLM_OK;
}
