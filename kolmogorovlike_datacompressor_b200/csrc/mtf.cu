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

// pass A (encode): tlast[tile][c] = 1 + block-local position of the last c in the tile, 0 if absent
__global__ void __launch_bounds__(KOLM_THREADS) k_mtf_last(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u32* __restrict__ tlast) {
    __shared__ u32 last[256];
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    last[threadIdx.x] = 0;
    __syncthreads();
    u32 t0 = td.start - bi.pbase;
    const u8* src = in + bi.ioff + t0;
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u8 s = src[x];
        if (x + 1 == td.count || src[x + 1] != s) atomicMax(&last[s], t0 + x + 1);
    }
    __syncthreads();
    tlast[(size_t)blockIdx.x * 256 + threadIdx.x] = last[threadIdx.x];
}

// pass B (encode): exclusive running max over the tiles of each block
__global__ void __launch_bounds__(256) k_mtf_scan_max(u32* __restrict__ tlast, const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        u32 nt = tilen[b];
        u32* base = tlast + (size_t)tile0[b] * 256 + threadIdx.x;
        u32 run = 0;
        for (u32 t = 0; t < nt; ++t) { u32 v = base[(size_t)t * 256]; base[(size_t)t * 256] = run; run = max(run, v); }
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
__global__ void __launch_bounds__(256) k_mtf_dec_compose(u8* __restrict__ tperm, const u32* __restrict__ tile0, const u32* __restrict__ tilen, int nblocks) {
    __shared__ u8 cur[256];
    __shared__ u8 nxt[256];
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        u32 nt = tilen[b];
        u8* base = tperm + (size_t)tile0[b] * 256;
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

int kolm_mtf_impl(kolm_ctx* c, const u8* in, u8* out, bool decode, cudaStream_t s) {
    const int nt = c->ntiles, nb = c->nblocks;
    if (!nt) return KOLM_OK;
    int sgrid = nb < 4 * c->sm_count ? nb : 4 * c->sm_count;
    int wgrid = (nt + MTF_WARPS - 1) / MTF_WARPS;
    if (!decode) {
        const i64 N = c->total_bytes;
        KL(c, KC_MTF_PRE, N + (i64)nt * 1024, s, k_mtf_last<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_thist));
        KL(c, KC_MTF_SCAN, (i64)nt * 2048, s, k_mtf_scan_max<<<sgrid, 256, 0, s>>>(c->d_thist, c->d_btile0, c->d_btilen, nb));
        KL(c, KC_MTF_MAIN, 2 * N + (i64)nt * 1024, s, k_mtf_enc<<<wgrid, MTF_WARPS * 32, 0, s>>>(in, out, c->d_tiles, c->d_binfo, c->d_thist, nt));
    } else {
        u8* tperm = (u8*)c->d_thist;
        const i64 N = c->total_bytes;
        KL(c, KC_MTF_PRE, N + (i64)nt * 256, s, k_mtf_dec_perm<<<wgrid, MTF_WARPS * 32, 0, s>>>(in, c->d_tiles, c->d_binfo, tperm, nt));
        KL(c, KC_MTF_SCAN, (i64)nt * 512, s, k_mtf_dec_compose<<<sgrid, 256, 0, s>>>(tperm, c->d_btile0, c->d_btilen, nb));
        KL(c, KC_MTF_MAIN, 2 * N + (i64)nt * 256, s, k_mtf_dec<<<wgrid, MTF_WARPS * 32, 0, s>>>(in, out, c->d_tiles, c->d_binfo, tperm, nt));
    }
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}
