// bbwt_inv.cu — inverse bijective BWT.
//
// Replaces bbwt_inverse (kolm_final.py:327-369 == kolm_final_researched_v2-2.py:425-454):
//   pi  = stable argsort of positions by (L[i], i)           -> one 8-bit radix pass (same engine as the forward sort)
//   cycles of pi; a cycle with minimum index i0 and length d emits L[pi^k(i0)], k = 1..d;
//   cycles are ordered by i0 ascending and concatenated in REVERSE.
//
// Cycle structure by pointer jumping on (next, window minimum, steps to that minimum): after
// ceil(log2 n) doublings every element knows its cycle minimum i0 and its distance to it, hence its
// output slot  off(i0) + (d - dist - 1)  where off(i0) = bytes of all cycles with a larger minimum
// (a look-back prefix sum over positions of the cycle lengths stored at the minima).
#include "common.cuh"

struct __align__(16) JumpNode { u32 nxt, mn, dmn, pad; };

__global__ void __launch_bounds__(KOLM_THREADS) k_inv_keys(const u8* __restrict__ in, const TileDesc* __restrict__ tiles,
                                                           const BlockInfo* __restrict__ binfo, u32* __restrict__ K, u32* __restrict__ V) {
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff + (td.start - bi.pbase);
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) { K[td.start + x] = src[x]; V[td.start + x] = td.start + x; }
}

// node[r] for sorted slot r: pi[r] = V[r]  (positions are padded-global; mn/dmn initialised for a window of 1)
__global__ void __launch_bounds__(KOLM_THREADS) k_inv_init(const u32* __restrict__ V, const TileDesc* __restrict__ tiles, JumpNode* __restrict__ a) {
    TileDesc td = tiles[blockIdx.x];
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 i = td.start + x;
        JumpNode n; n.nxt = V[i]; n.mn = i; n.dmn = 0; n.pad = 0;
        a[i] = n;
    }
}

__global__ void __launch_bounds__(KOLM_THREADS) k_inv_jump(const JumpNode* __restrict__ a, JumpNode* __restrict__ b, const TileDesc* __restrict__ tiles, u32 span) {
    TileDesc td = tiles[blockIdx.x];
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 i = td.start + x;
        JumpNode me = a[i];
        JumpNode nx = a[me.nxt];
        JumpNode o;
        o.nxt = nx.nxt;
        if (me.mn <= nx.mn) { o.mn = me.mn; o.dmn = me.dmn; } else { o.mn = nx.mn; o.dmn = span + nx.dmn; }
        o.pad = 0;
        b[i] = o;
    }
}

// cycle length at each cycle minimum (0 elsewhere) -> prefix sum over positions -> off(i0) = len - inclusive_prefix(i0)
__global__ void __launch_bounds__(KOLM_THREADS) k_inv_offsets(const JumpNode* __restrict__ a, const u32* __restrict__ pi, const TileDesc* __restrict__ tiles,
                                                              const BlockInfo* __restrict__ binfo, u64* lb, u32* __restrict__ cyc_len, u32* __restrict__ cyc_off) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_last[KOLM_THREADS / 32];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    u32 d[KOLM_IPT];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < KOLM_IPT; ++k) {
        u32 r = tid * KOLM_IPT + k;
        d[k] = 0;
        if (r < td.count) {
            u32 i = td.start + r;
            if (a[i].mn == i) d[k] = a[pi[i]].dmn + 1;      // i is its cycle's minimum: length = steps from pi(i) back to i, plus one
            sum += d[k];
        }
    }
    u64 tot;
    u64 incl = block_scan_incl(sum, 0ull, OpAdd(), s_warp, &tot);
    u64 prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((tid & 31) == 31) s_last[tid >> 5] = incl;
    __syncthreads();
    if ((tid & 31) == 0) prev = (tid >> 5) ? s_last[(tid >> 5) - 1] : 0ull;
    if (tid < 32) {
        u64 e = lb_exclusive(lb, tile, (td.flags & 1u) != 0, tot, 0ull, OpAdd());
        if (tid == 0) s_excl = e;
    }
    __syncthreads();
    u64 run = s_excl + prev;
#pragma unroll
    for (int k = 0; k < KOLM_IPT; ++k) {
        u32 r = tid * KOLM_IPT + k;
        if (r < td.count) {
            run += d[k];
            if (d[k]) { cyc_len[td.start + r] = d[k]; cyc_off[td.start + r] = bi.len - (u32)run; }
        }
    }
}

__global__ void __launch_bounds__(KOLM_THREADS) k_inv_emit(const u8* __restrict__ in, u8* __restrict__ out, const JumpNode* __restrict__ a,
                                                           const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                           const u32* __restrict__ cyc_len, const u32* __restrict__ cyc_off) {
    TileDesc td = tiles[blockIdx.x];
    BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff + (td.start - bi.pbase);
    for (u32 x = threadIdx.x; x < td.count; x += KOLM_THREADS) {
        u32 i = td.start + x;
        JumpNode n = a[i];
        u32 d = cyc_len[n.mn], o = cyc_off[n.mn];
        u32 kidx = (n.mn == i) ? d - 1 : d - n.dmn - 1;
        out[bi.ioff + o + kidx] = src[x];
    }
}

int kolm_bbwt_inv_impl(kolm_ctx* c, const u8* in, u8* out, cudaStream_t s) {
    const int nt = c->ntiles;
    if (!nt) return KOLM_OK;
    const i64 N = c->total_bytes;
    KL(c, KC_INV, N * 9, s, k_inv_keys<<<nt, KOLM_THREADS, 0, s>>>(in, c->d_tiles, c->d_binfo, c->d_k0, c->d_v0));
    u32 *K, *V;
    KOLM_TRY(radix_sort(c, c->d_tiles, nt, N, c->d_btile0, c->d_btilen, 8, c->d_k0, c->d_v0, c->d_k1, c->d_v1, &K, &V, s));
    // V (= d_v1 after one pass) is pi.  Jump nodes ping-pong in (d_sa,d_rank,d_nr,d_fstart) viewed as two 16-byte arrays.
    // The four u32 arrays are separate allocations, so use two dedicated views: A over d_k0/d_v0 is unsafe (K/V live there);
    // node buffers are carved from d_sa+d_rank (A) and d_nr+d_fstart (B) only if contiguous -- they are not, so allocate lazily.
    if (!c->d_jump) {
        CUDA_TRY(cudaMalloc((void**)&c->d_jump, 2 * c->max_elems * sizeof(JumpNode)));
    }
    JumpNode* A = (JumpNode*)c->d_jump;
    JumpNode* B = A + c->max_elems;
    KL(c, KC_INV, N * 20, s, k_inv_init<<<nt, KOLM_THREADS, 0, s>>>(V, c->d_tiles, A));
    u32 span = 1;
    while (span < c->max_len) {
        KL(c, KC_INV, N * 48, s, k_inv_jump<<<nt, KOLM_THREADS, 0, s>>>(A, B, c->d_tiles, span));
        JumpNode* t = A; A = B; B = t;
        span <<= 1;
    }
    int lgrid = nt;
    KOLM_TRY(kolm_lb_reset_mode(c, false, nt, &lgrid, 1, s));
    KL(c, KC_INV, N * 24, s, k_inv_offsets<<<lgrid, KOLM_THREADS, 0, s>>>(A, V, c->d_tiles, c->d_binfo, c->d_lb, c->d_sa, c->d_rank));
    KL(c, KC_INV, N * 26, s, k_inv_emit<<<nt, KOLM_THREADS, 0, s>>>(in, out, A, c->d_tiles, c->d_binfo, c->d_sa, c->d_rank));
    CUDA_TRY(cudaGetLastError());
    return KOLM_OK;
}
