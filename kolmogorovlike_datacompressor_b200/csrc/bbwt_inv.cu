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


// ------------------------------------------------------------------------------------------------
// Work-efficient cycle ranking with rulers.  Every 32nd position of a block is a ruler; a ruler walks pi until it meets the
// next ruler (its segment: length, minimum index, offset of that minimum), the (next, window-min, steps-to-min, length)
// doubling then runs on the rulers only (32x fewer nodes), and each ruler finally re-walks its segment to place its
// elements.  Cycles that contain no ruler (short ones, or adversarial inputs) are detected by the covered-length check and
// send the whole batch down the element-level pointer-jumping path below.
// ------------------------------------------------------------------------------------------------
#define RULER_SHIFT 5
#define RULER_MAXWALK (1u << 22)
struct __align__(16) RulerNode { u32 nxt, mn, dmn, wlen; };     // nxt = ruler slot (padded position >> 5)

__global__ void __launch_bounds__(KOLM_TILE >> RULER_SHIFT) k_inv_walk(const u32* __restrict__ pi, const BlockInfo* __restrict__ binfo, const TileDesc* __restrict__ tiles,
                                                  RulerNode* __restrict__ seg, RulerNode* __restrict__ node, u32* __restrict__ visited,
                                                  u32* __restrict__ fallback) {
    // one tile = 4096 positions = 128 rulers = 128 threads (with 256-thread CTAs half of the resident threads idled, and these
    // pointer-chasing kernels are bound by the number of dependent loads in flight)
    const TileDesc td = tiles[blockIdx.x];
    const u32 k = threadIdx.x;
    const u32 r = td.start + (k << RULER_SHIFT);
    if (k < (KOLM_TILE >> RULER_SHIFT) && (k << RULER_SHIFT) < td.count) {
        u32 cur = pi[r], len = 1, mn = r, tm = 0;
        atomicOr(visited + (r >> 5), 1u << (r & 31));
        while (cur & ((1u << RULER_SHIFT) - 1)) {
            atomicOr(visited + (cur >> 5), 1u << (cur & 31));
            if (cur < mn) { mn = cur; tm = len; }
            cur = pi[cur];
            if (++len > RULER_MAXWALK) { atomicExch(fallback, 1u); break; }
        }
        RulerNode n; n.nxt = cur >> RULER_SHIFT; n.mn = mn; n.dmn = tm; n.wlen = len;
        seg[r >> RULER_SHIFT] = n; node[r >> RULER_SHIFT] = n;
    }
}

// cycles without a ruler: every unvisited element walks its cycle until it sees a smaller index (then it is not the
// minimum and stops); the minimum completes the loop and records the cycle length.
#define ORPHAN_MAXWALK (1u << 16)
__global__ void __launch_bounds__(KOLM_THREADS) k_inv_orphan_len(const u32* __restrict__ pi, const TileDesc* __restrict__ tiles,
                                                                 const u32* __restrict__ visited, u32* __restrict__ cyc_len, u32* __restrict__ fallback) {
    const TileDesc td = tiles[blockIdx.x];
    for (u32 x0 = threadIdx.x; x0 < td.count; x0 += KOLM_THREADS) {
        const u32 x = td.start + x0;
        if ((visited[x >> 5] >> (x & 31)) & 1u) continue;
        u32 cur = pi[x], len = 1;
        while (cur > x) { cur = pi[cur]; if (++len > ORPHAN_MAXWALK) { atomicExch(fallback, 1u); break; } }
        if (cur == x) cyc_len[x] = len;
    }
}

__global__ void __launch_bounds__(KOLM_THREADS) k_inv_orphan_emit(const u8* __restrict__ in, u8* __restrict__ out, const u32* __restrict__ pi,
                                                                  const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo,
                                                                  const u32* __restrict__ visited, const u32* __restrict__ cyc_len,
                                                                  const u32* __restrict__ cyc_off) {
    const TileDesc td = tiles[blockIdx.x];
    const BlockInfo bi = binfo[td.block];
    const u8* src = in + bi.ioff;
    for (u32 x0 = threadIdx.x; x0 < td.count; x0 += KOLM_THREADS) {
        const u32 x = td.start + x0;
        if ((visited[x >> 5] >> (x & 31)) & 1u) continue;
        const u32 d = cyc_len[x];
        if (!d) continue;                                    // an orphan that is not its cycle's minimum
        u8* dst = out + bi.ioff + cyc_off[x];
        u32 cur = x;
        for (u32 k = 0; k < d; ++k) { cur = pi[cur]; dst[k] = src[cur - bi.pbase]; }
    }
}

__global__ void __launch_bounds__(KOLM_TILE >> RULER_SHIFT) k_inv_rjump(const RulerNode* __restrict__ a, RulerNode* __restrict__ b, const TileDesc* __restrict__ tiles) {
    const TileDesc td = tiles[blockIdx.x];
    const u32 k = threadIdx.x;
    if (k >= (KOLM_TILE >> RULER_SHIFT) || (k << RULER_SHIFT) >= td.count) return;
    const u32 slot = (td.start >> RULER_SHIFT) + k;
    RulerNode me = a[slot], nx = a[me.nxt], o;
    o.nxt = nx.nxt;
    if (me.mn <= nx.mn) { o.mn = me.mn; o.dmn = me.dmn; } else { o.mn = nx.mn; o.dmn = me.wlen + nx.dmn; }
    u32 w = me.wlen + nx.wlen; o.wlen = w < me.wlen ? 0xffffffffu : w;
    b[slot] = o;
}

// cycle length at the cycle minimum M: the ruler whose segment contains M closes the loop
__global__ void __launch_bounds__(KOLM_TILE >> RULER_SHIFT) k_inv_rlen(const RulerNode* __restrict__ seg, const RulerNode* __restrict__ node, const TileDesc* __restrict__ tiles,
                                                  u32* __restrict__ cyc_len) {
    const TileDesc td = tiles[blockIdx.x];
    const u32 k = threadIdx.x;
    if (k >= (KOLM_TILE >> RULER_SHIFT) || (k << RULER_SHIFT) >= td.count) return;
    const u32 slot = (td.start >> RULER_SHIFT) + k;
    RulerNode sg = seg[slot], me = node[slot];
    if (me.dmn < sg.wlen && sg.mn == me.mn) cyc_len[me.mn] = (sg.wlen - me.dmn) + node[sg.nxt].dmn;
}

__global__ void __launch_bounds__(KOLM_THREADS) k_inv_roffsets(const TileDesc* __restrict__ tiles, const BlockInfo* __restrict__ binfo, u64* lb,
                                                               const u32* __restrict__ cyc_len, u32* __restrict__ cyc_off) {
    __shared__ u64 s_warp[KOLM_THREADS / 32];
    __shared__ u64 s_last[KOLM_THREADS / 32];
    __shared__ u64 s_excl;
    const u32 tid = threadIdx.x;
    const u32 tile = lb_take_ticket(lb);
    if (tile == LB_NO_TILE) return;
    const TileDesc td = tiles[tile];
    const BlockInfo bi = binfo[td.block];
    u32 d[KOLM_IPT]; u64 sum = 0;
#pragma unroll
    for (int k = 0; k < KOLM_IPT; ++k) { u32 r = tid * KOLM_IPT + k; d[k] = r < td.count ? cyc_len[td.start + r] : 0; sum += d[k]; }
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
    for (int k = 0; k < KOLM_IPT; ++k) { u32 r = tid * KOLM_IPT + k; if (r < td.count) { run += d[k]; if (d[k]) cyc_off[td.start + r] = bi.len - (u32)run; } }
}

__global__ void __launch_bounds__(KOLM_TILE >> RULER_SHIFT) k_inv_remit(const u8* __restrict__ in, u8* __restrict__ out, const u32* __restrict__ pi,
                                                   const RulerNode* __restrict__ node, const TileDesc* __restrict__ tiles,
                                                   const BlockInfo* __restrict__ binfo, const u32* __restrict__ cyc_len, const u32* __restrict__ cyc_off) {
    const TileDesc td = tiles[blockIdx.x];
    const u32 k = threadIdx.x;
    if (k >= (KOLM_TILE >> RULER_SHIFT) || (k << RULER_SHIFT) >= td.count) return;
    const BlockInfo bi = binfo[td.block];
    const u32 r = td.start + (k << RULER_SHIFT);
    const RulerNode me = node[r >> RULER_SHIFT];
    const u32 M = me.mn, d = cyc_len[M], o = cyc_off[M];
    const u8* src = in + bi.ioff;
    u8* dst = out + bi.ioff + o;
    u32 dist = me.dmn;                                       // steps from the current element to M (going forward)
    u32 cur = r;
    do {
        dst[d - dist - 1] = src[cur - bi.pbase];             // element at distance `dist` before M is the (d-dist)-th emitted symbol
        cur = pi[cur];
        dist = dist ? dist - 1 : d - 1;
    } while (cur & ((1u << RULER_SHIFT) - 1));
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
    static int rulers = -1;
    if (rulers < 0) { const char* e = getenv("KOLM_INV_RULERS"); rulers = e ? atoi(e) : 1; }
    if (rulers) {
        // ruler path: three 16-byte node arrays of max_elems/32 entries carved from the jump scratch
        RulerNode* seg = (RulerNode*)c->d_jump;
        RulerNode* RA = seg + (c->max_elems >> RULER_SHIFT) + 1;
        RulerNode* RB = RA + (c->max_elems >> RULER_SHIFT) + 1;
        u32* visited = c->d_single;
        CUDA_TRY(cudaMemsetAsync(visited, 0, ((size_t)c->total_elems / 32 + 2) * 4, s));
        CUDA_TRY(cudaMemsetAsync(c->d_stats + 8, 0, 4, s));
        CUDA_TRY(cudaMemsetAsync(c->d_sa, 0, (size_t)c->total_elems * 4, s));                 // cyc_len: non-zero only at cycle minima
        KL(c, KC_INV, N * 5, s, k_inv_walk<<<nt, KOLM_TILE >> RULER_SHIFT, 0, s>>>(V, c->d_binfo, c->d_tiles, seg, RA, visited, c->d_stats + 8));
        KL(c, KC_INV, N / 8, s, k_inv_orphan_len<<<nt, KOLM_THREADS, 0, s>>>(V, c->d_tiles, visited, c->d_sa, c->d_stats + 8));
        CUDA_TRY(cudaMemcpyAsync(c->h_stats + 8, c->d_stats + 8, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (c->h_stats[8] == 0) {                            // no walk hit its cap (else: element-level pointer jumping below)
            u32 nr = (c->max_len >> RULER_SHIFT) + 2;
            for (u32 span = 1; span < nr; span <<= 1) {
                KL(c, KC_INV, (N >> RULER_SHIFT) * 48, s, k_inv_rjump<<<nt, KOLM_TILE >> RULER_SHIFT, 0, s>>>(RA, RB, c->d_tiles));
                RulerNode* t = RA; RA = RB; RB = t;
            }
            KL(c, KC_INV, (N >> RULER_SHIFT) * 40, s, k_inv_rlen<<<nt, KOLM_TILE >> RULER_SHIFT, 0, s>>>(seg, RA, c->d_tiles, c->d_sa));
            int lgrid = nt;
            KOLM_TRY(kolm_lb_reset_mode(c, false, nt, &lgrid, 1, s));
            KL(c, KC_INV, N * 8, s, k_inv_roffsets<<<lgrid, KOLM_THREADS, 0, s>>>(c->d_tiles, c->d_binfo, c->d_lb, c->d_sa, c->d_rank));
            KL(c, KC_INV, N * 6, s, k_inv_remit<<<nt, KOLM_TILE >> RULER_SHIFT, 0, s>>>(in, out, V, RA, c->d_tiles, c->d_binfo, c->d_sa, c->d_rank));
            KL(c, KC_INV, N / 8, s, k_inv_orphan_emit<<<nt, KOLM_THREADS, 0, s>>>(in, out, V, c->d_tiles, c->d_binfo, visited, c->d_sa, c->d_rank));
            CUDA_TRY(cudaGetLastError());
            return KOLM_OK;
        }
    }
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
