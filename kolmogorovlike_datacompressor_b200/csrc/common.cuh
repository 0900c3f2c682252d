// common.cuh — shared device/host infrastructure of libkolm_b200 (sm_100a only).
//
// Data model: a *batch* is nblocks independent blocks laid back to back in one
// device byte buffer, described by host offsets off[nblocks+1].  Every per-element scratch array
// lives in a *padded index space*: block b owns [pbase[b], pbase[b]+len[b]) with pbase a multiple
// of 32 elements, so that every tile start is 128-byte aligned (1-D TMA bulk copies need 16 B).
// Work is cut into tiles of <= TILE elements that never straddle a block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define KOLM_TILE 4096            // elements per tile
#define KOLM_THREADS 256          // threads per tile CTA
#define KOLM_IPT (KOLM_TILE / KOLM_THREADS)
#define KOLM_PAD 32               // block base alignment in the padded index space (elements)

#include "../../include/kolm_abi.h"   // KOLM_OK / KOLM_E_* codes

struct TileDesc {
    u32 start;   // padded global index of the first element
    u32 count;   // elements in this tile (1..TILE)
    u32 block;   // owning block
    u32 flags;   // bit0: first tile of its block, bit1: last tile of its block
};

struct BlockInfo {
    i64 ioff;    // byte offset of the block in the caller's buffers
    u32 pbase;   // base in the padded index space
    u32 len;     // block length in bytes / elements
};

#define CUDA_TRY(x)                                   \
    do {                                              \
        cudaError_t e__ = (x);                        \
        if (e__ != cudaSuccess) { kolm_set_cuda_error(e__, __FILE__, __LINE__); return KOLM_E_CUDA; } \
    } while (0)
#define KOLM_TRY(x) do { int r__ = (x); if (r__ != KOLM_OK) return r__; } while (0)
void kolm_set_cuda_error(cudaError_t e, const char* file, int line);

// ------------------------------------------------------------------------------------------------
// context: all scratch is allocated once at create time (no allocation in the hot path)
// ------------------------------------------------------------------------------------------------
struct kolm_ctx {
    int device;
    size_t max_elems;        // capacity in padded elements
    int max_blocks;
    int max_tiles;
    int sm_count;
    // per-block
    BlockInfo* d_binfo;      // [max_blocks]
    u32* d_btile0;           // [max_blocks] first static tile of block
    u32* d_btilen;           // [max_blocks] number of static tiles
    u32* d_atile0;           // [max_blocks] first active tile
    u32* d_atilen;           // [max_blocks]
    u32* d_active;           // [max_blocks] active records of the block this round
    u32* d_newcls;           // [max_blocks] classes created this round
    u32* d_done;             // [max_blocks]
    u32* d_nfac;             // [max_blocks] Lyndon factor count
    u32* d_stats;            // [16] device counters
    u32* h_stats;            // pinned mirror
    u64* d_bacc;             // [max_blocks*64] per-block 64-bit accumulators (Rice cost sums, ...)
    u64* h_bacc;             // pinned mirror
    i64* d_poff; i64* h_poff;      // [max_blocks+1] payload offsets
    int* d_params; int* h_params;  // [4*max_blocks]
    i64* d_sizes; i64* h_sizes;    // [5*max_blocks]
    // tiles
    TileDesc* d_tiles;       // static tiles
    TileDesc* d_atiles;      // active tiles
    u64* d_lb;               // look-back state: [0] ticket, [8..8+max_tiles) tile states
    u32* d_thist;            // [max_tiles*256]
    // per-element (padded index space)
    u32 *d_k0, *d_v0, *d_k1, *d_v1;   // sort ping-pong
    u32 *d_sa;               // current order
    u32 *d_rank;             // rank by position (block-local group start)
    u32 *d_nr;               // new ranks aligned with sorted records
    u32 *d_lo;               // deep bootstrap: low key half by position (k_boot_lo -> k_rerank<2>); local rounds: fallback flags by order index
    u32 *d_grp;              // group start of every order index (block-local), the local refinement rounds' view of the partition
    u8  *d_live;             // [max_tiles] static tile still holds unsettled records
    u32 *d_lact;             // [max_blocks] unsettled records at the start of a local round
    u32 *d_single;           // bitmap by position: rank is final and unique
    u32 *d_fstart;           // Lyndon factor starts, front-packed per block (block-local positions)
    u8  *d_tmp8a, *d_tmp8b;  // byte staging (bbwt out -> mtf -> rice)
    void* d_jump;            // inverse BBWT pointer-jumping nodes (2 x 16 B / element), allocated on first decode
    int light, light_ok;             // KOLM_CTX_REPAIR_ONLY context (kolm_create_ex): only kolm_repair_enc may run on it
    void* d_rpb; size_t rpb_bytes;   // slab pool of the incremental Re-Pair kernel (repair_big.cu), allocated on first use
    int* d_err; int* h_err;  // [max_blocks] per-block decode status
    // host mirrors
    BlockInfo* h_binfo;      // pinned
    u32* h_u32;              // pinned scratch [4*max_blocks]
    int nblocks; u32 total_elems; u32 max_len; int ntiles;
    i64 total_bytes;
    // accounting / profiling
    u32 static_rows, active_rows;   // max tiles per block of the static / active tile map
    i64 launches[32]; i64 algbytes[32];
    int prof_on; int prof_n; int prof_cat[8192]; cudaEvent_t* prof_ev;
    i64 counters[8];
    u32 sort_serial;         // sorts run on this context (stamps of the local rounds' per-group flags)
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// 1-D TMA (cp.async.bulk) global -> shared with mbarrier completion.  SASS: UBLKCP.
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KOLM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KOLM_DONE;\n\t"
        "bra KOLM_WAIT;\n\t"
        "KOLM_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------------
// decoupled look-back over tiles of one block (single pass scans).
//   state word: [63:62] status (0 invalid, 1 aggregate, 2 inclusive prefix), [61:0] payload.
//   Tiles are taken in ticket order (atomic counter) so predecessors are always resident.
// ------------------------------------------------------------------------------------------------
#define LB_INVALID 0ull
#define LB_AGG 1ull
#define LB_PREFIX 2ull
#define LB_PAYLOAD_MASK ((1ull << 62) - 1)

// header words: [0] ticket counter, [1] rows (0: ticket == tile index), [2] nblocks, [3] tile0 ptr, [4] tilen ptr, [5] group size.
// With rows > 0 tickets are handed out "row-major": ticket t -> tile k = t / nblocks of block b = t % nblocks, so the
// co-resident CTAs cover a few tiles of EVERY block instead of many tiles of a few blocks; each block's look-back chain
// then finds a published prefix within one 32-wide window.  A tile's predecessor always has a smaller ticket.
#define LB_NO_TILE 0xffffffffu
__device__ __forceinline__ u32 lb_take_ticket(u64* lb) {
    __shared__ u32 s_ticket;
    if (threadIdx.x == 0) {
        u32 t = (u32)atomicAdd((unsigned long long*)lb, 1ull);
        u32 tile = t;
        u64 rows = lb[1];
        if (rows) {
            u32 nb = (u32)lb[2], G = (u32)lb[5];             // groups of G blocks: row-major inside a group, groups in sequence
            u32 per = (u32)rows * G;
            u32 g = t / per, tg = t - g * per;
            u32 k = tg / G, b = g * G + (tg - k * G);
            const u32* t0 = (const u32*)lb[3]; const u32* tn = (const u32*)lb[4];
            tile = (b < nb && k < tn[b]) ? t0[b] + k : LB_NO_TILE;
        }
        s_ticket = tile;
    }
    __syncthreads();
    return s_ticket;
}
__device__ __forceinline__ void lb_store(u64* st, u64 status, u64 payload) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(st), "l"((status << 62) | (payload & LB_PAYLOAD_MASK)) : "memory");
}
__device__ __forceinline__ u64 lb_load(const u64* st) {
    u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(st) : "memory"); return v;
}
// Called by ALL lanes of warp 0 (others must not call).  Returns the exclusive prefix of `tile`.
template <class Op>
__device__ __forceinline__ u64 lb_exclusive(u64* lb, u32 tile, bool first_in_block, u64 aggregate, u64 identity, Op op) {
    u64* states = lb + 8;
    const u32 lane = lane_id();
    if (first_in_block) {
        if (lane == 0) lb_store(states + tile, LB_PREFIX, aggregate);
        return identity;
    }
    if (lane == 0) lb_store(states + tile, LB_AGG, aggregate);
    u64 excl = identity;
    i64 t = (i64)tile - 1;
    for (;;) {
        i64 mine = t - lane;
        u64 v = 0; u32 status; u32 pmask, imask;
        do {   // wait until every tile between us and the nearest published prefix is valid
            if (mine >= 0) { v = lb_load(states + mine); status = (u32)(v >> 62); }
            else { v = identity; status = (u32)LB_PREFIX; }
            pmask = __ballot_sync(0xffffffffu, status == (u32)LB_PREFIX);
            imask = __ballot_sync(0xffffffffu, status == (u32)LB_INVALID);
            if (pmask) imask &= ((pmask & (0u - pmask)) - 1u);
        } while (imask);
        u32 upto = pmask ? (u32)(__ffs(pmask) - 1) : 31u;    // nearest prefix lane (inclusive)
        u64 val = (lane <= upto) ? (v & LB_PAYLOAD_MASK) : identity;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val = op(val, __shfl_xor_sync(0xffffffffu, val, o));
        excl = op(excl, val);
        if (pmask) break;
        t -= 32;
    }
    if (lane == 0) lb_store(states + tile, LB_PREFIX, op(excl, aggregate));
    return excl;
}

struct OpAdd { __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a + b; } };
struct OpMax { __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a > b ? a : b; } };
// two independent 31-bit lanes: hi = bits [61:31], lo = bits [30:0]
struct OpMax2 {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const {
        u64 ah = a >> 31, bh = b >> 31, al = a & 0x7fffffffull, bl = b & 0x7fffffffull;
        return ((ah > bh ? ah : bh) << 31) | (al > bl ? al : bl);
    }
};
// hi = running minimum (31 bit), lo = running sum (31 bit)
struct OpMinAdd {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const {
        u64 ah = a >> 31, bh = b >> 31, al = a & 0x7fffffffull, bl = b & 0x7fffffffull;
        return ((ah < bh ? ah : bh) << 31) | ((al + bl) & 0x7fffffffull);
    }
};

// block-wide inclusive scan of one u64 per thread (KOLM_THREADS threads) with an associative op.
template <class Op>
__device__ __forceinline__ u64 block_scan_incl(u64 v, u64 identity, Op op, u64* s_warp /*[THREADS/32]*/, u64* total) {
    const u32 lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u64 n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v = op(n, v); }
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    u64 pre = identity, tot = identity;
#pragma unroll
    for (int i = 0; i < KOLM_THREADS / 32; ++i) { u64 x = s_warp[i]; if ((u32)i < w) pre = op(pre, x); tot = op(tot, x); }
    __syncthreads();
    if (total) *total = tot;
    return op(pre, v);
}

// ------------------------------------------------------------------------------------------------
// launch accounting + optional per-category CUDA-event timing (bench.py's roofline numbers)
// ------------------------------------------------------------------------------------------------
enum KernelCat { KC_TILES, KC_BOOT, KC_HIST, KC_SCAN, KC_SCATTER, KC_RERANK, KC_APPLY, KC_GATHER, KC_PLAN, KC_LYNDON, KC_EMIT,
                 KC_MTF_PRE, KC_MTF_SCAN, KC_MTF_MAIN, KC_RICE_COST, KC_RICE_PLAN, KC_RICE_PACK, KC_ZERO, KC_INV, KC_MISC, KC_COUNT };
#define KOLM_PROF_MAX 8192
static inline void prof_begin(kolm_ctx* c, int cat, i64 bytes, cudaStream_t s) {
    c->launches[cat]++; c->algbytes[cat] += bytes;
    if (c->prof_on && c->prof_n < KOLM_PROF_MAX) { c->prof_cat[c->prof_n] = cat; cudaEventRecord(c->prof_ev[2 * c->prof_n], s); }
}
static inline void prof_end(kolm_ctx* c, cudaStream_t s) {
    if (c->prof_on && c->prof_n < KOLM_PROF_MAX) { cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], s); c->prof_n++; }
}
#define KL(c, cat, bytes, s, ...) do { prof_begin((c), (cat), (i64)(bytes), (s)); __VA_ARGS__; prof_end((c), (s)); } while (0)

// host-side helpers shared between translation units
int kolm_set_batch(kolm_ctx* c, const i64* off_host, int nblocks, cudaStream_t s);
// zero the look-back state for a launch over the static (active=false) or active tile map; returns the grid size
int kolm_lb_reset(kolm_ctx* c, bool active, int ntiles, int* grid, cudaStream_t s);
int kolm_lb_reset_mode(kolm_ctx* c, bool active, int ntiles, int* grid, int mode, cudaStream_t s);
