// cdc.cu — content-defined chunking with the O(n) part on the GPU.
//
// Replaces the scan inside cdc_fast_boundaries (kolm_final.py:161-194) and cdc_fast_boundaries_strict
// (kolm_final_researched_v2-2.py:210-309).  Both reference loops roll  h = (h << 1) + GEAR[byte]  from h = 0 at
// start+min_size and cut where the low k bits of h are zero.  Bit j of h only depends on the last j+1 bytes, so once k
// bytes of a chunk have been consumed the test equals a test on the WINDOW hash of the last k bytes — a function of the
// position alone.  k_cdc_candidates evaluates that window test at every position in parallel and emits the (rare)
// positions that pass; the chain over chunks — which needs the previous cut, the min/avg/max rules and, for the first k-1
// bytes after start+min_size, the truncated hash — is a walk over that short list (kolm_cdc_walk_*, host C++).
// The result is bit-identical to kolm_cdc_kf / kolm_cdc_v22 (tests/test_gpu_cdc.py compares them and the oracle).
#include "common.cuh"

#define CDC_SEG 256          // positions per thread
#define CDC_WARM 32          // bytes of history that determine all 32 bits of h (k <= 20 are used)

struct GearTab { u32 g[256]; };

// out[0] = number of candidates found (may exceed cap: then the list is truncated and the caller must fall back);
// out[1 + i] = (position << 1) | strict, for every position p in [lo, hi) whose window hash has (h & mask_loose) == 0;
// strict = ((h & mask_strict) == 0).  Positions are reported as base + p.  Order is arbitrary (the walker sorts).
__global__ void __launch_bounds__(256) k_cdc_candidates(const u8* __restrict__ data, i64 lo, i64 hi, i64 base, const GearTab tab,
                                                        u32 mask_loose, u32 mask_strict, unsigned long long* __restrict__ out, u64 cap) {
    __shared__ u32 G[256];
    G[threadIdx.x] = tab.g[threadIdx.x];
    __syncthreads();
    const i64 p0 = lo + ((i64)blockIdx.x * blockDim.x + threadIdx.x) * CDC_SEG;
    if (p0 >= hi) return;
    const i64 p1 = p0 + CDC_SEG < hi ? p0 + CDC_SEG : hi;
    u32 h = 0;
    for (i64 p = p0 >= CDC_WARM ? p0 - CDC_WARM : 0; p < p0; ++p) h = (h << 1) + G[data[p]];
    auto step = [&](u32 byte, i64 p) {
        h = (h << 1) + G[byte];
        if ((h & mask_loose) == 0) {
            unsigned long long idx = atomicAdd(out, 1ull);
            if (idx < cap) out[1 + idx] = ((unsigned long long)(base + p) << 1) | ((h & mask_strict) == 0 ? 1ull : 0ull);
        }
    };
    i64 p = p0;
    if ((((uintptr_t)(data + p0)) & 15) == 0) {
        for (; p + 16 <= p1; p += 16) {
            const uint4 q = *reinterpret_cast<const uint4*>(data + p);
            const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) step((w[i >> 2] >> (8 * (i & 3))) & 0xFF, p + i);
        }
    }
    for (; p < p1; ++p) step(data[p], p);
}

static void v22_gear_table(u32* g) { u32 x = 0x243F6A88u; for (int i = 0; i < 256; ++i) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; g[i] = x | 1u; } }
static int cdc_bits(i64 avg) { int bl = 0; for (i64 a = avg; a > 0; a >>= 1) ++bl; int k = bl - 1; if (k > 20) k = 20; if (k < 6) k = 6; return k; }
static const u32* cdc_gear(int variant) {
    static u32 G[2][256]; static bool init = false;
    if (!init) { kf_gear_table(G[0]); v22_gear_table(G[1]); init = true; }
    return G[variant ? 1 : 0];
}

extern "C" {

int kolm_cdc_candidates(kolm_ctx* c, const uint8_t* d_data, int64_t lo, int64_t hi, int64_t base, int variant, int64_t avg_size,
                        uint64_t* d_out, int64_t cap, int64_t* count, cudaStream_t s) {
    if (!c || !d_data || !d_out || !count || lo < 0 || hi < lo || cap < 0 || avg_size <= 0) return KOLM_E_ARG;
    *count = 0;
    if (hi == lo) return KOLM_OK;
    const int k = cdc_bits(avg_size);
    u32 ml, ms;
    if (variant == 0) ml = ms = (1u << k) - 1u;                                            // KF.py:170-172
    else { const int ks = (k + 2 <= 20) ? k + 2 : 20, kl = (k > 2) ? k - 2 : 1; ms = (1u << ks) - 1u; ml = (1u << kl) - 1u; }   // V22.py:233-240
    GearTab tab; memcpy(tab.g, cdc_gear(variant), sizeof tab.g);
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 8, s));
    const i64 nthreads = (hi - lo + CDC_SEG - 1) / CDC_SEG;
    const int grid = (int)((nthreads + 255) / 256);
    KL(c, KC_MISC, hi - lo, s, k_cdc_candidates<<<grid, 256, 0, s>>>(d_data, lo, hi, base, tab, ml, ms, (unsigned long long*)d_out, (u64)cap));
    unsigned long long found = 0;
    CUDA_TRY(cudaMemcpyAsync(&found, d_out, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *count = (int64_t)found;
    return found > (unsigned long long)cap ? KOLM_E_CAPACITY : KOLM_OK;
}

// cdc_fast_boundaries (KF.py:161-194) over a candidate list from kolm_cdc_candidates(variant 0).  cand is sorted in place.
int64_t kolm_cdc_walk_kf(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, uint64_t* cand, int64_t ncand,
                         int64_t* ends, int64_t cap) {
    if (n <= 0) return 0;
    const u32* G = cdc_gear(0);
    const int k = cdc_bits(avg_size);
    const u32 mask = (1u << k) - 1u;
    std::sort(cand, cand + ncand);
    i64 i = 0, cnt = 0, ci = 0;
    while (i < n) {
        const i64 start = i;
        const i64 emin = start + min_size < n ? start + min_size : n, emax = start + max_size < n ? start + max_size : n;
        const i64 tend = emin + (k - 1) < emax ? emin + (k - 1) : emax;      // fewer than k bytes consumed: truncated hash, computed here
        u32 h = 0; i64 cut = -1;
        for (i64 p = emin; p < tend; ++p) { h = (h << 1) + G[data[p]]; if ((h & mask) == 0) { cut = p + 1; break; } }
        if (cut < 0) {
            while (ci < ncand && (i64)(cand[ci] >> 1) < tend) ++ci;
            cut = (ci < ncand && (i64)(cand[ci] >> 1) < emax) ? (i64)(cand[ci] >> 1) + 1 : emax;
        }
        if (cnt >= cap) return KOLM_E_CAPACITY;
        ends[cnt++] = cut; i = cut;
        if (i == start) return KOLM_E_ARG;
    }
    return cnt;
}

// cdc_fast_boundaries_strict (V22.py:210-309, merge_orphan_tail=True) over a candidate list from kolm_cdc_candidates(variant 1)
int64_t kolm_cdc_walk_v22(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, uint64_t* cand, int64_t ncand,
                          int64_t* ends, int64_t cap) {
    if (n <= 0) return 0;
    if (!(min_size > 0 && min_size <= avg_size && avg_size <= max_size) || avg_size < 64) return KOLM_E_ARG;
    const u32* G = cdc_gear(1);
    const int k = cdc_bits(avg_size);
    const int ks = (k + 2 <= 20) ? k + 2 : 20, kl = (k > 2) ? k - 2 : 1;
    const u32 ms = (1u << ks) - 1u, ml = (1u << kl) - 1u;
    std::sort(cand, cand + ncand);
    i64 i = 0, cnt = 0, ci = 0, last_start = 0;
    while (i < n) {
        const i64 start = i, rem = n - start; last_start = start;
        if (cnt >= cap) return KOLM_E_CAPACITY;
        if (rem <= min_size) { ends[cnt++] = n; break; }
        const i64 lmax = rem < max_size ? rem : max_size, normal = avg_size < lmax ? avg_size : lmax;
        const i64 s0 = start + min_size, en = start + normal, el = start + lmax;
        const i64 tend = s0 + (ks - 1) < el ? s0 + (ks - 1) : el;             // truncated while fewer than ks bytes are in the hash
        u32 fp = 0; i64 cut = -1;
        for (i64 p = s0; p < tend; ++p) { fp = (fp << 1) + G[data[p]]; if ((fp & (p < en ? ms : ml)) == 0) { cut = p + 1; break; } }
        if (cut < 0) {
            while (ci < ncand && (i64)(cand[ci] >> 1) < tend) ++ci;
            for (i64 cj = ci; cj < ncand; ++cj) {
                const i64 p = (i64)(cand[cj] >> 1);
                if (p >= el) break;
                if (p >= en || (cand[cj] & 1u)) { cut = p + 1; break; }   // strict mask before the normal point, loose after
            }
            if (cut < 0) cut = el;
        }
        ends[cnt++] = cut; i = cut;
    }
    if (cnt >= 2 && ends[cnt - 1] - last_start < min_size) { ends[cnt - 2] = ends[cnt - 1]; --cnt; }   // orphan tail merged (V22.py:301-306)
    return cnt;
}

}  // extern "C"
