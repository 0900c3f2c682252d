/* kolm_abi.h — C ABI of libkolm_b200.so: the B200-native (sm_100a) per-block transform
 * encode/decode hot path of KolmogorovLike-DataCompressor.
 *
 * The reference has no FFI layer; its de-facto operator API is the set of pure per-block
 * functions listed below.  Each entry point replaces the named reference
 * function(s) for a *batch* of independent blocks:
 *
 *   KF  = final/kolm_final.py                              ('KOLM' container)
 *   V22 = final_researched/kolm_final_researched_v2-2.py   ('KOLR' container)
 *
 * Conventions
 *   - `in`, `out`, payload buffers: raw DEVICE pointers owned by the caller.
 *   - `off`: HOST array of nblocks+1 byte offsets (off[0]=0 is not required; block b is
 *     bytes [off[b], off[b+1]) of every per-byte buffer of the call).
 *   - every call is asynchronous on `stream` except where it returns host scalars
 *     (those calls synchronise `stream` before returning).
 *   - return value: 0 = ok, <0 = KOLM_E_* (kolm_strerror()).  No exceptions, no allocation
 *     in the hot path; all scratch lives in the context.
 *   - a context is bound to one device and must not be used from two threads at once;
 *     distinct contexts are independent.
 *   - encoders with an `out_cap`: the exact total payload size is known on the device before any byte is
 *     emitted; when it exceeds out_cap (the Rice packers, which store whole words, want 8 bytes of slack) the
 *     emit kernels write NOTHING and the call returns KOLM_E_CAPACITY with out_off filled in, so the caller can
 *     retry with out_off[nblocks] (+ 8) bytes.  Worst cases: residual / LZ77 2n, KF model 2 ~2.2n, V22 Rice 8.25n + 8
 *     per block, Re-Pair 5n + 16 per block (ULEB128 symbols of blocks >= 512 MiB need five bytes).
 */
#ifndef KOLM_ABI_H
#define KOLM_ABI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kolm_ctx kolm_ctx;
typedef void* kolm_stream_t; /* cudaStream_t */

#define KOLM_PROFILE_KOLM 1 /* kolm_final.py: models 0..3 */
#define KOLM_PROFILE_KOLR 2 /* kolm_final_researched_v2-2.py: models 0..10 */

/* error codes */
#define KOLM_OK 0
#define KOLM_E_CUDA (-1)
#define KOLM_E_ARG (-2)
#define KOLM_E_CAPACITY (-3)
#define KOLM_E_TRUNCATED (-4)   /* reference raises EOFError   */
#define KOLM_E_CORRUPT (-5)     /* reference raises ValueError */
#define KOLM_E_UNSUPPORTED (-6)
#define KOLM_E_INDEX (-7)       /* reference raises IndexError */

int kolm_abi_version(void);
const char* kolm_strerror(int code);
const char* kolm_last_cuda_error(void);

/* context: scratch for batches of up to max_batch_bytes bytes in up to max_blocks blocks */
size_t kolm_scratch_bytes(size_t max_batch_bytes, int max_blocks);
int kolm_create(int device, size_t max_batch_bytes, int max_blocks, kolm_ctx** out);
/* flags = KOLM_CTX_REPAIR_ONLY: a context on which only kolm_repair_enc may run (every other operator returns
 * KOLM_E_UNSUPPORTED); it holds 8 instead of 38 bytes of scratch per element, so one context can take a whole container's
 * long blocks for the Re-Pair candidate (repair_compress, V22.py:1841-1911) beside the contexts of the other candidates. */
#define KOLM_CTX_REPAIR_ONLY 1u
int kolm_create_ex(int device, size_t max_batch_bytes, int max_blocks, unsigned flags, kolm_ctx** out);
void kolm_destroy(kolm_ctx* ctx);

/* ---- stage operators (batched) ------------------------------------------------------------- */

/* duval_lyndon  KF.py:200-225 == V22.py:326-349.
 * start_flags[i] = 1 iff a Lyndon factor starts at byte i (same indexing as `in`). */
int kolm_lyndon(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* start_flags, kolm_stream_t stream);

/* bbwt_forward  KF.py:227-325 == V22.py:351-423 */
int kolm_bbwt_fwd(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream);
/* bbwt_inverse  KF.py:327-369 == V22.py:425-454 */
int kolm_bbwt_inv(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream);

/* mtf_encode / mtf_decode  KF.py:375-405 == V22.py:460-478 */
int kolm_mtf_enc(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream);
int kolm_mtf_dec(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, kolm_stream_t stream);

/* KF model-2 token coder on an MTF sequence: choose_rice_grid/cost_gamma + BitWriter pack
 * (KF.py:499-529, 636-691).  Payload b is written at out + out_off[b]; out_off (HOST, nblocks+1)
 * receives the exclusive prefix sums of the payload sizes.  `out` must hold out_cap bytes.
 * params (HOST, 4*nblocks ints, may be NULL): k0, k1, use_rice_zero, use_rice_nz per block. */
int kolm_rice_kf_enc(kolm_ctx* ctx, const uint8_t* mtf, const int64_t* off, int nblocks, uint8_t* out, size_t out_cap,
                     int64_t* out_off, int* params, kolm_stream_t stream);
/* decode_model_bbwt_mtf's bit parser (KF.py:771-794): payloads -> MTF sequences of orig length */
int kolm_rice_kf_dec(kolm_ctx* ctx, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks,
                     uint8_t* mtf_out, kolm_stream_t stream);

/* V22 models 2-6: byte transform (flags 0,1,4,8,16 = none, 8x8 bit-plane, nibble swap, bit reverse,
 * Gray) + rice_encode(k=2)  (V22.py:1100-1120, 1413-1421, 1650-1680, 2044-2065).
 * sizes (HOST, 5*nblocks int64, may be NULL) receives every variant's exact payload size in the
 * order flags {0,1,4,8,16}; the variant `flags` is packed. */
int kolm_rice_k2_enc(kolm_ctx* ctx, const uint8_t* mtf, const int64_t* off, int nblocks, int flags, uint8_t* out, size_t out_cap,
                     int64_t* out_off, int64_t* sizes, kolm_stream_t stream);
/* Both coders of one MTF batch in one call: the cost pass reads the MTF bytes ONCE for the KF parameters / size and for the five
 * V22 variants' sizes (what one iteration of both references' selection loops needs, KF.py:821-864 / V22.py:2350-2369), then the KF
 * payload and the V22 variant `k2_flags` are packed.  Arguments as in kolm_rice_kf_enc / kolm_rice_k2_enc. */
int kolm_rice_dual_enc(kolm_ctx* ctx, const uint8_t* mtf, const int64_t* off, int nblocks, int k2_flags, uint8_t* kf_out, size_t kf_cap,
                       int64_t* kf_off, int* kf_params, uint8_t* k2_out, size_t k2_cap, int64_t* k2_off, int64_t* k2_sizes, kolm_stream_t stream);
int kolm_rice_k2_dec(kolm_ctx* ctx, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, int flags,
                     uint8_t* mtf_out, kolm_stream_t stream);

/* LZ77: exhaustive nearest-among-longest match (min 3, overlap allowed) + greedy parse, tokens [0,byte] / [1,ULEB len,ULEB dist].
 *   KF  encode_model_lz77 (KF.py:567-617):  window 255,  max_len 127
 *   V22 encode_lz77 (V22.py:1711-1763):     window 4096, max_len 0 (= unbounded)
 * Any other (window, max_len) is a parameterisation of the same semantics (BASELINE cfg 3 uses 65536 / 0). */
int kolm_lz77_enc(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint32_t window, uint32_t max_len, uint8_t* out,
                  size_t out_cap, int64_t* out_off, kolm_stream_t stream);
/* decode_model_lz77 (KF.py:723-760; window_check 0) / decode_lz77 (V22.py:1765-1812; window_check 4096) */
int kolm_lz77_dec(kolm_ctx* ctx, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint32_t window_check,
                  uint8_t* out, kolm_stream_t stream);

/* byte-predictor residual coders, each residual one ULEB128 value.  kind 0 = XOR (KF.py:545-565, 702-721),
 * 1 = delta (V22 "xor", V22.py:2105-2122), 2 = LFSR predictor (V22.py:1984-2019).
 * kolm_residual_sizes: exact payload sizes of all three kinds (HOST int64[3*nblocks]) from one read. */
int kolm_residual_sizes(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, int64_t* sizes3, kolm_stream_t stream);
int kolm_residual_enc(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, int kind, uint8_t* out, size_t out_cap,
                      int64_t* out_off, kolm_stream_t stream);
int kolm_residual_dec(kolm_ctx* ctx, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, int kind,
                      uint8_t* out, kolm_stream_t stream);

/* Re-Pair grammar candidate: repair_compress / repair_decompress (V22.py:1841-1911, 1916-1978).
 * Blocks of up to kolm_repair_max_block() bytes are coded by one CTA with the sequence in shared memory; longer blocks take the
 * exact incremental kernel (one CTA per block in a slab of global memory, ~150 bytes per input byte, allocated on first use and
 * kept by the context; KOLM_E_CAPACITY when not even one slab fits in free memory). */
int kolm_repair_enc(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, size_t out_cap, int64_t* out_off,
                    kolm_stream_t stream);
int kolm_repair_dec(kolm_ctx* ctx, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint8_t* out,
                    kolm_stream_t stream);
int kolm_repair_max_block(void);

/* decode_new_pipeline  V22.py:1578-1648 (method 10, "v2_new": header, eight bit planes RAW or Rice(k)-coded run lengths of the
 * plane's BBWT, inverse of the circuit_map_automaton model V22.py:1054-1092).  The shipped reference never emits this method
 *  but decodes it.  The context must have been created for >= 8x the batch bytes and 8x the blocks (the planes
 * are inverse-transformed as one batch of 8*nblocks blocks); otherwise KOLM_E_CAPACITY.  ValueError cases -> KOLM_E_CORRUPT. */
/* encode_new_pipeline  V22.py:1498-1576 with circuit_map_automaton_forward(parallel=False) (V22.py:1013-1052: 13 candidate
 * models scored by fp64 zero-order entropy, _pick_better tie rules :936-946), then per plane RAW vs BBWT -> RLE -> Rice(best k of
 * 0..15).  The shipped reference raises NameError before reaching this code, so the KOLR drop-in offers the
 * candidate only as an opt-in; the bytes equal the reference function's output once that path is taken.  Same context
 * requirement as kolm_v2new_dec.  The entropy scores are computed on the host in fp64 (glibc log2) from device histograms. */
int kolm_v2new_enc(kolm_ctx* ctx, const uint8_t* in, const int64_t* off, int nblocks, uint8_t* out, size_t out_cap, int64_t* out_off,
                   kolm_stream_t stream);
int kolm_v2new_dec(kolm_ctx* ctx, const uint8_t* payload, const int64_t* pay_off, const int64_t* off, int nblocks, uint8_t* out,
                   kolm_stream_t stream);

/* Content-defined chunking, HOST buffers (adjacent to the hot path; bit-identical boundaries are a precondition of
 * byte-identical containers).  ends[i] = exclusive end of chunk i; returns the chunk count or a negative error.
 *   kolm_cdc_kf : cdc_fast_boundaries (KF.py:161-194), gear table of KF.py:148-159
 *   kolm_cdc_v22: cdc_fast_boundaries_strict (V22.py:210-309, merge_orphan_tail=True), gear table of V22.py:152-167 */
int64_t kolm_cdc_kf(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, int64_t* ends, int64_t cap);
int64_t kolm_cdc_v22(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, int64_t* ends, int64_t cap);

/* GPU-assisted chunking: the per-byte scan of the two functions above runs on the device, the chain over
 * chunks on the host.  Both reference loops roll h = (h << 1) + GEAR[byte] from h = 0 at start+min_size (KF.py:176-190,
 * V22.py:262-296); once k bytes are in the hash its low k bits depend on the last k bytes only, i.e. on the position alone.
 *   kolm_cdc_candidates : evaluates that window test at every position lo <= p < hi of d_data (device pointer; bytes
 *                         d_data[max(0, lo-32) .. lo) are read as history) and writes d_out[0] = number found,
 *                         d_out[1 + i] = ((base + p) << 1) | strict  in arbitrary order.  variant 0 = KF gear, one mask of
 *                         k bits (strict == loose); variant 1 = V22 gear, loose mask k-2 bits, strict mask k+2 bits
 *                         (V22.py:233-240).  d_out must hold 1 + cap words; *count (host) receives the number found;
 *                         KOLM_E_CAPACITY if it exceeds cap (the list is then incomplete: enlarge cap or use kolm_cdc_*).
 *   kolm_cdc_walk_kf / _v22 : host.  data = the same bytes in host memory (only the <= 19 bytes after each start+min_size are
 *                         read: the truncated hash), cand/ncand = the list (sorted in place).  Return value, ends[] and error
 *                         codes exactly as kolm_cdc_kf / kolm_cdc_v22. */
int kolm_cdc_candidates(kolm_ctx* ctx, const uint8_t* d_data, int64_t lo, int64_t hi, int64_t base, int variant, int64_t avg_size,
                        uint64_t* d_out, int64_t cap, int64_t* count, cudaStream_t stream);
int64_t kolm_cdc_walk_kf(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, uint64_t* cand, int64_t ncand,
                         int64_t* ends, int64_t cap);
int64_t kolm_cdc_walk_v22(const uint8_t* data, int64_t n, int64_t min_size, int64_t avg_size, int64_t max_size, uint64_t* cand, int64_t ncand,
                          int64_t* ends, int64_t cap);

/* Per-block model selection: _encode_block KF.py:821-864 (ids 0..3 in order, keep `plen < best`) and the selection loops of
 * V22.py:2233-2252 / 2350-2369 over _select_encoders() (strict '<'): the winner is the FIRST minimum of the candidates' exact
 * payload sizes — the lowest id on ties.  sizes: HOST array, row-major [nblocks][ncand], as produced by the candidates'
 * encoders / cost kernels (a candidate the reference skips — it raised — carries a huge size); method[b] (HOST) receives the
 * winner's index, best[b] its size.  The comparison runs on the device (one thread per block). */
int kolm_select_blocks(kolm_ctx* ctx, const int64_t* sizes, int nblocks, int ncand, int32_t* method, int64_t* best, kolm_stream_t stream);

/* Payload gather after model selection: block b's winning payload is len[b] bytes at DEVICE address src_addr[b] (any of
 * the per-model payload buffers, or the input itself for RAW); they are laid out back to back at `out` in block order —
 * the payload area of a container (KF.py:896-901; V22.py:2443-2444).  src_addr/len: HOST arrays; out_off (HOST, nblocks+1)
 * receives the exclusive prefix sums. */
int kolm_gather_payloads(kolm_ctx* ctx, const uint64_t* src_addr, const int64_t* len, int nblocks, uint8_t* out, int64_t* out_off,
                         kolm_stream_t stream);

/* Batched device-to-device copy: len[b] bytes from src_addr[b] to dst_addr[b] (absolute DEVICE addresses in HOST arrays).
 * Used on the decode side to put each method group's decoded blocks at their final offsets. */
int kolm_copy_blocks(kolm_ctx* ctx, const uint64_t* src_addr, const uint64_t* dst_addr, const int64_t* len, int nblocks, kolm_stream_t stream);

/* ---- fused hot path: what compress() / decompress() call per batch ---------------- */
/* Per-block model selection with every candidate evaluated on the device: _encode_block (KF.py:821-864, ids 0..3 = raw, xor,
 * bbwt+mtf+model 2, lz77) for KOLM_PROFILE_KOLM; the selection loops of V22.py:2233-2252 / 2350-2369 over _select_encoders()
 * (V22.py:2152-2178, ids 0..9 = raw, xor[delta], bbwt, bbwt_bp, bbwt_nib, bbwt_br, bbwt_gray, lz77, lfsr_pred, repair; id 10 = v2_new
 * raises in the shipped reference and is never selected) for KOLM_PROFILE_KOLR.  The winner is the first minimum of the exact
 * payload sizes (strict '<': the lowest id on ties).  Sizes, offsets and method ids stay in device memory until the single copy home
 * at the end; only the winners are emitted, directly at their final offsets.
 *   off[0] must be 0.  cand_mask: bit id = candidate offered (0 = all; raw is always offered).
 *   ext_id >= 0: candidate ext_id was computed elsewhere (e.g. Re-Pair of long blocks on a KOLM_CTX_REPAIR_ONLY context): ext_sizes
 *                (HOST int64[nblocks]) and ext_addr (HOST: DEVICE address of each block's payload); -1 = none.
 *   scratch (DEVICE, 256-byte aligned, kolm_encode_blocks_scratch(profile, batch bytes, nblocks) bytes).
 *   payload_out (DEVICE, cap bytes, 8 bytes of slack wanted): the winners' payloads back to back in block order (KF.py:896-901;
 *                V22.py:2443-2444); payload_off (HOST, nblocks+1); method_ids (HOST u8[nblocks]); sizes_out (HOST, may be NULL):
 *                every candidate's size, row-major [nblocks][4 or 10], 2^62-1 for a candidate that was not offered.
 *   Re-Pair (V22.py:1841-1911, the last candidate of the list) of blocks of up to 16 KiB is evaluated after all the others and, when
 *   sizes_out is NULL, only as far as it can still win: a block's rounds stop once a lower bound on the final payload (the rules
 *   made so far + one left symbol per distinct adjacent pair of the current sequence + 7) reaches the best other size.  The
 *   selection is unchanged — the reference picks Re-Pair on a strictly smaller payload only; KOLM_REPAIR_STOP=0 runs every block
 *   to the end.  kolm_encode_blocks_stats: out2[0] = blocks of the last call whose Re-Pair candidate stopped early. */
size_t kolm_encode_blocks_scratch(int profile, size_t batch_bytes, int nblocks);
int kolm_encode_blocks(kolm_ctx* ctx, int profile, const uint8_t* in, const int64_t* off, int nblocks, uint32_t cand_mask,
                       int ext_id, const int64_t* ext_sizes, const uint64_t* ext_addr, uint8_t* scratch, size_t scratch_bytes,
                       uint8_t* payload_out, size_t cap, int64_t* payload_off, uint8_t* method_ids, int64_t* sizes_out,
                       kolm_stream_t stream);
int kolm_encode_blocks_stats(kolm_ctx* ctx, int64_t* out2);
/* The decode loops of decompress (KF.py:925-949 over _DECODERS; V22.py:2530-2540 over _select_decoders()): block b's payload is
 * payload[payload_start[b] .. + payload_len[b]) (DEVICE buffer, HOST arrays — KOLM containers interleave 9-byte block headers with
 * the payloads, so starts and lengths are separate), its method method_ids[b], and it decodes to
 * out[out_off[b] .. out_off[b+1]) (DEVICE).  Blocks are grouped by method and each group runs as one batch.  On a bad block the
 * reference's error class comes back (KOLM_E_TRUNCATED / CORRUPT / INDEX) and *bad_block (may be NULL) is the lowest failing block
 * index; an unknown method id is KOLM_E_CORRUPT.  Method 10 (v2_new) -> KOLM_E_UNSUPPORTED here (kolm_v2new_dec needs its own,
 * 8x larger context).  scratch: DEVICE, kolm_decode_blocks_scratch(payload bytes, output bytes, nblocks) bytes. */
size_t kolm_decode_blocks_scratch(size_t payload_bytes, size_t out_bytes, int nblocks);
int kolm_decode_blocks(kolm_ctx* ctx, int profile, const uint8_t* payload, const int64_t* payload_start, const int64_t* payload_len,
                       const uint8_t* method_ids, const int64_t* out_off, int nblocks, uint8_t* scratch, size_t scratch_bytes,
                       uint8_t* out, int* bad_block, kolm_stream_t stream);

/* ---- diagnostics --------------------------------------------------------------------------- */
/* counters of the last call on this context: [0] plain-suffix doubling rounds, [1] rotation doubling
 * rounds, [2] kernels launched since the last profile reset, [3] records sorted (sum over rounds) */
int kolm_last_counters(kolm_ctx* ctx, int64_t* out4);

/* launch accounting and optional CUDA-event timing per kernel category (used by bench.py):
 * enable(1) brackets every kernel launch with events on the launching stream; read() synchronises
 * the device and returns, per category, summed milliseconds, launch count and algorithmic bytes
 * since the last reset(). */
int kolm_profile_categories(void);
const char* kolm_profile_name(int cat);
int kolm_profile_enable(kolm_ctx* ctx, int on);
int kolm_profile_reset(kolm_ctx* ctx);
int kolm_profile_read(kolm_ctx* ctx, double* ms, int64_t* launches, int64_t* algbytes);

#ifdef __cplusplus
}
#endif
#endif
