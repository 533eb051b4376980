/*
 * lz4b200.h -- C ABI of the B200-native LZ4 block codec for lzbench's chunked mode (-b#).
 *
 * Plain C, plain pointers and sizes: this is exactly what a reference-side binding links
 * against (see INTEGRATION.md for the rows/wrappers a maintainer adds to
 * _lzbench/compressors.cpp, compressors.h and lzbench.h).  Every entry point cites the
 * reference interface it replaces.  There is no CPU fallback anywhere behind this header:
 * every call either runs the sm_100a kernels or fails (negative return, message available
 * from lz4b200_last_error()).
 *
 * Data layout (identical to the reference's chunk loops):
 *   uncompressed image : chunk i at  sum(chunk_sizes[0..i))          (contiguous)
 *   compressed image   : chunk i at  sum(compr_sizes[0..i))          (contiguous, no header)
 *   each compressed chunk is one standard LZ4 block, or the raw bytes when
 *   compr_sizes[i] == chunk_sizes[i]  (_lzbench/lzbench.cpp:284-289, :311-315).
 */
#ifndef LZ4B200_H
#define LZ4B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZ4B200_VERSION_STRING "0.1"

typedef struct lz4b200_ctx lz4b200_ctx;

/* ---- library ------------------------------------------------------------------------- */

/* Library version string ("0.1"); the `version` column of the comp_desc[] row
 * (_lzbench/lzbench.h:142-219). */
const char* lz4b200_version(void);

/* Thread-local message for the last failing call in this thread ("" if none). */
const char* lz4b200_last_error(void);

/* Worst-case size of one compressed chunk: n + n/255 + 16 (lz4/lz4.h:171 LZ4_COMPRESSBOUND;
 * lz4/lz4.c LZ4_compressBound).  0 if n exceeds LZ4_MAX_INPUT_SIZE (lz4/lz4.h:170). */
size_t lz4b200_compress_bound(size_t n);

/* Number of CUDA devices visible, or a negative value if the runtime cannot be initialised. */
int lz4b200_device_count(void);

/* ---- context (replaces init_func / deinit_func, _lzbench/lzbench.h:114-115; precedent
 *      lzbench_nvcomp_init/_deinit, _lzbench/compressors.cpp:1856-1914) ------------------- */

/* Create a codec context bound to CUDA device `device`.  `max_total_bytes` and
 * `max_chunk_bytes` are sizing hints (0 allowed): buffers grow on demand.  One context
 * serves both directions; calls on one context must be serial (lzbench is single-threaded,
 * _lzbench/lzbench.cpp:479-534).  Returns 0 or a negative error. */
int lz4b200_create(int device, size_t max_total_bytes, size_t max_chunk_bytes, lz4b200_ctx** ctx);
void lz4b200_destroy(lz4b200_ctx* ctx);

/* ---- batched host-buffer entry points: the hot path ------------------------------------ */

/* Replaces the whole of lzbench_compress (_lzbench/lzbench.cpp:266-298) for one codec:
 * compresses n_chunks independent chunks of `in` (sizes chunk_sizes[]) into the contiguous
 * concatenation `out`, fills compr_sizes[], stores a chunk raw iff its LZ4 block would be
 * exactly chunk_sizes[i] bytes (:284-289).  Returns the total compressed size, 0 if it
 * does not fit in out_capacity (:286), negative on CUDA failure.  Host pointers; copies
 * to/from the device are inside the call. */
int64_t lz4b200_compress_batch(lz4b200_ctx* ctx, const char* in, const size_t* chunk_sizes,
                               size_t n_chunks, char* out, size_t out_capacity, size_t* compr_sizes);

/* Replaces the whole of lzbench_decompress (_lzbench/lzbench.cpp:301-329): chunk i is
 * copied raw iff compr_sizes[i] == chunk_sizes[i] (:311-315), otherwise decoded as one LZ4
 * block of exactly chunk_sizes[i] bytes.  Returns the total decompressed size, or a value
 * <= 0 if any chunk is malformed (:321; the bounds-checked decoder follows
 * LZ4_decompress_safe's rules, lz4/lz4.c:2171, not LZ4_decompress_fast's trust). */
int64_t lz4b200_decompress_batch(lz4b200_ctx* ctx, const char* in, const size_t* compr_sizes,
                                 const size_t* chunk_sizes, size_t n_chunks, char* out);

/* ---- device-resident entry points (inputs and outputs stay in HBM; used for the roofline
 *      measurement and by callers that already hold data on the GPU) ----------------------- */

/* All pointers are device pointers on the context's device; `stream` is a cudaStream_t
 * (0 = default stream).  Size arrays are uint64_t.  Calls are asynchronous on `stream`.
 * d_total_out receives the total compressed size.  Returns 0 on successful launch.
 * d_out must hold lz4b200_compress_device_bound(n_chunks, max_chunk_bytes) bytes: there is no capacity argument
 * and a smaller buffer is overrun on the device.  The device entry points share per-context scratch
 * (offsets, status, tickets, sequence table): calls on one context must be ordered -- issue them on the same
 * stream, or synchronize `stream` before the next call (device or host entry point) on that context. */
size_t lz4b200_compress_device_bound(size_t n_chunks, size_t max_chunk_bytes);
int lz4b200_compress_device(lz4b200_ctx* ctx, const void* d_in, const uint64_t* d_chunk_sizes,
                            size_t n_chunks, size_t max_chunk_bytes, void* d_out,
                            uint64_t* d_compr_sizes, uint64_t* d_total_out, void* stream);

/* d_bad_chunks (may be NULL) receives the number of malformed chunks (0 = success). */
int lz4b200_decompress_device(lz4b200_ctx* ctx, const void* d_in, const uint64_t* d_compr_sizes,
                              const uint64_t* d_chunk_sizes, size_t n_chunks, size_t max_chunk_bytes,
                              void* d_out, uint32_t* d_bad_chunks, void* stream);

/* Number of kernels this library launched through `ctx` since creation (for bench.py's
 * gpu_launches claim). */
uint64_t lz4b200_launch_count(const lz4b200_ctx* ctx);

/* Per-launch device time of the codec kernel (the dominant kernel of each direction), measured
 * with CUDA events on the launching stream.  Enable with lz4b200_set_option(ctx,"timing","1");
 * every compress/decompress call then records one event pair (up to 256 between reads).  This
 * call waits for the recorded launches, writes up to max_n durations in milliseconds in launch
 * order, clears the list and returns how many were written (negative on error). */
int lz4b200_kernel_times_ms(lz4b200_ctx* ctx, float* out_ms, int max_n);

/* Debug: cumulative per-phase cycle counters of the tiled kernels (enable with option "prof"="1");
 * copies 16 counters to out16 and clears them. */
int lz4b200_debug_prof(lz4b200_ctx* ctx, uint64_t* out16);

/* Options (key, value as strings).  Returns 0, or -1 for an unknown key or value.
 *   "kernels"  decode generation: "v0" warp per chunk; "v1" one CTA per chunk, TMA-staged tile, speculative parallel parse;
 *              "v2" sub-warp teams; "v3" + every output byte owned by a fixed lane, match store pipelined behind its load;
 *              "v4" + warp-uniform predicated fast path; "v5" (default) + compressed stream through a cp.async shared-memory
 *              ring, long literal runs on the fast path, slow path out of line; "v6" experiment: v5 with the pending match
 *              bytes staged by cp.async ("depth" = "2".."4" iterations in flight); "v7": sequence table by one lane per
 *              chunk, then one lane per sequence (kernels_v7.cuh), forced for every chunk size up to 64 KiB; "default":
 *              v7 for chunks above 16 KiB up to 64 KiB, v5 otherwise.  Also sets the matching compressor ("v0" / "v1": their own,
 *              "v2" .. "v7": v2, "default": default).
 *              The environment variable LZ4B200_KERNELS, read at lz4b200_create, is the same option.
 *   "team"     lanes per chunk for the team decoders: "4" / "8" / "16" / "32", "0" = by batch size (default)
 *   "compress" compressor generation: "v0" warp per chunk, "v1" one CTA per chunk, "v2" two warp groups per CTA, "v3"
 *              five stages on their own warps (kernels_c3.cuh; chunks up to 64 KiB at acceleration 1, the rest stays on v2),
 *              "default": v3 where it applies, v2 otherwise
 *   "accel"    LZ4_compress_fast's acceleration (lz4/lz4.c:1352-1375; the lz4fast levels of lzbench,
 *              _lzbench/compressors.cpp:348-356): 1 default, 2-3 probe every 2nd position, >= 4 every 4th
 *   "slices"   pipeline slices of the host entry points (0 = automatic); "ramp" = "0"/"1" small first slices when
 *              decompressing (default 1); "h2d" = "own"/"shared" stream for the inbound copies (default own)
 *   "host_register" = "1" page-locks caller buffers on first sight (what the lzbench wrappers turn on); a registered
 *              buffer must outlive the context or be released with lz4b200_unregister_host before it is freed
 *   "split"    v7: percent of a large batch decoded by the team decoder on a second stream ("0" default; measured slower);
 *              "coarse" = "0"/"1" (default 1): chunks above 64 KiB that are a handful of long copies are decoded piecewise by the
 *              whole machine (kernels_coarse.cuh);
 *              "v7big" = "0"/"1" (default 1): compressible chunks above 64 KiB (up to 16 MiB) through the sequence table too;
 *              "planes" = "8"/"16"/"32" chunks per warp in v7's table pass (default 32; measured: 16 the same, 8 slower)
 *   "timing" = "0"/"1" (see lz4b200_kernel_times_ms), "prof" = "0"/"1" (see lz4b200_debug_prof), "c1mode" (match-finder variant) */
int lz4b200_set_option(lz4b200_ctx* ctx, const char* key, const char* value);

/* Releases the page-lock that "host_register" put on the caller buffer containing `p` (0 = released, -1 = none
 * found).  The context keeps up to 16 registrations and drops the oldest when a 17th buffer comes along. */
int lz4b200_unregister_host(lz4b200_ctx* ctx, const void* p);

/* ---- lzbench plugin-signature wrappers (drop-in comp_desc[] row "lz4_b200") -------------- */

/* Exactly the typedefs at _lzbench/lzbench.h:113-115.  Called per chunk they run a
 * one-chunk batch, so `lzbench -elz4_b200 -b64` works on an unpatched harness next to
 * `-elz4` (_lzbench/compressors.cpp:343-362); INTEGRATION.md shows the <=30-line harness
 * patch that hands the whole chunk list to the batched entry points instead. */
char*   lzbench_lz4_b200_init(size_t insize, size_t level, size_t param2);
void    lzbench_lz4_b200_deinit(char* workmem);
int64_t lzbench_lz4_b200_compress(char* in, size_t insize, char* out, size_t outsize,
                                  size_t level, size_t param2, char* workmem);
int64_t lzbench_lz4_b200_decompress(char* in, size_t insize, char* out, size_t outsize,
                                    size_t level, size_t param2, char* workmem);

/* Batched forms behind the same workmem, for the patched harness: same arguments as
 * lzbench_compress / lzbench_decompress minus the function pointer. */
int64_t lzbench_lz4_b200_compress_batch(const size_t* chunk_sizes, size_t n_chunks, size_t* compr_sizes,
                                        char* in, char* out, size_t outsize, char* workmem);
int64_t lzbench_lz4_b200_decompress_batch(const size_t* chunk_sizes, const size_t* compr_sizes, size_t n_chunks,
                                          char* in, char* out, char* workmem);

/* ---- PCIe yardstick -------------------------------------------------
 * What the reference's `cudaMemcpy` row does (lzbench_cuda_memcpy, _lzbench/compressors.cpp:1828-1833): n bytes host ->
 * device -> host, no kernel, through this library's staging (page-locked buffers, overlapped slices).  Returns n. */
int64_t lz4b200_memcpy_batch(lz4b200_ctx* ctx, const char* in, size_t n, char* out);

/* ---- LZ4 frame container around the chunked path -------------------
 * A frame with independent blocks is the chunked hot path plus framing: magic, descriptor, one LE32 size word per
 * block (bit 31 = stored uncompressed), end mark, optional xxHash32 checksums (lz4/lz4frame.h:119-186 for the types
 * these arguments mirror).  Frames with linked blocks or a dictionary are outside the chunked path: rejected, -2. */
#define LZ4B200_FRAME_CONTENT_CHECKSUM 1   /* LZ4F_contentChecksumEnabled, lz4/lz4frame.h:147; serial, computed on the host */
#define LZ4B200_FRAME_BLOCK_CHECKSUM   2   /* LZ4F_blockChecksumEnabled,   lz4/lz4frame.h:154; per block, on the device   */
#define LZ4B200_FRAME_CONTENT_SIZE     4   /* frameInfo.contentSize = n,   lz4/lz4frame.h:180                             */
#define LZ4B200_FRAME_SKIP_CHECKSUMS   1   /* decompress flag: LZ4F_decompressOptions_t is not offered by 1.9.3; default verifies */

/* Upper bound of a frame holding n bytes in blocks of id block_id (4..7 = 64 KiB, 256 KiB, 1 MiB, 4 MiB;
 * LZ4F_blockSizeID_t, lz4/lz4frame.h:123-133).  Replaces LZ4F_compressFrameBound, lz4/lz4frame.h:215.  0 = bad id. */
size_t  lz4b200_frame_bound(size_t n, int block_id);
/* Replaces LZ4F_compressFrame (lz4/lz4frame.h:224, lz4/lz4frame.c:429-470) with blockMode = LZ4F_blockIndependent.
 * Host buffers.  Returns the frame size, or < 0. */
int64_t lz4b200_frame_compress(lz4b200_ctx* ctx, const char* in, size_t n, char* out, size_t out_capacity,
                               int block_id, int flags);
/* Replaces one complete LZ4F_decompress session (lz4/lz4frame.h:465, lz4/lz4frame.c:1384-1830) over one frame.
 * Returns the decoded size; -1 malformed / checksum mismatch / does not fit, -2 valid but outside the chunked path. */
int64_t lz4b200_frame_decompress(lz4b200_ctx* ctx, const char* in, size_t in_size, char* out, size_t out_capacity,
                                 int flags);
/* Replaces LZ4F_getFrameInfo(...).contentSize (lz4/lz4frame.h:430).  Host only.  -3 = the frame does not say. */
int64_t lz4b200_frame_content_size(const char* in, size_t in_size);

/* ---- synthetic corpora; host-side, deterministic ----------------------- */
int lz4b200_corpus_fill(int kind, uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
