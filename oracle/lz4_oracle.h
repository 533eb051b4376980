/*
 * oracle/lz4_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's LZ4 block codec as lzbench drives it in
 * chunked mode.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path (lzbench_b200/csrc)
 * never links, loads or calls anything declared here.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 4), so the pin is the reference itself compiled here
 * (oracle/_ref/liblz4ref.so, built by oracle/Makefile from /root/reference/lz4/lz4.c):
 * tests/test_oracle.py checks that oracle_lz4_compress() is byte-identical to
 * LZ4_compress_default() and that oracle_lz4_decompress_safe() agrees with
 * LZ4_decompress_safe() on every corpus, chunk size and hand-built block of the
 * SURVEY section-4 accept/reject table; tests/golden/ holds vectors generated from _ref.
 */
#ifndef LZ4_ORACLE_H
#define LZ4_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lz4/lz4.h:171  LZ4_COMPRESSBOUND */
int oracle_lz4_compress_bound(int n);

/* lz4/lz4.c:1373 LZ4_compress_default -> :1354 -> :1284 -> :851 (acceleration = 1).
 * Returns bytes written, 0 if dst too small (limitedOutput path). */
int oracle_lz4_compress(const uint8_t* src, int n, uint8_t* dst, int cap);
/* LZ4_compress_fast, lz4/lz4.c:1352-1375 */
int oracle_lz4_compress_fast(const uint8_t* src, int n, uint8_t* dst, int cap, int acceleration);

/* lz4/lz4.c:2171 LZ4_decompress_safe semantics (endOnInputSize, full block, noDict).
 * Returns bytes written or a negative value on malformed input. */
int oracle_lz4_decompress_safe(const uint8_t* src, int clen, uint8_t* dst, int cap);

/* Sequence statistics gathered by the validator (SURVEY.md appendix B columns). */
typedef struct {
    uint64_t n_seq;          /* sequences, including the final literal-only one   */
    uint64_t lit_bytes;      /* total literal bytes                                */
    uint64_t match_bytes;    /* total match bytes                                  */
    uint64_t zero_lit_seq;   /* sequences that carry no literals (non-final)       */
    uint64_t overlap_seq;    /* matches with offset < length                       */
    uint64_t short_off_seq;  /* matches with offset < 16                           */
    uint64_t max_seq_in_block;
} oracle_seq_stats;

/* Walk one block and check the format rules the reference decoder enforces
 * (SURVEY.md section 4 table / appendix A): exact consumption of `clen`, exact
 * production of `n`, 1 <= offset <= position, last 5 bytes literal, last match
 * starts <= n-12, n < 13 => single literal run.  Returns 0 if valid, else a
 * positive rule number.  `stats` may be NULL; counters are accumulated. */
int oracle_lz4_validate(const uint8_t* src, int clen, int n, oracle_seq_stats* stats);

/* lzbench chunk loops, _lzbench/lzbench.cpp:266-298 and :301-329, driving a codec
 * given as lzbench-style function pointers (:113 typedef compress_func). */
typedef int64_t (*oracle_codec_fn)(char* in, size_t insize, char* out, size_t outsize,
                                   size_t level, size_t param2, char* workmem);

int64_t oracle_lzbench_compress(oracle_codec_fn compress, const size_t* chunk_sizes, size_t n_chunks,
                                size_t* compr_sizes, uint8_t* inbuf, uint8_t* outbuf, size_t outsize);
int64_t oracle_lzbench_decompress(oracle_codec_fn decompress, const size_t* chunk_sizes,
                                  const size_t* compr_sizes, size_t n_chunks,
                                  uint8_t* inbuf, uint8_t* outbuf);

/* The two lz4 plugin wrappers, _lzbench/compressors.cpp:343-346 and :358-362, on top of
 * the restated codec. */
int64_t oracle_plugin_lz4_compress(char* in, size_t insize, char* out, size_t outsize,
                                   size_t level, size_t p2, char* wm);
int64_t oracle_plugin_lz4_decompress(char* in, size_t insize, char* out, size_t outsize,
                                     size_t level, size_t p2, char* wm);

/* _lzbench/lzbench.h:17 GET_COMPRESS_BOUND */
size_t oracle_lzbench_compress_bound(size_t insize);

/* Multi-threaded driver used for the CPU baseline (bench.py cpu_baseline / --impl
 * reference): contiguous chunk ranges per thread, each thread running the serial
 * lzbench loop on its range.  `codec` is the compress or decompress entry point
 * (the port above, or the functions dlsym'd from oracle/_ref).  direction: 0 =
 * compress (fills compr_sizes and writes each chunk's output at its bound-strided
 * slot in outbuf), 1 = decompress from the same strided layout.  Returns the wall
 * time of the slowest thread in nanoseconds, or <0 on codec failure. */
int64_t oracle_mt_run(oracle_codec_fn codec, int direction, int n_threads,
                      const size_t* chunk_sizes, size_t* compr_sizes, size_t n_chunks,
                      uint8_t* plain, uint8_t* packed, size_t slot_stride);

#ifdef __cplusplus
}
#endif
#endif
