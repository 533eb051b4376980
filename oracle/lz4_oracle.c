/*
 * oracle/lz4_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see lz4_oracle.h).
 *
 * Plain-C restatement of the reference's LZ4 block codec (lz4 1.9.3 as vendored by
 * lzbench) and of lzbench's two chunk loops.  Written from the algorithm, not copied:
 * index-based instead of pointer-based, one table type switch, no dictionary modes
 * (lzbench never uses them).  Each function cites the reference lines it follows.
 * Pinned byte-for-byte against oracle/_ref (the reference compiled here) by
 * tests/test_oracle.py.
 */
#define _GNU_SOURCE
#include "lz4_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* lz4/lz4.c:214-235 and lz4/lz4.h:556-558, :170 */
enum {
    MINMATCH = 4,
    LASTLITERALS = 5,
    MFLIMIT = 12,
    MIN_LENGTH = MFLIMIT + 1,            /* lz4/lz4.c:221 LZ4_minLength */
    ML_MASK = 15,
    RUN_MASK = 15,
    DISTANCE_MAX = 65535,
    LIMIT_64K = 65536 + (MFLIMIT - 1),   /* lz4/lz4.c:633 LZ4_64Klimit */
    SKIP_TRIGGER = 6,                    /* lz4/lz4.c:634 */
    HASHLOG_U32 = 12,                    /* lz4/lz4.h:578 LZ4_HASHLOG */
    HASHLOG_U16 = 13,
    MAX_INPUT = 0x7E000000
};

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

int oracle_lz4_compress_bound(int n)
{
    if (n < 0 || (unsigned)n > (unsigned)MAX_INPUT) return 0;
    return n + n / 255 + 16;
}

/* lz4/lz4.c:698-722: 4-byte multiplicative hash to 13 bits for the u16 table; 5-byte
 * hash of a 64-bit little-endian read to 12 bits for the u32 table. */
static inline uint32_t hash_at(const uint8_t* p, int tab16)
{
    if (tab16) return (rd32(p) * 2654435761U) >> (32 - HASHLOG_U16);
    return (uint32_t)(((rd64(p) << 24) * 889523592379ULL) >> (64 - HASHLOG_U32));
}

/* lz4/lz4.c:603-626 LZ4_count: common prefix length of src[a..] and src[b..], a < limit */
static inline uint32_t count_eq(const uint8_t* src, uint32_t a, uint32_t b, uint32_t limit)
{
    uint32_t s = a;
    while (a + 8 <= limit) {
        uint64_t d = rd64(src + a) ^ rd64(src + b);
        if (d) return (a - s) + (uint32_t)(__builtin_ctzll(d) >> 3);
        a += 8; b += 8;
    }
    while (a < limit && src[a] == src[b]) { a++; b++; }
    return a - s;
}

/* lz4/lz4.c:851-1240 LZ4_compress_generic_validated, noDict / noDictIssue, acceleration 1,
 * outputDirective in {notLimited, limitedOutput} picked as in :1290-1303. */
int oracle_lz4_compress(const uint8_t* src, int n, uint8_t* dst, int cap)
{
    return oracle_lz4_compress_fast(src, n, dst, cap, 1);
}

/* LZ4_compress_fast (lz4/lz4.c:1352-1375: acceleration < 1 -> 1, > 65537 -> 65537; the `lz4fast` rows of lzbench,
 * _lzbench/compressors.cpp:348-356): the skip counter of the search loop starts at acceleration << 6 (:985). */
int oracle_lz4_compress_fast(const uint8_t* src, int n, uint8_t* dst, int cap, int acceleration)
{
    if (acceleration < 1) acceleration = 1;
    if (acceleration > 65537) acceleration = 65537;
    uint32_t table[1 << HASHLOG_U32];            /* 16 KB, doubles as u16[8192] logically */
    if (n < 0 || (unsigned)n > (unsigned)MAX_INPUT) return 0;
    const int limited = cap < oracle_lz4_compress_bound(n);
    const int tab16 = n < LIMIT_64K;
    uint16_t* t16 = (uint16_t*)table;
    memset(table, 0, sizeof table);               /* :1286 LZ4_initStream */

#define TGET(h)    (tab16 ? (uint32_t)t16[h] : table[h])
#define TPUT(h, v) do { if (tab16) t16[h] = (uint16_t)(v); else table[h] = (v); } while (0)

    uint32_t ip = 0, anchor = 0, op = 0;
    const int64_t mflimit_plus1 = (int64_t)n - MFLIMIT + 1;  /* :883 */
    const uint32_t matchlimit = (uint32_t)(n - LASTLITERALS);

    if (n < MIN_LENGTH) goto last_literals;       /* :921 */

    TPUT(hash_at(src, tab16), 0);                 /* :924 */
    ip = 1;
    uint32_t fwd_h = hash_at(src + ip, tab16);

    for (;;) {
        uint32_t match, tok_pos;
        {   /* :983-1014 search loop with skip acceleration */
            uint32_t fwd_ip = ip, step = 1, nb = (uint32_t)acceleration << SKIP_TRIGGER;
            for (;;) {
                uint32_t h = fwd_h, cur = fwd_ip;
                uint32_t cand = TGET(h);
                ip = fwd_ip;
                fwd_ip += step;
                step = nb++ >> SKIP_TRIGGER;
                if ((int64_t)fwd_ip > mflimit_plus1) goto last_literals;
                fwd_h = hash_at(src + fwd_ip, tab16);
                TPUT(h, cur);
                if (!tab16 && cand + DISTANCE_MAX < cur) continue;       /* :1003 too far */
                if (rd32(src + cand) == rd32(src + ip)) { match = cand; break; }
            }
        }
        /* :1019 catch up */
        while (ip > anchor && match > 0 && src[ip - 1] == src[match - 1]) { ip--; match--; }

        {   /* :1022-1046 literal run */
            uint32_t lit = ip - anchor;
            tok_pos = op++;
            if (limited && (int64_t)op + lit + (2 + 1 + LASTLITERALS) + lit / 255 > cap) return 0;
            if (lit >= RUN_MASK) {
                uint32_t len = lit - RUN_MASK;
                dst[tok_pos] = RUN_MASK << 4;
                for (; len >= 255; len -= 255) dst[op++] = 255;
                dst[op++] = (uint8_t)len;
            } else dst[tok_pos] = (uint8_t)(lit << 4);
            memcpy(dst + op, src + anchor, lit);
            op += lit;
        }

        for (;;) {  /* :1048 _next_match */
            uint32_t off = ip - match;
            dst[op++] = (uint8_t)off; dst[op++] = (uint8_t)(off >> 8);          /* :1072 */
            uint32_t mc = count_eq(src, ip + MINMATCH, match + MINMATCH, matchlimit); /* :1092 */
            ip += mc + MINMATCH;
            if (limited && (int64_t)op + (1 + LASTLITERALS) + (mc + 240) / 255 > cap) return 0;
            if (mc >= ML_MASK) {                                               /* :1123-1135 */
                dst[tok_pos] += ML_MASK;
                mc -= ML_MASK;
                for (; mc >= 255; mc -= 255) dst[op++] = 255;
                dst[op++] = (uint8_t)mc;
            } else dst[tok_pos] += (uint8_t)mc;

            anchor = ip;
            if ((int64_t)ip >= mflimit_plus1) goto last_literals;              /* :1143 */
            TPUT(hash_at(src + ip - 2, tab16), ip - 2);                        /* :1146 */
            {   /* :1159-1196 immediate re-probe at ip */
                uint32_t h = hash_at(src + ip, tab16), cur = ip, cand = TGET(h);
                TPUT(h, cur);
                if ((tab16 || cand + DISTANCE_MAX >= cur) && rd32(src + cand) == rd32(src + ip)) {
                    tok_pos = op++; dst[tok_pos] = 0; match = cand;
                    continue;
                }
            }
            break;
        }
        fwd_h = hash_at(src + (++ip), tab16);                                  /* :1199 */
    }

last_literals:
    {   /* :1204-1231 */
        uint32_t last = (uint32_t)n - anchor;
        if (limited && (int64_t)op + last + 1 + (last + 255 - RUN_MASK) / 255 > cap) return 0;
        if (last >= RUN_MASK) {
            uint32_t acc = last - RUN_MASK;
            dst[op++] = RUN_MASK << 4;
            for (; acc >= 255; acc -= 255) dst[op++] = 255;
            dst[op++] = (uint8_t)acc;
        } else dst[op++] = (uint8_t)(last << 4);
        memcpy(dst + op, src + anchor, last);
        op += last;
    }
    return (int)op;
#undef TGET
#undef TPUT
}

/* Shared sequence walker: lz4/lz4.c:1929-2151 (the careful loop; the fast loop and the
 * shortcut are buffer-slack optimisations that accept the same well-formed streams).
 * dst == NULL => validate only.  Deviation, documented: offset 0 is rejected (the
 * reference accepts it and emits indeterminate bytes). */
static int64_t walk(const uint8_t* src, int64_t clen, uint8_t* dst, int64_t cap,
                    oracle_seq_stats* st, int* rule)
{
    int64_t ip = 0, op = 0;
    uint64_t nseq = 0;
    *rule = 0;
    if (cap == 0) { if (clen == 1 && src[0] == 0) return 0; *rule = 1; return -1; }  /* :1776 */
    if (clen <= 0) { *rule = 1; return -1; }                                        /* :1782 */
    for (;;) {
        if (ip >= clen) { *rule = 2; return -ip - 1; }
        uint32_t token = src[ip++];
        int64_t len = token >> 4;
        if (len == RUN_MASK) {                              /* :1976-1982, read_variable_length :1707 */
            if (ip >= clen - RUN_MASK) { *rule = 3; return -ip - 1; }
            uint32_t s;
            do {
                s = src[ip++]; len += s;
                if (s == 255 && ip >= clen - RUN_MASK) { *rule = 3; return -ip - 1; }
            } while (s == 255);
        }
        int64_t cpy = op + len;
        nseq++;
        if (cpy > cap - MFLIMIT || ip + len > clen - (2 + 1 + LASTLITERALS)) {       /* :1990 */
            if (ip + len != clen || cpy > cap) { *rule = 4; return -ip - 1; }      /* :2027-2036 */
            if (dst) memmove(dst + op, src + ip, (size_t)len);
            if (st) st->lit_bytes += (uint64_t)len;
            ip += len; op += len;
            break;
        }
        if (dst) memcpy(dst + op, src + ip, (size_t)len);
        if (st) { st->lit_bytes += (uint64_t)len; if (len == 0) st->zero_lit_seq++; }
        ip += len; op = cpy;

        int64_t off = src[ip] | (src[ip + 1] << 8); ip += 2;                        /* :2055 */
        int64_t ml = token & ML_MASK;
        if (ml == ML_MASK) {                                                        /* :2062-2067 */
            uint32_t s;
            do {
                s = src[ip++]; ml += s;
                if (ip >= clen - LASTLITERALS + 1) { *rule = 5; return -ip - 1; }
            } while (s == 255);
        }
        ml += MINMATCH;
        if (off > op || off == 0) { *rule = 6; return -ip - 1; }                    /* :2073 */
        cpy = op + ml;
        if (cpy > cap - LASTLITERALS) { *rule = 7; return -ip - 1; }                /* :2139 */
        if (dst) { for (int64_t i = 0; i < ml; i++) dst[op + i] = dst[op + i - off]; }
        if (st) {
            st->match_bytes += (uint64_t)ml;
            if (off < ml) st->overlap_seq++;
            if (off < 16) st->short_off_seq++;
        }
        op = cpy;
    }
    if (st) { st->n_seq += nseq; if (nseq > st->max_seq_in_block) st->max_seq_in_block = nseq; }
    return op;
}

int oracle_lz4_decompress_safe(const uint8_t* src, int clen, uint8_t* dst, int cap)
{
    int rule;
    if (!src) return -1;
    return (int)walk(src, clen, dst, cap, NULL, &rule);
}

int oracle_lz4_validate(const uint8_t* src, int clen, int n, oracle_seq_stats* stats)
{
    int rule;
    int64_t r = walk(src, clen, NULL, n, stats, &rule);
    if (r < 0) return rule ? rule : 99;
    if (r != n) return 8;
    return 0;
}

size_t oracle_lzbench_compress_bound(size_t insize) { return insize + insize / 6 + 16 * 1024; }

/* _lzbench/lzbench.cpp:266-298 */
int64_t oracle_lzbench_compress(oracle_codec_fn compress, const size_t* chunk_sizes, size_t n_chunks,
                                size_t* compr_sizes, uint8_t* inbuf, uint8_t* outbuf, size_t outsize)
{
    size_t sum = 0;
    for (size_t i = 0; i < n_chunks; i++) {
        size_t part = chunk_sizes[i];
        size_t outpart = oracle_lzbench_compress_bound(part);
        if (outpart > outsize) outpart = outsize;
        int64_t clen = compress((char*)inbuf, part, (char*)outbuf, outpart, 0, 0, NULL);
        if (clen <= 0 || (size_t)clen == part) {          /* :284-289 raw store */
            if (part > outsize) return 0;
            memcpy(outbuf, inbuf, part);
            clen = (int64_t)part;
        }
        inbuf += part; outbuf += clen; outsize -= (size_t)clen;
        compr_sizes[i] = (size_t)clen;
        sum += (size_t)clen;
    }
    return (int64_t)sum;
}

/* _lzbench/lzbench.cpp:301-329 */
int64_t oracle_lzbench_decompress(oracle_codec_fn decompress, const size_t* chunk_sizes,
                                  const size_t* compr_sizes, size_t n_chunks,
                                  uint8_t* inbuf, uint8_t* outbuf)
{
    size_t sum = 0;
    for (size_t i = 0; i < n_chunks; i++) {
        size_t part = compr_sizes[i];
        int64_t dlen;
        if (part == chunk_sizes[i]) { memcpy(outbuf, inbuf, part); dlen = (int64_t)part; }  /* :311-315 */
        else dlen = decompress((char*)inbuf, part, (char*)outbuf, chunk_sizes[i], 0, 0, NULL);
        if (dlen <= 0) return dlen;
        inbuf += part; outbuf += dlen; sum += (size_t)dlen;
    }
    return (int64_t)sum;
}

/* _lzbench/compressors.cpp:343-346 */
int64_t oracle_plugin_lz4_compress(char* in, size_t insize, char* out, size_t outsize,
                                   size_t level, size_t p2, char* wm)
{
    (void)level; (void)p2; (void)wm;
    return oracle_lz4_compress((const uint8_t*)in, (int)insize, (uint8_t*)out,
                               outsize > 0x7FFFFFFF ? 0x7FFFFFFF : (int)outsize);
}

/* _lzbench/compressors.cpp:358-362.  The reference calls LZ4_decompress_fast (trusts the
 * stream, ignores insize) and returns outsize; the restatement uses the bounds-checked
 * walker and reports failure instead of reading out of bounds. */
int64_t oracle_plugin_lz4_decompress(char* in, size_t insize, char* out, size_t outsize,
                                     size_t level, size_t p2, char* wm)
{
    (void)level; (void)p2; (void)wm;
    int r = oracle_lz4_decompress_safe((const uint8_t*)in, (int)insize, (uint8_t*)out, (int)outsize);
    return r == (int)outsize ? (int64_t)outsize : -1;
}

/* ---- multi-threaded CPU baseline driver (bench only) ---------------------------------- */
typedef struct {
    oracle_codec_fn codec; int direction;
    const size_t* chunk_sizes; size_t* compr_sizes; size_t lo, hi;
    uint8_t* plain; uint8_t* packed; size_t slot_stride; size_t plain_off;
    int64_t ns; int failed;
} mt_job;

static int64_t now_ns(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (int64_t)ts.tv_sec * 1000000000LL + ts.tv_nsec;
}

static void* mt_worker(void* arg)
{
    mt_job* j = (mt_job*)arg;
    int64_t t0 = now_ns();
    uint8_t* p = j->plain + j->plain_off;
    for (size_t i = j->lo; i < j->hi; i++) {
        size_t part = j->chunk_sizes[i];
        uint8_t* slot = j->packed + i * j->slot_stride;
        if (j->direction == 0) {
            int64_t c = j->codec((char*)p, part, (char*)slot, j->slot_stride, 0, 0, NULL);
            if (c <= 0 || (size_t)c == part) { memcpy(slot, p, part); c = (int64_t)part; }
            j->compr_sizes[i] = (size_t)c;
        } else {
            size_t c = j->compr_sizes[i];
            if (c == part) memcpy(p, slot, part);
            else if (j->codec((char*)slot, c, (char*)p, part, 0, 0, NULL) <= 0) j->failed = 1;
        }
        p += part;
    }
    j->ns = now_ns() - t0;
    return NULL;
}

int64_t oracle_mt_run(oracle_codec_fn codec, int direction, int n_threads,
                      const size_t* chunk_sizes, size_t* compr_sizes, size_t n_chunks,
                      uint8_t* plain, uint8_t* packed, size_t slot_stride)
{
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_chunks && n_chunks > 0) n_threads = (int)n_chunks;
    mt_job* jobs = (mt_job*)calloc((size_t)n_threads, sizeof(mt_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    size_t off = 0, lo = 0;
    for (int t = 0; t < n_threads; t++) {
        size_t hi = n_chunks * (size_t)(t + 1) / (size_t)n_threads;
        jobs[t] = (mt_job){ codec, direction, chunk_sizes, compr_sizes, lo, hi, plain, packed, slot_stride, off, 0, 0 };
        for (size_t i = lo; i < hi; i++) off += chunk_sizes[i];
        lo = hi;
    }
    int64_t t0 = now_ns();
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    int failed = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); failed |= jobs[t].failed; }
    int64_t wall = now_ns() - t0;
    free(jobs); free(th);
    return failed ? -1 : wall;
}
