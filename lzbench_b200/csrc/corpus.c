/*
 * corpus.c -- deterministic synthetic corpora for the LZ4 chunked path.
 *
 * The reference ships no data generator (lz4's datagen is not vendored), so the six named
 * inputs are defined here.  Every corpus is a concatenation of independent 64 KiB
 * generation blocks: block b of corpus (kind, seed) depends only on (kind, seed, b), so any
 * byte range can be produced by any number of threads (or ranks) and is bit-reproducible.
 * PRNG = splitmix64.
 *
 *   kind 0 "mixed"   50 % Zipf-like text (4096-word vocabulary plus 1024 stock phrases), 25 % 16-byte
 *                    binary records, 15 % uniform random bytes, 10 % single-byte runs
 *   kind 1 "random"  uniform random bytes (incompressible)
 *   kind 2 "repeat"  one 64-byte pattern repeated (offset-64 matches)
 *   kind 3 "zero"    all zero bytes (offset-1 RLE)
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GEN_BLOCK 65536u
#define VOCAB 4096

static inline uint64_t splitmix64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static char vocab_txt[VOCAB * 16];
static uint8_t vocab_len[VOCAB];
static uint16_t vocab_off[VOCAB];
#define PHRASES 1024
static uint16_t phrase_words[PHRASES][6];
static uint8_t phrase_n[PHRASES];
static pthread_once_t vocab_once = PTHREAD_ONCE_INIT;

/* cubic skew towards low ranks, integer-only: floor(n * u1*u2*u3) with u in [0,1) as 16-bit fractions */
static inline uint32_t zipf3(uint64_t q, uint32_t n)
{
    uint64_t u1 = q & 0xFFFF, u2 = (q >> 16) & 0xFFFF, u3 = (q >> 32) & 0xFFFF;
    return (uint32_t)((((u1 * u2) >> 16) * u3 * n) >> 32);
}

static void build_vocab(void)
{
    /* letters repeated roughly by English frequency */
    static const char alpha[] =
        "eeeeeeeeeeeetttttttttaaaaaaaaooooooooiiiiiiinnnnnnnssssssrrrrrrhhhhhhddddlllluuucccmmmffyywwggppbbvkxqjz";
    uint64_t s = 0x4C5A34423230304BULL;   /* fixed: the vocabulary does not depend on the corpus seed */
    uint32_t off = 0;
    for (int w = 0; w < VOCAB; w++) {
        uint64_t r = splitmix64(&s);
        /* common (low-rank) words are short, rare words longer: 2..5 for the first 64, up to 12 later */
        uint32_t maxlen = w < 64 ? 4 : (w < 512 ? 7 : 11);
        uint32_t len = 2 + (uint32_t)(r % maxlen);
        vocab_off[w] = (uint16_t)off;
        vocab_len[w] = (uint8_t)len;
        for (uint32_t i = 0; i < len; i++) {
            r = splitmix64(&s);
            vocab_txt[off++] = alpha[r % (sizeof(alpha) - 1)];
        }
    }
    for (int p = 0; p < PHRASES; p++) {
        uint64_t r = splitmix64(&s);
        phrase_n[p] = (uint8_t)(2 + r % 5);
        for (int k = 0; k < phrase_n[p]; k++) phrase_words[p][k] = (uint16_t)zipf3(splitmix64(&s), VOCAB);
    }
}

static void gen_mixed_block(uint64_t seed, uint64_t b, uint8_t* out /* GEN_BLOCK bytes */)
{
    uint64_t s = seed * 0xD1342543DE82EF95ULL + b * 0x2545F4914F6CDD1DULL + 0x1234567ULL;
    uint32_t pos = 0;
    (void)splitmix64(&s);
    while (pos < GEN_BLOCK) {
        uint64_t r = splitmix64(&s);
        uint32_t kind = (uint32_t)(r % 100);
        uint32_t room = GEN_BLOCK - pos;
        if (kind < 50) {                                   /* text */
            uint32_t len = 200 + (uint32_t)((r >> 8) % 3801);
            if (len > room) len = room;
            uint32_t end = pos + len, col = 0;
            uint32_t ph = 0, ph_left = 0;
            while (pos < end) {
                uint64_t q = splitmix64(&s);
                uint32_t rank;
                if (ph_left == 0 && ((q >> 59) < 16)) {          /* 1/2: start a stock phrase */
                    ph = zipf3(splitmix64(&s), PHRASES); ph_left = phrase_n[ph];
                }
                if (ph_left) { rank = phrase_words[ph][phrase_n[ph] - ph_left]; ph_left--; q |= 0xFFFFull << 48; }
                else rank = zipf3(q, VOCAB);
                const char* w = vocab_txt + vocab_off[rank];
                uint32_t wl = vocab_len[rank];
                uint32_t flags = (uint32_t)(q >> 48);
                for (uint32_t i = 0; i < wl && pos < end; i++) {
                    char c = w[i];
                    if (i == 0 && (flags & 0x1F) == 0) c = (char)(c - 32);   /* occasional capital */
                    out[pos++] = (uint8_t)c;
                }
                col += wl + 1;
                if (pos < end) {
                    if ((flags & 0x3E0) == 0) { out[pos++] = (flags & 0x400) ? ',' : '.'; if (pos < end) out[pos++] = ' '; }
                    else if (col > 72) { out[pos++] = '\n'; col = 0; }
                    else out[pos++] = ' ';
                }
            }
        } else if (kind < 75) {                            /* 16-byte binary records */
            uint32_t nrec = 16 + (uint32_t)((r >> 8) % 241);
            uint32_t counter = (uint32_t)(r >> 24);
            uint32_t tag = (uint32_t)splitmix64(&s);
            for (uint32_t k = 0; k < nrec && pos < GEN_BLOCK; k++, counter++) {
                uint8_t rec[16];
                uint32_t rnd = (uint32_t)splitmix64(&s) & 0xFFFFu;   /* 2 random + 2 zero bytes */
                memcpy(rec, &counter, 4); memcpy(rec + 4, &tag, 4); memcpy(rec + 8, &rnd, 4);
                memset(rec + 12, 0xFF, 4);
                uint32_t n = GEN_BLOCK - pos < 16 ? GEN_BLOCK - pos : 16;
                memcpy(out + pos, rec, n); pos += n;
            }
        } else if (kind < 90) {                            /* uniform random */
            uint32_t len = 256 + (uint32_t)((r >> 8) % 7937);
            if (len > room) len = room;
            uint32_t end = pos + len;
            while (pos < end) {
                uint64_t q = splitmix64(&s);
                uint32_t n = end - pos < 8 ? end - pos : 8;
                memcpy(out + pos, &q, n); pos += n;
            }
        } else {                                           /* single-byte run */
            uint32_t len = 64 + (uint32_t)((r >> 8) % 4033);
            if (len > room) len = room;
            memset(out + pos, (int)((r >> 40) & 0xFF), len); pos += len;
        }
    }
}

static void gen_block(int kind, uint64_t seed, uint64_t b, uint8_t* out)
{
    switch (kind) {
    case 0: gen_mixed_block(seed, b, out); break;
    case 1: {
        uint64_t s = seed * 0x9E3779B97F4A7C15ULL + b * 0xC2B2AE3D27D4EB4FULL + 7;
        for (uint32_t i = 0; i < GEN_BLOCK; i += 8) { uint64_t q = splitmix64(&s); memcpy(out + i, &q, 8); }
        break;
    }
    case 2: {
        uint8_t pat[64]; uint64_t s = seed + 0x51ED27;
        for (int i = 0; i < 64; i += 8) { uint64_t q = splitmix64(&s); memcpy(pat + i, &q, 8); }
        for (uint32_t i = 0; i < GEN_BLOCK; i += 64) memcpy(out + i, pat, 64);
        break;
    }
    default: memset(out, 0, GEN_BLOCK); break;
    }
}

typedef struct { int kind; uint64_t seed; uint64_t b0, b1; uint8_t* out; uint64_t first_block; uint64_t len; } job_t;

static void* worker(void* a)
{
    job_t* j = (job_t*)a;
    uint8_t tmp[GEN_BLOCK];
    for (uint64_t b = j->b0; b < j->b1; b++) {
        uint64_t o = (b - j->first_block) * GEN_BLOCK;
        uint64_t n = j->len - o < GEN_BLOCK ? j->len - o : GEN_BLOCK;
        if (n == GEN_BLOCK) gen_block(j->kind, j->seed, b, j->out + o);
        else { gen_block(j->kind, j->seed, b, tmp); memcpy(j->out + o, tmp, n); }
    }
    return NULL;
}

/* Fill out[0..len) with bytes [offset, offset+len) of corpus (kind, seed).
 * offset must be a multiple of 64 KiB.  Returns 0, or -1 on bad arguments. */
int lz4b200_corpus_fill(int kind, uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out, int n_threads)
{
    if (kind < 0 || kind > 3 || (offset % GEN_BLOCK) != 0 || !out) return -1;
    pthread_once(&vocab_once, build_vocab);
    uint64_t first = offset / GEN_BLOCK, nblk = (len + GEN_BLOCK - 1) / GEN_BLOCK;
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > nblk) n_threads = (int)(nblk ? nblk : 1);
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof *th);
    job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof *jobs);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (job_t){ kind, seed, first + nblk * (uint64_t)t / (uint64_t)n_threads,
                           first + nblk * (uint64_t)(t + 1) / (uint64_t)n_threads, out, first, len };
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
