// compress_sim.c -- CPU model of the c3 match finder (kernels_c3.cuh), used to choose its parameters BEFORE writing CUDA:
// what does the compressed size cost when the hash table is updated in groups of G positions (all G read, then all G
// write) instead of position by position, and which "near" distances (same hash d lanes earlier in the group) are needed
// to get the short-period data back?  Size only (sequence sizes are summed, no bytes are emitted); the reference size
// comes from the oracle (byte-identical to LZ4_compress_default, tests/test_oracle.py).
//
//   gcc -O2 -o /tmp/compress_sim scripts/sim/compress_sim.c lzbench_b200/csrc/corpus.c oracle/lz4_oracle.c -lpthread
//   /tmp/compress_sim <kind 0..3 | file> <MiB> <chunk>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int lz4b200_corpus_fill(int kind, uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out, int n_threads);
int oracle_lz4_compress(const uint8_t* src, int n, uint8_t* dst, int cap);

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hash13(uint32_t v) { return (v * 2654435761u) >> 19; }
static uint32_t ext_bytes(uint32_t x) { return x >= 15 ? 1 + (x - 15) / 255 : 0; }

typedef struct { int G; int nd; int d[8]; int winner_low; int runlen; } cfg_t;

// returns the compressed size of one chunk under the model
static uint32_t model(const uint8_t* in, uint32_t n, const cfg_t* c, uint16_t* cand, uint8_t* ok)
{
    if (n < 13) return 1 + ext_bytes(n) + n;
    static uint16_t tab[8192];
    memset(tab, 0, sizeof tab);
    const uint32_t mfl = n - 12, mlimit = n - 5;
    for (uint32_t g = 0; g <= mfl; g += c->G) {
        uint32_t h[256];
        const uint32_t m = (g + c->G <= mfl + 1) ? (uint32_t)c->G : mfl + 1 - g;
        for (uint32_t l = 0; l < m; l++) { h[l] = hash13(rd32(in + g + l)); cand[g + l] = tab[h[l]]; }
        for (uint32_t l = 0; l < m; l++)
            for (int k = 0; k < c->nd; k++) {                     // smallest distance first
                const uint32_t d = (uint32_t)c->d[k];
                if (l >= d && h[l] == h[l - d]) { cand[g + l] = (uint16_t)(g + l - d); break; }
            }
        if (c->winner_low) for (uint32_t l = m; l-- > 0;) tab[h[l]] = (uint16_t)(g + l);
        else for (uint32_t l = 0; l < m; l++) tab[h[l]] = (uint16_t)(g + l);
    }
    for (uint32_t p = 0; p <= mfl; p++) ok[p] = cand[p] < p && rd32(in + cand[p]) == rd32(in + p);
    uint32_t anchor = 0, pos = 0, out = 0;
    while (pos <= mfl) {
        if (!ok[pos]) { pos++; continue; }
        uint32_t p = pos, q = cand[p], len = 4;
        if (c->runlen) {                                          // length from the run of matched positions with the same offset
            uint32_t r = p;
            while (r + 1 <= mfl && ok[r + 1] && r + 1 - cand[r + 1] == p - q) r++;
            len = r - p + 4;
            if (p + len > mlimit) len = mlimit - p;
            if (c->runlen == 2) while (p + len < mlimit && in[p + len] == in[q + len]) len++;   // ... extended at its end (= exact)
            if (c->runlen == 3) { uint32_t e = 0; while (e < 12 && p + len < mlimit && in[p + len] == in[q + len]) { len++; e++; } }  // ... by at most 12 bytes
        } else
        while (p + len < mlimit && in[p + len] == in[q + len]) len++;
        while (p > anchor && q > 0 && in[p - 1] == in[q - 1]) { p--; q--; len++; }
        const uint32_t lit = p - anchor;
        out += 1 + ext_bytes(lit) + lit + 2 + ext_bytes(len - 4);
        anchor = pos = p + len;
    }
    const uint32_t last = n - anchor;
    return out + 1 + ext_bytes(last) + last;
}

int main(int argc, char** argv)
{
    const char* kind = argc > 1 ? argv[1] : "0";
    const uint64_t total = (uint64_t)(argc > 2 ? atoi(argv[2]) : 64) << 20;
    const uint32_t chunk = argc > 3 ? (uint32_t)atoi(argv[3]) : 65536;
    uint8_t* buf = malloc(total + 65536);
    uint64_t have = total;
    if (kind[0] >= '0' && kind[0] <= '3' && !kind[1]) lz4b200_corpus_fill(kind[0] - '0', 2, 0, total, buf, 8);
    else { FILE* f = fopen(kind, "rb"); have = fread(buf, 1, total, f); fclose(f); }
    const cfg_t cfgs[] = {
        { 1, 0, {0}, 0, 0 },
        { 32, 0, {0}, 0, 0 },
        { 32, 1, {1}, 0, 0 },
        { 32, 3, {1, 2, 4}, 0, 0 },
        { 32, 4, {1, 2, 3, 4}, 0, 0 },
        { 32, 6, {1, 2, 3, 4, 8, 16}, 0, 0 },
        { 32, 6, {1, 2, 3, 4, 8, 16}, 1, 0 },
        { 32, 1, {1}, 0, 1 },
        { 32, 1, {1}, 0, 2 },
        { 32, 1, {1}, 0, 3 },
        { 32, 0, {0}, 0, 1 },
    };
    const int ncfg = sizeof cfgs / sizeof cfgs[0];
    uint64_t ref = 0, sz[16] = { 0 };
    uint16_t* cand = malloc(2 * 65536); uint8_t* ok = malloc(65536); uint8_t* tmp = malloc(70000);
    for (uint64_t o = 0; o < have; o += chunk) {
        const uint32_t n = have - o < chunk ? (uint32_t)(have - o) : chunk;
        int r = oracle_lz4_compress(buf + o, (int)n, tmp, 70000);
        ref += (r <= 0 || (uint32_t)r >= n) ? n : (uint32_t)r;
        for (int k = 0; k < ncfg; k++) { uint32_t s = model(buf + o, n, &cfgs[k], cand, ok); sz[k] += s >= n ? n : s; }
    }
    printf("%s: %llu bytes, chunk %u, reference %.3f %%\n", kind, (unsigned long long)have, chunk, 100.0 * ref / have);
    for (int k = 0; k < ncfg; k++) {
        printf("  G=%3d near={", cfgs[k].G);
        for (int i = 0; i < cfgs[k].nd; i++) printf("%d%s", cfgs[k].d[i], i + 1 < cfgs[k].nd ? "," : "");
        printf("} runlen=%d%s: %.3f %%  x%.4f\n", cfgs[k].runlen, cfgs[k].winner_low ? " low-wins" : "", 100.0 * sz[k] / have, (double)sz[k] / ref);
    }
    return 0;
}
