// decode_sim.c -- CPU model of the lane-per-sequence decoder (design study for kernels_v7):
// how many dependency rounds does a group of 32 consecutive sequences need, and how long are the
// per-round copy loops?  Input: the bench corpus compressed by the oracle (test infrastructure).
//   gcc -O2 -o /tmp/decode_sim scripts/sim/decode_sim.c lzbench_b200/csrc/corpus.c oracle/lz4_oracle.c -lpthread
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int lz4b200_corpus_fill(int kind, uint64_t seed, uint64_t offset, uint64_t len, uint8_t* out, int n_threads);
int oracle_lz4_compress(const uint8_t* src, int n, uint8_t* dst, int cap);
typedef struct { uint32_t L, off, M, op, ip; } Seq;
#define G 32
int main(int argc, char** argv)
{
    int kind = argc > 1 ? atoi(argv[1]) : 0;
    size_t n = (argc > 2 ? atoi(argv[2]) : 64) << 20, chunk = argc > 3 ? atoi(argv[3]) : 65536;
    int CAP = argc > 4 ? atoi(argv[4]) : 32; int FWD = argc > 5 ? atoi(argv[5]) : 0; double tot_fsteps = 0;
    uint8_t* data = malloc(n); uint8_t* comp = malloc(chunk * 2);
    lz4b200_corpus_fill(kind, 2, 0, n, data, 8);
    Seq* s = malloc(sizeof(Seq) * chunk);
    double tot_seq = 0, tot_groups = 0, tot_rounds_exact = 0, tot_rounds_front = 0, tot_lit_iter = 0, tot_match_iter = 0;
    double far = 0, nmatch = 0, longlit = 0, longmatch = 0, overlap = 0, tot_c = 0, intra = 0;
    uint64_t hist[G + 1] = {0};
    double tot_depth = 0; int max_depth = 0;
    static uint16_t lvl[70000];
    for (size_t c0 = 0; c0 < n; c0 += chunk) {
        int cl = oracle_lz4_compress(data + c0, chunk, comp, chunk * 2);
        tot_c += cl;
        uint32_t ip = 0, op = 0, ns = 0;
        while (1) {
            uint32_t t = comp[ip], st = ip; ip++;
            uint32_t L = t >> 4; if (L == 15) { uint32_t b; do { b = comp[ip++]; L += b; } while (b == 255); }
            s[ns].L = L; s[ns].op = op; s[ns].ip = st; ip += L; op += L;
            if (ip >= (uint32_t)cl) { s[ns].M = 0; s[ns].off = 0; ns++; break; }
            uint32_t off = comp[ip] | (comp[ip + 1] << 8); ip += 2;
            uint32_t M = t & 15; if (M == 15) { uint32_t b; do { b = comp[ip++]; M += b; } while (b == 255); }
            M += 4; s[ns].M = M; s[ns].off = off; op += M; ns++;
        }
        tot_seq += ns;
        // exact byte-level DAG depth over the chunk (levels per byte)
        {
            int md = 0;
            for (uint32_t i = 0; i < ns; i++) {
                for (uint32_t k = 0; k < s[i].L; k++) lvl[s[i].op + k] = 0;
                uint32_t d = s[i].op + s[i].L; int l = 0;
                uint32_t span = s[i].M < s[i].off ? s[i].M : s[i].off;
                for (uint32_t k = 0; k < span; k++) if (lvl[d - s[i].off + k] > l) l = lvl[d - s[i].off + k];
                for (uint32_t k = 0; k < s[i].M; k++) lvl[d + k] = l + 1;
                if (s[i].M && l + 1 > md) md = l + 1;
            }
            tot_depth += md; if (md > max_depth) max_depth = md;
        }
        for (uint32_t g0 = 0; g0 < ns; g0 += G) {
            uint32_t ng = ns - g0 < G ? ns - g0 : G;
            tot_groups++;
            uint32_t base = s[g0].op;
            // literal loop iterations: max over lanes of min(L, CAP); long literals counted
            uint32_t ml = 0;
            for (uint32_t i = 0; i < ng; i++) { uint32_t L = s[g0 + i].L; if (L > (uint32_t)CAP) { longlit++; L = 0; } if (L > ml) ml = L; }
            tot_lit_iter += ml;
            // source forwarding (pointer jumping): while the whole source of match i lies inside the match region of ONE
            // in-group lane j, read from j's (already forwarded) source instead: shift_i += shift_j
            uint32_t shift[G]; int fsteps = 0;
            for (uint32_t i = 0; i < ng; i++) shift[i] = s[g0 + i].off;
            if (FWD) for (int it = 0; it < 6; it++) {
                uint32_t ns2[G]; int changed = 0;
                for (uint32_t i = 0; i < ng; i++) {
                    ns2[i] = shift[i]; Seq* a = &s[g0 + i]; if (!a->M) continue;
                    uint32_t dst = a->op + a->L, src = dst - shift[i]; uint32_t span = a->M < a->off ? a->M : a->off;
                    if (a->off < a->M) continue;            // overlapping matches are not forwarded
                    for (uint32_t j = 0; j < i; j++) {
                        Seq* b = &s[g0 + j]; if (!b->M || b->off < b->M) continue;
                        uint32_t bd = b->op + b->L;
                        if (src >= bd && src + span <= bd + b->M) { ns2[i] = shift[i] + shift[j]; changed = 1; break; }
                    }
                }
                if (!changed) break;
                memcpy(shift, ns2, sizeof shift); fsteps++;
            }
            tot_fsteps += fsteps;
            // rounds (exact intra-group): level[i] = 1 + max level of in-group matches whose region intersects source
            int level[G]; int done_round[G];
            int R = 0;
            for (uint32_t i = 0; i < ng; i++) {
                Seq* a = &s[g0 + i]; level[i] = 0;
                if (!a->M) continue;
                nmatch++;
                uint32_t dst = a->op + a->L, src = dst - shift[i];
                uint32_t span = a->M < a->off ? a->M : a->off;
                if (a->off < a->M) overlap++;
                if (a->M > (uint32_t)CAP) longmatch++;
                if (src + span <= base) { if (base - src > 2048) far++; }
                int l = 0;
                for (uint32_t j = 0; j < i; j++) {
                    Seq* b = &s[g0 + j]; if (!b->M) continue;
                    uint32_t bd = b->op + b->L;
                    if (bd < src + span && bd + b->M > src && level[j] > l) l = level[j];
                }
                if (l) intra++;
                level[i] = l + 1; if (level[i] > R) R = level[i];
            }
            tot_rounds_exact += R; hist[R]++;
            // copy-loop iterations: per round, max over lanes in that round of min(M, CAP)
            for (int r = 1; r <= R; r++) {
                uint32_t mm = 0;
                for (uint32_t i = 0; i < ng; i++) if (level[i] == r) { uint32_t M = s[g0 + i].M; if (M > (uint32_t)CAP) M = 0; if (M > mm) mm = M; }
                tot_match_iter += mm;
            }
            // frontier rule: lane ready iff src+span <= dst of first not-done match lane (or it is that lane)
            {
                int pend[G]; int np = 0; for (uint32_t i = 0; i < ng; i++) { pend[i] = s[g0 + i].M != 0; np += pend[i]; }
                int rr = 0;
                while (np) {
                    uint32_t f = 0; while (!pend[f]) f++;
                    uint32_t fd = s[g0 + f].op + s[g0 + f].L;
                    int rdy[G];
                    for (uint32_t i = 0; i < ng; i++) {
                        rdy[i] = 0; if (!pend[i]) continue;
                        Seq* a = &s[g0 + i]; uint32_t dst = a->op + a->L, src = dst - a->off; uint32_t span = a->M < a->off ? a->M : a->off;
                        if (i == f || src + span <= fd) rdy[i] = 1;
                    }
                    for (uint32_t i = 0; i < ng; i++) if (rdy[i]) { pend[i] = 0; np--; }
                    rr++;
                }
                tot_rounds_front += rr;
            }
            (void)done_round;
        }
    }
    printf("chunks %zu  ratio %.2f%%  seq/chunk %.0f  groups/chunk %.1f\n", n / chunk, 100.0 * tot_c / n, tot_seq / (n / chunk), tot_groups / (n / chunk));
    printf("chunk DAG depth (byte-exact): avg %.1f max %d\n", tot_depth / (n / chunk), max_depth);
    printf("rounds/group exact %.2f  frontier %.2f\n", tot_rounds_exact / tot_groups, tot_rounds_front / tot_groups);
    printf("lit-loop iters/group %.1f  match-loop iters/group (sum over rounds) %.1f  (cap %d)\n", tot_lit_iter / tot_groups, tot_match_iter / tot_groups, CAP);
    printf("matches: far(>2KB before group) %.1f%%  intra-group dep %.1f%%  overlap %.2f%%  long(>cap) %.2f%%  long literals/seq %.2f%%\n",
           100 * far / nmatch, 100 * intra / nmatch, 100 * overlap / nmatch, 100 * longmatch / nmatch, 100 * longlit / tot_seq);
    printf("forwarding steps/group %.2f\n", tot_fsteps / tot_groups);
    printf("round histogram:"); for (int i = 0; i <= G; i++) if (hist[i]) printf(" %d:%.1f%%", i, 100.0 * hist[i] / tot_groups); printf("\n");
    return 0;
}
