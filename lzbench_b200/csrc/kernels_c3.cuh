// kernels_c3.cuh -- generation-3 compressor: the c2 algorithm taken apart once more, for instruction count.
//
// The c2 profile (profiles/r01_final_compress_c2_ncu_*): 6.3 G warp instructions per GiB (5.9 per input byte), IPC 2.25,
// every phase issue-bound.  Where they go: the match finder spends ~85 thread instructions per position (unaligned
// reads, a batch-local table kept with atomics, two verifications), and the parse walks every speculative window
// through the data (extension compares inside a divergent per-lane loop, 2.4 rounds per tile).  c3 keeps the result
// (LZ4_compress_default's greedy parse, lz4/lz4.c:851-1240, over an every-position hash table) and changes who does what:
//
//   H  hash, all positions, 4 consecutive positions per lane from two aligned words (one funnel shift each):
//      13-bit hash (lz4/lz4.c:698-702) -> u16 per position.
//   T  table, ONE warp per chunk, strictly in position order, 32 positions per step: candidate = table[h],
//      table[h] = position (lz4/lz4.c:736-787; last writer wins).  Nothing but shared-memory traffic: 4 instructions
//      per 32 positions and a shuffle for "same hash as the position before" (runs).  A step reads before it writes,
//      so a position misses a repeat that lies in its own step; the CPU model (scripts/sim/compress_sim.c) puts
//      that at 0.05 % of the output on the bench corpus and 0.3 % on real binaries (the parse's backward extension,
//      lz4/lz4.c:1019, gets most of it back).
//   V  verify, all positions, 4 per lane: 4 bytes at the candidate == 4 bytes here -> offset (u16, in place) and two
//      bit maps: `ok` (has a match) and `cont` (has a match with the same offset as the position before).
//   R  extension, once per RUN of positions with the same offset (not once per window walk): every position of a
//      run is the same match minus its first bytes, so they all END at the same place -- the run's last position + 4
//      + what a compare from there on adds.  Run ends are compacted inside a warp and extended densely, at most
//      C3_CAP bytes (longer ones are finished by a whole warp when the parse really takes them).
//   P  parse, one warp, bit maps only: the greedy chain "first matched position at or after the end of the previous
//      match; jump to its end" (lz4/lz4.c:956-1014) over 32 windows of 64 positions, entry states repaired by
//      prefix-max rounds as in c1; a step is a few bit operations and one byte load, no access to the data.
//   E  emit, one lane per sequence: backward extension (lz4/lz4.c:1019), sizes, scan, bytes (lz4/lz4.c:1022-1135).
//
// The T warp runs one tile ahead of the others (it needs only H of its tile), everything else is a sequence of
// short phases over a 2048-position tile.  Same block format rules as c1/c2 (end-of-block rules lz4/lz4.c:883-884,
// last literals :1204-1231); chunks above 64 KiB and the lz4fast levels stay on c2.
#pragma once
#include "common.cuh"
#include "kernels_c1.cuh"

namespace lz4b200 {

constexpr u32 C3_TILE = 2048;                    // positions per tile
constexpr u32 C3_SLICE = 128;                    // positions per warp step of H and V (4 per lane)
constexpr u32 C3_CAP = 60;                       // bytes a lane compares past a run end; the byte 255 stands for "more"
constexpr u32 C3_MAXSEQ = C3_TILE / MINMATCH;    // sequences that can start inside one tile
constexpr u32 C3_LONG_LIT = 64;                  // literal runs this long are copied by a whole warp
constexpr u32 C3_NLONGQ = C3_TILE / C3_LONG_LIT + 2;
constexpr u32 C3_NLONGM = 32;                    // long matches per tile whose end is remembered (a tile holds at most 32 of >= 64 bytes)

template <u32 NWORK>
struct C3Smem {
    static constexpr u32 OFF_IN = 0;
    static constexpr u32 IN_CAP = C1_CHUNK_MAX + 64;
    static constexpr u32 OFF_TAB = OFF_IN + IN_CAP;                     // u16[8192]
    static constexpr u32 OFF_HC = OFF_TAB + 2 * (1u << C1_HBITS);       // u16[3][TILE]: hash -> candidate -> offset, in place
    static constexpr u32 OFF_EXT = OFF_HC + 3 * 2 * C3_TILE;            // u8[TILE]: bytes a run end's match goes on past its 4 (255: >= C3_CAP)
    static constexpr u32 OFF_OKB = OFF_EXT + C3_TILE;                   // u32[TILE / 32]
    static constexpr u32 OFF_CONT = OFF_OKB + 4 * (C3_TILE / 32);       // u32[TILE / 32]
    static constexpr u32 OFF_SEQP = OFF_CONT + 4 * (C3_TILE / 32);      // u16[MAXSEQ] start of the sequence's match
    static constexpr u32 OFF_SEQE = OFF_SEQP + 2 * C3_MAXSEQ;           // u16[MAXSEQ] ... its end (65536 does not occur: matches end 5 bytes before the chunk does)
    static constexpr u32 OFF_SEQF = OFF_SEQE + 2 * C3_MAXSEQ;           // u16[MAXSEQ] ... its offset
    static constexpr u32 OFF_SEQQ = OFF_SEQF + 2 * C3_MAXSEQ;           // u16[MAXSEQ] ... its start after the backward extension
    static constexpr u32 OFF_SEQO = OFF_SEQQ + 2 * C3_MAXSEQ;           // u32[MAXSEQ] bytes the earlier sequences of its block of 32 take
    static constexpr u32 OFF_LIST = OFF_SEQO + 4 * C3_MAXSEQ;           // u8[NWORK][SLICE] run ends of a slice, per warp
    static constexpr u32 OFF_LONGQ = (OFF_LIST + NWORK * C3_SLICE + 15) & ~15u;   // uint4[NLONGQ] long literal runs
    static constexpr u32 OFF_LONGM = OFF_LONGQ + 16 * C3_NLONGQ;        // u16[2][NLONGM] long matches: run end, match end
    static constexpr u32 OFF_BLK = OFF_LONGM + 4 * C3_NLONGM;           // u32[MAXSEQ / 32] bytes per block of 32 sequences
    static constexpr u32 OFF_MISC = OFF_BLK + 4 * (C3_MAXSEQ / 32);     // mbarrier (8 B) + scalars
    static constexpr u32 BYTES = OFF_MISC + 64;
};

template <u32 COUNT> __device__ __forceinline__ void c3_bar_work() { asm volatile("bar.sync 1, %0;" ::"n"(COUNT) : "memory"); }

// last position of the run that position p (tile-relative index i) belongs to: p + the `cont` bits set from i + 1 on
__device__ __forceinline__ u32 c3_run_end(const u32* contb, u32 i)
{
    u32 j = i + 1;
    for (;;) {
        if (j >= C3_TILE) return C3_TILE - 1;
        const u32 sh = j & 31u;
        const u32 t = ~(contb[j >> 5] >> sh);              // the zeros shifted in at the top end the count at the word's last bit
        const u32 n1 = t ? (u32)__ffs(t) - 1u : 32u;
        j += n1;
        if (n1 + sh < 32u) return j - 1;
    }
}

template <u32 NWORK>
__global__ void __launch_bounds__((NWORK + 1) * 32, 2)
compress_c3_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ usz,
                   u8* __restrict__ stage, u64 slot_stride, u64* __restrict__ csz, u32 n_chunks, u32* ticket,
                   u64* __restrict__ prof)
{
    typedef C3Smem<NWORK> S;
    constexpr u32 NTHR = (NWORK + 1) * 32, NWT = NWORK * 32;
    extern __shared__ __align__(128) u8 smem[];
    u8* const s_raw = smem + S::OFF_IN;
    u16* const s_tab = reinterpret_cast<u16*>(smem + S::OFF_TAB);
    u16* const s_hc = reinterpret_cast<u16*>(smem + S::OFF_HC);
    u8* const s_ext = smem + S::OFF_EXT;
    u32* const s_okb = reinterpret_cast<u32*>(smem + S::OFF_OKB);
    u32* const s_cont = reinterpret_cast<u32*>(smem + S::OFF_CONT);
    u16* const s_seqp = reinterpret_cast<u16*>(smem + S::OFF_SEQP);
    u16* const s_seqe = reinterpret_cast<u16*>(smem + S::OFF_SEQE);
    u16* const s_seqf = reinterpret_cast<u16*>(smem + S::OFF_SEQF);
    u16* const s_seqq = reinterpret_cast<u16*>(smem + S::OFF_SEQQ);
    u32* const s_seqo = reinterpret_cast<u32*>(smem + S::OFF_SEQO);
    u8* const s_list = smem + S::OFF_LIST;
    uint4* const s_longq = reinterpret_cast<uint4*>(smem + S::OFF_LONGQ);
    u16* const s_longm = reinterpret_cast<u16*>(smem + S::OFF_LONGM);
    u32* const s_blk = reinterpret_cast<u32*>(smem + S::OFF_BLK);
    u64* const s_bar = reinterpret_cast<u64*>(smem + S::OFF_MISC);
    u32* const s_misc = reinterpret_cast<u32*>(smem + S::OFF_MISC + 16);   // [0] chunk [1] long-literal count [2] long-match count [3] sequences of the tile [4] chain state after the tile [5] cur, [6] op for the last literals

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_t = warp == NWORK;                  // the table warp
    u32 phase = 0;
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();

    for (;;) {
        if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 chunk = s_misc[0];
        if (chunk >= n_chunks) break;
        const u8* src = in + in_off[chunk];
        const u32 n = (u32)usz[chunk];
        u8* dst = stage + (u64)chunk * slot_stride;
        const long long pt0 = prof ? clock64() : 0;

        // ---- stage the chunk with position 0 on a 4-byte boundary, clear the table ----
        const u32 head = (u32)(reinterpret_cast<uintptr_t>(src) & 15);
        const bool tma = (head & 3u) == 0;
        u8* const s_in = tma ? s_raw + head : s_raw;
        if (tma) {                                    // TMA bulk copy of the 16-byte aligned interior, edges by hand (as c1)
            const u32 lead = head ? 16 - head : 0;
            const u32 lead_n = lead < n ? lead : n;
            const u32 bulk = (n - lead_n) & ~15u;
            if (tid == 0 && bulk) {
                fence_async_smem();
                mbar_expect_tx(s_bar, bulk);
                bulk_g2s(s_in + lead_n, src + lead_n, bulk, s_bar);
            }
            if (tid < lead_n) s_in[tid] = src[tid];
            const u32 tail0 = lead_n + bulk;
            if (tid >= 32 && tid - 32 < n - tail0) s_in[tail0 + tid - 32] = src[tail0 + tid - 32];
            uint4* t4 = reinterpret_cast<uint4*>(s_tab);
            for (u32 i = tid; i < (2u << C1_HBITS) / 16; i += NTHR) t4[i] = make_uint4(0, 0, 0, 0);
            if (bulk) { mbar_wait(s_bar, phase); phase ^= 1; }
        } else {                                      // source not on a 4-byte boundary: realigning word copy
            const u32 sh = (head & 3u) * 8u;
            const u32* sw = reinterpret_cast<const u32*>(src - (head & 3u));
            const u32 nw = n >= 8 ? n / 4 - 1 : 0;   // words whose two source words lie inside the chunk
            u32* dw = reinterpret_cast<u32*>(s_in);
            for (u32 k = tid; k < nw; k += NTHR) dw[k] = __funnelshift_r(__ldg(sw + k), __ldg(sw + k + 1), sh);
            for (u32 i = nw * 4 + tid; i < n; i += NTHR) s_in[i] = src[i];
            uint4* t4 = reinterpret_cast<uint4*>(s_tab);
            for (u32 i = tid; i < (2u << C1_HBITS) / 16; i += NTHR) t4[i] = make_uint4(0, 0, 0, 0);
        }
        if (tid == 0) { s_misc[1] = 0; s_misc[5] = 0; s_misc[6] = 0; }
        __syncthreads();
        if (prof && tid == 0) atomicAdd((unsigned long long*)&prof[5], (unsigned long long)(clock64() - pt0));

        if (n >= MIN_LENGTH) {
            const u32 mfl = n - MFLIMIT;              // last position a match may start at
            const u32 mlimit = n - LASTLITERALS;      // matches end at or before this
            const u32 n_tiles = mfl / C3_TILE + 1;
            const u32* const s_in32 = reinterpret_cast<const u32*>(s_in);

            // H of one tile: hash of every position, 4 consecutive positions per lane
            auto hash_tile = [&](u32 t, u32 w0, u32 wstep) {
                u16* const hc = s_hc + (t % 3) * C3_TILE;
                for (u32 s = w0; s < C3_TILE / C3_SLICE; s += wstep) {
                    const u32 i0 = s * C3_SLICE + 4 * lane;
                    const u32* w = s_in32 + ((t * C3_TILE + i0) >> 2);
                    const u32 a = w[0], b = w[1];
                    const u32 h0 = hash13(a), h1 = hash13(__funnelshift_r(a, b, 8));
                    const u32 h2 = hash13(__funnelshift_r(a, b, 16)), h3 = hash13(__funnelshift_r(a, b, 24));
                    *reinterpret_cast<uint2*>(hc + i0) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
                }
            };
            // T of one tile: candidate = table[h], table[h] = position, 32 positions per step in position order
            auto table_tile = [&](u32 t) {
                u16* const hc = s_hc + (t % 3) * C3_TILE;
                const u32 t0 = t * C3_TILE;
                #pragma unroll 4
                for (u32 s = 0; s < C3_TILE / 32; s++) {
                    const u32 i = s * 32 + lane;
                    const u32 h = hc[i];
                    const u32 hp = __shfl_up_sync(FULL, h, 1);
                    u32 c = s_tab[h];
                    s_tab[h] = (u16)(t0 + i);
                    if (lane && hp == h) c = t0 + i - 1;            // the table is a step behind: a run would miss its own first step
                    hc[i] = (u16)c;
                    __syncwarp();                                   // the next step reads what this one wrote
                }
            };

            if (!is_t) { hash_tile(0, warp, NWORK); if (n_tiles > 1) hash_tile(1, warp, NWORK); }
            __syncthreads();
            if (is_t) table_tile(0);
            __syncthreads();

            u32 cur = 0;                              // chain state: end of the last match (= literal anchor)
            u32 op = 0;                               // bytes emitted so far
            for (u32 t = 0; t < n_tiles; t++) {
                const u32 t0 = t * C3_TILE;
                if (is_t) {
                    if (t + 1 < n_tiles) table_tile(t + 1);
                    __syncthreads();
                    continue;
                }
                u16* const hc = s_hc + (t % 3) * C3_TILE;
                const bool skip = cur >= t0 + C3_TILE;              // the tile lies inside a match: nothing to find in it
                const long long q0 = prof ? clock64() : 0;
                // ---- long literal runs of the previous tile, whole warps ----
                {
                    const u32 nlong = s_misc[1];
                    for (u32 e = warp; e < nlong; e += NWORK) {
                        const uint4 q = s_longq[e];
                        for (u32 i = lane; i < q.z; i += 32) dst[q.y + i] = s_in[q.x + i];
                    }
                }
                // ---- V + R ----
                if (!skip) {
                    u8* const list = s_list + warp * C3_SLICE;
                    for (u32 s = warp; s < C3_TILE / C3_SLICE; s += NWORK) {
                        const u32 i0 = s * C3_SLICE + 4 * lane, p0 = t0 + i0;
                        const u32* w = s_in32 + (p0 >> 2);
                        const u32 a = w[0], b = w[1];
                        const uint2 cc = *reinterpret_cast<const uint2*>(hc + i0);
                        u32 v[4] = { a, __funnelshift_r(a, b, 8), __funnelshift_r(a, b, 16), __funnelshift_r(a, b, 24) };
                        u32 c[4] = { cc.x & 0xFFFFu, cc.x >> 16, cc.y & 0xFFFFu, cc.y >> 16 };
                        u32 off[4];
                        #pragma unroll
                        for (u32 k = 0; k < 4; k++) {
                            const bool ok = p0 + k <= mfl && c[k] < p0 + k && c1_rd32(s_in + c[k]) == v[k];
                            off[k] = ok ? p0 + k - c[k] : 0u;
                        }
                        *reinterpret_cast<uint2*>(hc + i0) = make_uint2(off[0] | (off[1] << 16), off[2] | (off[3] << 16));
                        u32 prev = __shfl_up_sync(FULL, off[3], 1);
                        if (lane == 0) prev = 0;                    // a slice starts a new run
                        u32 okn = 0, cn = 0;
                        #pragma unroll
                        for (u32 k = 0; k < 4; k++) {
                            okn |= (u32)(off[k] != 0) << k;
                            cn |= (u32)(off[k] != 0 && off[k] == (k ? off[k - 1] : prev)) << k;
                        }
                        u32 nextc = __shfl_down_sync(FULL, cn & 1u, 1);
                        if (lane == 31) nextc = 0;
                        const u32 ren = okn & ~((cn >> 1) | (nextc << 3));   // run ends: matched, and the next position does not continue
                        // bit maps: lanes 8k .. 8k+7 hold word k of the slice
                        u32 okw = okn << (4 * (lane & 7)), cw = cn << (4 * (lane & 7));
                        #pragma unroll
                        for (u32 d = 1; d < 8; d <<= 1) { okw |= __shfl_xor_sync(FULL, okw, d); cw |= __shfl_xor_sync(FULL, cw, d); }
                        if ((lane & 7) == 0) { s_okb[(i0 >> 5)] = okw; s_cont[(i0 >> 5)] = cw; }
                        // run ends, compacted (any order will do) and extended densely
                        u32 total = 0;
                        #pragma unroll
                        for (u32 k = 0; k < 4; k++) {
                            const unsigned bk = __ballot_sync(FULL, (ren >> k) & 1u);
                            if ((ren >> k) & 1u) list[total + __popc(bk & ((1u << lane) - 1u))] = (u8)(4 * lane + k);
                            total += __popc(bk);
                        }
                        __syncwarp();
                        for (u32 k = lane; k < total; k += 32) {
                            const u32 j = s * C3_SLICE + list[k];
                            const u32 o = hc[j], a4 = t0 + j + MINMATCH;
                            const u32 x = c1_count(s_in, a4, a4 - o, mlimit, C3_CAP);
                            s_ext[j] = (u8)(x >= C3_CAP ? 255u : x);
                        }
                        __syncwarp();
                    }
                }
                const long long q1 = prof ? clock64() : 0;
                c3_bar_work<NWT>();
                // ---- P (warp 0) beside H of tile t + 2 (the other warps) ----
                if (warp == 0) {
                    u32 cur_out = cur, nseq = 0;
                    if (lane == 0) { s_misc[1] = 0; s_misc[2] = 0; }
                    __syncwarp();
                    if (!skip) {
                        const u32 ws = t0 + 64 * lane, we = ws + 64;
                        const u64 okw = (u64)s_okb[2 * lane] | ((u64)s_okb[2 * lane + 1] << 32);
                        u64 sel = 0;                               // starts of the matches the chain takes in my window
                        u32 exitv = 0;                             // end of the last of them (0: none)
                        u32 from = cur > ws ? cur : ws;            // where the chain enters my window, as far as known
                        bool need = from < we;
                        for (u32 round = 0;; round++) {
                            u64 nsel = 0; u32 pos = from, nexit = 0; bool act = need;
                            while (__any_sync(FULL, act)) {
                                bool lng = false; u32 r = 0;
                                if (act) {
                                    const u64 m = okw >> (pos - ws);
                                    if (!m) act = false;
                                    else {
                                        const u32 p = pos + (u32)__ffsll((long long)m) - 1u;
                                        const u64 bit = 1ull << (p - ws);
                                        if (round && (sel & bit)) {                 // joins the chain of the previous round
                                            nsel |= sel & ~(bit - 1ull); nexit = exitv; act = false;
                                        } else {
                                            r = c3_run_end(s_cont, p - t0);
                                            const u32 x = s_ext[r];
                                            nsel |= bit;
                                            if (x == 255u) lng = true;
                                            else { pos = t0 + r + MINMATCH + x; nexit = pos; if (pos >= we) act = false; }
                                        }
                                    }
                                }
                                unsigned lm = __ballot_sync(FULL, lng);
                                while (lm) {                                            // long matches: remembered, else finished by the warp
                                    const int srcl = __ffs(lm) - 1; lm &= lm - 1;
                                    const u32 rr = __shfl_sync(FULL, r, srcl);
                                    u32 e = 0;
                                    const u32 nl = s_misc[2];
                                    for (u32 i = 0; i < nl; i++) if (s_longm[i] == rr) e = s_longm[C3_NLONGM + i];
                                    if (!e) {
                                        const u32 o = hc[rr], p = t0 + rr + MINMATCH;
                                        u32 len = C3_CAP;
                                        for (;;) {
                                            const u32 a = p + len + lane * 4;
                                            u32 c = 0;
                                            if (a + 4 <= mlimit) {
                                                const u32 x = c1_rd32(s_in + a) ^ c1_rd32(s_in + a - o);
                                                c = x ? ((__ffs(x) - 1) >> 3) : 4;
                                            } else {
                                                while (a + c < mlimit && s_in[a + c] == s_in[a + c - o]) c++;
                                            }
                                            const unsigned full = __ballot_sync(FULL, c == 4);
                                            if (full == FULL) { len += 128; continue; }
                                            const int fl = __ffs(~full) - 1;
                                            len += fl * 4 + __shfl_sync(FULL, c, fl);
                                            break;
                                        }
                                        e = p + len;
                                        if (lane == 0 && nl < C3_NLONGM) { s_longm[nl] = (u16)rr; s_longm[C3_NLONGM + nl] = (u16)e; s_misc[2] = nl + 1; }
                                        __syncwarp();
                                    }
                                    if ((int)lane == srcl) { pos = e; nexit = e; if (pos >= we) act = false; }
                                }
                            }
                            if (need) { sel = nsel; exitv = nexit; }
                            // repair: the chain enters a window at max(cur, ends of the matches of the earlier windows); a window that lies
                            // inside an earlier match takes nothing (and contributes nothing)
                            u32 incl, nfrom;
                            for (;;) {
                                incl = exitv;
                                #pragma unroll
                                for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, d); if (lane >= d && y > incl) incl = y; }
                                u32 ent = __shfl_up_sync(FULL, incl, 1);
                                if (lane == 0 || ent < cur) ent = cur;
                                nfrom = ent > ws ? ent : ws;
                                const bool cov = nfrom >= we && (exitv != 0 || sel != 0);
                                if (!__any_sync(FULL, cov)) break;
                                if (cov) { sel = 0; exitv = 0; from = 0xFFFFFFFFu; }
                            }
                            need = nfrom < we && nfrom != from;
                            if (need) from = nfrom;
                            if (!__any_sync(FULL, need)) {
                                const u32 last = __shfl_sync(FULL, incl, 31);
                                cur_out = last > cur ? last : cur;
                                break;
                            }
                        }
                        // the sequences of the tile, in order
                        const u32 cnt = (u32)__popcll(sel);
                        u32 base = cnt;
                        #pragma unroll
                        for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, base, d); if (lane >= d) base += y; }
                        nseq = __shfl_sync(FULL, base, 31);
                        base -= cnt;
                        for (u64 s = sel; s; s &= s - 1ull) s_seqp[base++] = (u16)(ws + (u32)__ffsll((long long)s) - 1u);
                    }
                    if (lane == 0) { s_misc[3] = nseq; s_misc[4] = cur_out; }
                } else if (t + 2 < n_tiles) {
                    hash_tile(t + 2, warp - 1, NWORK - 1);
                }
                const long long q2 = prof ? clock64() : 0;
                c3_bar_work<NWT>();
                // ---- E: one lane per sequence.  Pass A: match end, backward extension, encoded size ----
                const u32 nseq = s_misc[3], cur_next = s_misc[4];
                auto match_end = [&](u32 i) -> u32 {                 // i: tile-relative start of a match the chain took
                    const u32 r = c3_run_end(s_cont, i);
                    const u32 x = s_ext[r];
                    if (x != 255u) return t0 + r + MINMATCH + x;
                    const u32 nl = s_misc[2];
                    for (u32 q = 0; q < nl; q++) if (s_longm[q] == r) return s_longm[C3_NLONGM + q];
                    const u32 a4 = t0 + r + MINMATCH;                // not remembered (more than C3_NLONGM long runs were looked at): once more
                    return a4 + c1_count(s_in, a4, a4 - hc[r], mlimit, 0xFFFFFFFFu);
                };
                for (u32 b = warp; b * 32 < nseq; b += NWORK) {
                    const u32 j = b * 32 + lane;
                    u32 enc = 0;
                    u32 e = 0;
                    if (j < nseq) e = match_end((u32)s_seqp[j] - t0);
                    u32 anchor = __shfl_up_sync(FULL, e, 1);
                    if (lane == 0) {
                        anchor = cur;
                        if (j && j < nseq) anchor = match_end((u32)s_seqp[j - 1] - t0);   // the sequence before mine belongs to another warp: its end once more
                    }
                    if (j < nseq) {
                        u32 p = s_seqp[j];
                        const u32 o = hc[p - t0];
                        while (p > anchor && p > o && s_in[p - 1] == s_in[p - o - 1]) p--;      // lz4/lz4.c:1019
                        const u32 lit = p - anchor, ml = e - p - MINMATCH;
                        enc = 1 + c1_len_ext_bytes(lit) + lit + 2 + c1_len_ext_bytes(ml);
                        s_seqq[j] = (u16)p; s_seqe[j] = (u16)e; s_seqf[j] = (u16)o;
                    }
                    u32 incl = enc;
                    #pragma unroll
                    for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += y; }
                    if (j < nseq) s_seqo[j] = incl - enc;
                    if (lane == 31) s_blk[b] = incl;
                }
                const long long q3 = prof ? clock64() : 0;
                c3_bar_work<NWT>();
                // ---- pass B: the bytes ----
                u32 tile_bytes = 0;
                {
                    const u32 nblk = (nseq + 31) / 32;
                    for (u32 b = warp; b < nblk; b += NWORK) {
                        const u32 j = b * 32 + lane;
                        u32 before = 0;
                        for (u32 q = 0; q < b; q++) before += s_blk[q];
                        if (j < nseq) {
                            const u32 p = s_seqq[j], e = s_seqe[j], o = s_seqf[j];
                            const u32 anchor = j ? (u32)s_seqe[j - 1] : cur;
                            const u32 lit = p - anchor, ml = e - p - MINMATCH;
                            u8* q = dst + op + before + s_seqo[j];
                            *q++ = (u8)(((lit < RUN_MASK ? lit : RUN_MASK) << 4) | (ml < ML_MASK ? ml : ML_MASK));
                            if (lit >= RUN_MASK) q = c1_put_ext(q, lit - RUN_MASK);
                            if (lit >= C3_LONG_LIT) {
                                const u32 slot = atomicAdd(&s_misc[1], 1u);
                                s_longq[slot] = make_uint4(anchor, (u32)(q - dst), lit, 0);
                            } else {
                                for (u32 i = 0; i < lit; i++) q[i] = s_in[anchor + i];
                            }
                            q += lit;
                            *q++ = (u8)o; *q++ = (u8)(o >> 8);
                            if (ml >= ML_MASK) q = c1_put_ext(q, ml - ML_MASK);
                        }
                    }
                    for (u32 q = 0; q < nblk; q++) tile_bytes += s_blk[q];
                }
                op += tile_bytes;
                cur = cur_next;
                if (tid == 0) { s_misc[5] = cur; s_misc[6] = op; }
                if (prof && tid == 0) {
                    const long long q4 = clock64();
                    atomicAdd((unsigned long long*)&prof[0], (unsigned long long)(q1 - q0));
                    atomicAdd((unsigned long long*)&prof[1], (unsigned long long)(q2 - q1));
                    atomicAdd((unsigned long long*)&prof[2], (unsigned long long)(q3 - q2));
                    atomicAdd((unsigned long long*)&prof[3], (unsigned long long)(q4 - q3));
                    atomicAdd((unsigned long long*)&prof[4], 1ull);
                }
                __syncthreads();
            }
        }
        __syncthreads();
        // ---- long literal runs of the last tile, then the last literals (lz4/lz4.c:1204-1231), all threads ----
        {
            const u32 nlong = s_misc[1];
            for (u32 e = warp; e < nlong; e += NWORK + 1) {
                const uint4 q = s_longq[e];
                for (u32 i = lane; i < q.z; i += 32) dst[q.y + i] = s_in[q.x + i];
            }
            const u32 cur = s_misc[5];
            u32 op = s_misc[6];
            const u32 last = n - cur;
            u8* o = dst + op;
            const u32 hdr = 1 + c1_len_ext_bytes(last);
            if (tid == 0) {
                *o = (u8)((last < RUN_MASK ? last : RUN_MASK) << 4);
                if (last >= RUN_MASK) c1_put_ext(o + 1, last - RUN_MASK);
            }
            for (u32 i = tid; i < last; i += NTHR) o[hdr + i] = s_in[cur + i];
            op += hdr + last;
            if (tid == 0) csz[chunk] = op;
        }
        __syncthreads();
    }
}

template <u32 NWORK>
inline cudaError_t c3_launch_n(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                               u32 n, u32* ticket, int sm_count, cudaStream_t st, u64* prof)
{
    typedef C3Smem<NWORK> S;
    cudaError_t e = cudaFuncSetAttribute(compress_c3_kernel<NWORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES);   // per device: every launch
    if (e != cudaSuccess) return e;
    unsigned grid = (unsigned)(sm_count * 2);
    if (grid > n) grid = n;
    e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    compress_c3_kernel<NWORK><<<grid, (NWORK + 1) * 32, S::BYTES, st>>>(in, in_off, usz, stage, stride, csz, n, ticket, prof);
    return cudaGetLastError();
}

inline cudaError_t c3_launch_compress(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                                      u32 n, u32* ticket, int sm_count, cudaStream_t st, u64* prof = nullptr, int workers = 8)
{
    if (workers == 16) return c3_launch_n<16>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, prof);
    if (workers == 12) return c3_launch_n<12>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, prof);
    return c3_launch_n<8>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, prof);
}

}  // namespace lz4b200
