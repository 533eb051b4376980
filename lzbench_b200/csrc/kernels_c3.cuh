// kernels_c3.cuh -- generation-3 compressor: the c2 algorithm taken apart once more, for instruction count and overlap.
//
// The c2 profile (profiles/r01_final_compress_c2_ncu_*): 6.3 G warp instructions per GiB (5.9 per input byte), IPC 2.25,
// every phase issue-bound.  Where they go: the match finder spends ~85 thread instructions per position (unaligned
// reads, a batch-local table kept with atomics, two verifications), and the parse walks every speculative window
// through the data (extension compares inside a divergent per-lane loop, 2.4 rounds per tile).  c3 keeps the result
// (LZ4_compress_default's greedy parse, lz4/lz4.c:851-1240, over an every-position hash table) and changes who does what.
// Five stages, each on its own warps, each on a different 2048-position tile at any moment (one CTA barrier per step):
//
//   H  hash (warps 0..7, tile k): 4 consecutive positions per lane from two aligned words, one funnel shift each;
//      13-bit hash (lz4/lz4.c:698-702) -> u16 per position.
//   T  table (one warp, tile k-1), strictly in position order, 32 positions per step: candidate = table[h],
//      table[h] = position (lz4/lz4.c:736-787; last writer wins).  Nothing but shared-memory traffic, plus a shuffle
//      for "same hash as the position before" (runs).  A step reads before it writes, so a position misses a repeat
//      that lies in its own step; the CPU model (scripts/sim/compress_sim.c) puts that at 0.05 % of the output on the
//      bench corpus and 0.3 % on real binaries (the backward extension, lz4/lz4.c:1019, gets most of it back).
//   V  verify + extend (warps 0..7, tile k-2), 4 positions per lane: 4 bytes at the candidate == 4 bytes here ->
//      offset (u16, in place) and the bit map of matched positions.  Positions that follow each other with the same
//      offset are the same match minus its first bytes, so they all END at the same place: the extension compare
//      is done once per RUN, at its last position, densely (run ends compacted inside the warp: up to 16 bytes by
//      the lane, one 128-byte step by the warp for the longer ones) and handed back along the run by a segmented
//      scan -> match length per position (u8; 255 = longer than that, finished when the parse really takes it).
//   P  parse (one warp, tile k-3), bit map + lengths only: the greedy chain "first matched position at or after
//      the end of the previous match; jump to its end" (lz4/lz4.c:956-1014) over 32 windows of 64 positions, entry
//      states repaired by prefix-max rounds as in c1; a step is a shift, a find-first-set and one byte load.
//   E  emit (4 warps, tile k-4), one lane per sequence: backward extension (lz4/lz4.c:1019), sizes, scan, bytes
//      (lz4/lz4.c:1022-1135).
//
// Same block format rules as c1/c2 (end-of-block rules lz4/lz4.c:883-884, last literals :1204-1231).  Chunks above 64 KiB
// come as jobs of 16 KiB history + 48 KiB new data (segmented.cuh); the lz4fast levels stay on c2.
#pragma once
#include "common.cuh"
#include "kernels_c1.cuh"

namespace lz4b200 {

constexpr u32 C3_TILE = 2048;                    // positions per tile
constexpr u32 C3_SLICE = 128;                    // positions per warp step of H and V (4 per lane)
constexpr u32 C3_CAP1 = 32;                      // bytes a lane compares past a run end (then one 128-byte step by the warp); measured 8 / 16 / 32 / 48 / 64: 101.7 / 111.5 / 117.7 / 117.2 / 115.7 GB/s
constexpr u32 C3_LONG = 255;                     // match length byte: 255 or more, to be measured when needed
constexpr u32 C3_KNOWN = (MINMATCH + C3_CAP1 + 128 < C3_LONG ? MINMATCH + C3_CAP1 + 128 : C3_LONG) & ~3u;   // bytes of a long match that are known to be equal
constexpr u32 C3_MAXSEQ = C3_TILE / MINMATCH;    // sequences that can start inside one tile
constexpr u32 C3_LONG_LIT = 64;                  // literal runs this long are copied by a whole warp
constexpr u32 C3_NLONGQ = C3_TILE / C3_LONG_LIT + 2;
constexpr u32 C3_NLONGM = 16;                    // long matches per half tile whose end is remembered
constexpr u32 C3_NVR = 8;                        // warps of H and V
constexpr u32 C3_NE = 4;                         // warps of E
// Warp numbers.  The table warp and the two parse warps are serial chains that everything else waits for: they have scheduler 0
// (warp number mod 4) to themselves -- tried against instruction fetch and issue competition, measured: no difference, their
// instructions wait for each other, not for the scheduler.  Warp 12, the rest of that scheduler, has no stage: it helps with the
// loading and clearing between chunks and takes part in the CTA barriers.
constexpr u32 C3_WT = 0;
constexpr u32 C3_WP = 4;
constexpr u32 C3_WP2 = 8;                        // the second parse warp
constexpr u32 C3_THREADS = 16 * 32;
constexpr u32 C3_LIVE = C3_THREADS;              // threads that share the work between chunks

struct C3Smem {
    static constexpr u32 OFF_IN = 0;
    static constexpr u32 IN_CAP = C1_CHUNK_MAX + 64;
    static constexpr u32 OFF_TAB = OFF_IN + IN_CAP;                     // u16[8192]
    static constexpr u32 OFF_HC = OFF_TAB + 2 * (1u << C1_HBITS);       // u16[4][TILE]: hash -> candidate -> offset, in place
    static constexpr u32 OFF_MLEN = OFF_HC + 4 * 2 * C3_TILE;           // u8[2][TILE]: length of the match that starts here (255: more)
    static constexpr u32 OFF_OKB = OFF_MLEN + 2 * C3_TILE;              // u32[2][TILE / 32]
    static constexpr u32 OFF_SEQ = OFF_OKB + 2 * 4 * (C3_TILE / 32);    // u16[2][3][MAXSEQ]: start, end, offset of the matches the chain took
    static constexpr u32 OFF_HDR = OFF_SEQ + 2 * 3 * 2 * C3_MAXSEQ;     // u32[2][2]: their number, chain state before the tile
    static constexpr u32 OFF_SEQQ = OFF_HDR + 16;                       // u16[MAXSEQ] start after the backward extension
    static constexpr u32 OFF_SEQO = OFF_SEQQ + 2 * C3_MAXSEQ;           // u32[MAXSEQ] bytes the earlier sequences of its block of 32 take
    static constexpr u32 OFF_BLK = OFF_SEQO + 4 * C3_MAXSEQ;            // u32[MAXSEQ / 32] bytes per block of 32 sequences
    static constexpr u32 OFF_LIST = OFF_BLK + 4 * (C3_MAXSEQ / 32);     // u8[NVR][2][SLICE] run ends of a slice and their extensions, per warp
    static constexpr u32 OFF_LONGQ = (OFF_LIST + C3_NVR * 2 * C3_SLICE + 15) & ~15u;   // uint4[NLONGQ] long literal runs
    static constexpr u32 OFF_LONGM = OFF_LONGQ + 16 * C3_NLONGQ;        // u16[2][2][NLONGM] long matches, per parse warp: start, end
    static constexpr u32 OFF_MISC = OFF_LONGM + 2 * 4 * C3_NLONGM;      // mbarrier (8 B) + scalars
    static constexpr u32 BYTES = OFF_MISC + 64;
};
static_assert(C3Smem::BYTES <= 115712, "two CTAs per SM");

__device__ __forceinline__ void c3_bar_emit() { asm volatile("bar.sync 1, %0;" ::"n"(C3_NE * 32) : "memory"); }

// common prefix of in[a..] and in[a - o..] from `len` bytes on, 128 bytes per warp step, never reading in[limit..] as data;
// at most `steps` steps (0: no bound).  Returns the length, or 0xFFFFFFFF when the bound was reached with all bytes equal.
__device__ __forceinline__ u32 c3_warp_count(const u8* s_in, u32 a, u32 o, u32 len, u32 limit, u32 steps, u32 lane)
{
    for (u32 it = 0;; it++) {
        const u32 x0 = a + len + lane * 4;
        u32 c = 0;
        if (x0 + 4 <= limit) {
            const u32 x = c1_rd32(s_in + x0) ^ c1_rd32(s_in + x0 - o);
            c = x ? ((__ffs(x) - 1) >> 3) : 4;
        } else {
            while (x0 + c < limit && s_in[x0 + c] == s_in[x0 + c - o]) c++;
        }
        const unsigned full = __ballot_sync(FULL, c == 4);
        if (full == FULL) {
            len += 128;
            if (steps && it + 1 == steps) return 0xFFFFFFFFu;
            continue;
        }
        const int fl = __ffs(~full) - 1;
        return len + fl * 4 + __shfl_sync(FULL, c, fl);
    }
}

// the same for matches that are known to be long: 16 bytes per lane, 512 per step, no bound
__device__ __forceinline__ u32 c3_warp_count16(const u8* s_in, u32 a, u32 o, u32 len, u32 limit, u32 lane)
{
    for (;;) {
        const u32 x0 = a + len + lane * 16;
        u32 c = 0;
        if (x0 + 16 <= limit) {
            const u32 y0 = c1_rd32(s_in + x0) ^ c1_rd32(s_in + x0 - o), y1 = c1_rd32(s_in + x0 + 4) ^ c1_rd32(s_in + x0 + 4 - o);
            const u32 y2 = c1_rd32(s_in + x0 + 8) ^ c1_rd32(s_in + x0 + 8 - o), y3 = c1_rd32(s_in + x0 + 12) ^ c1_rd32(s_in + x0 + 12 - o);
            c = y0 ? ((__ffs(y0) - 1) >> 3) : y1 ? 4 + ((__ffs(y1) - 1) >> 3) : y2 ? 8 + ((__ffs(y2) - 1) >> 3) : y3 ? 12 + ((__ffs(y3) - 1) >> 3) : 16;
        } else {
            while (c < 16 && x0 + c < limit && s_in[x0 + c] == s_in[x0 + c - o]) c++;
        }
        const unsigned full = __ballot_sync(FULL, c == 16);
        if (full == FULL) { len += 512; continue; }
        const int fl = __ffs(~full) - 1;
        return len + fl * 16 + __shfl_sync(FULL, c, fl);
    }
}

template <bool PROF>                             // PROF: per-stage cycle counters (scripts/gpu_probe.py PROF=1); the production instance carries none
__global__ void __launch_bounds__(C3_THREADS, 2)
compress_c3_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ usz,
                   u8* __restrict__ stage, u64 slot_stride, u64* __restrict__ csz, u32 n_chunks_arg, u32* ticket,
                   u64* __restrict__ prof_arg, const u64* __restrict__ n_dev, u32* __restrict__ tail_pos, u32* __restrict__ tail_len,
                   const u32* __restrict__ hist_arr)
{
    // segmented mode (segmented.cuh, chunks above 64 KiB): the job count lives in device memory, a job starts with `hist` bytes of
    // history (hashed and matchable, not encoded), a job that is not the end of its chunk is exempt from the end-of-block rules,
    // and every job reports where its last literals start and how many they are
    const u32 n_chunks = n_dev ? (u32)*n_dev : n_chunks_arg;
    typedef C3Smem S;
    u64* const prof = PROF ? prof_arg : nullptr;
    extern __shared__ __align__(128) u8 smem[];
    u8* const s_raw = smem + S::OFF_IN;
    u16* const s_tab = reinterpret_cast<u16*>(smem + S::OFF_TAB);
    u16* const s_hc = reinterpret_cast<u16*>(smem + S::OFF_HC);
    u8* const s_mlen = smem + S::OFF_MLEN;
    u32* const s_okb = reinterpret_cast<u32*>(smem + S::OFF_OKB);
    u16* const s_seq = reinterpret_cast<u16*>(smem + S::OFF_SEQ);
    u32* const s_hdr = reinterpret_cast<u32*>(smem + S::OFF_HDR);
    u16* const s_seqq = reinterpret_cast<u16*>(smem + S::OFF_SEQQ);
    u32* const s_seqo = reinterpret_cast<u32*>(smem + S::OFF_SEQO);
    u32* const s_blk = reinterpret_cast<u32*>(smem + S::OFF_BLK);
    u8* const s_list = smem + S::OFF_LIST;
    uint4* const s_longq = reinterpret_cast<uint4*>(smem + S::OFF_LONGQ);
    u16* const s_longm = reinterpret_cast<u16*>(smem + S::OFF_LONGM);
    u64* const s_bar = reinterpret_cast<u64*>(smem + S::OFF_MISC);
    volatile u32* const s_misc = reinterpret_cast<volatile u32*>(smem + S::OFF_MISC + 16);   // [0] chunk [1] long-literal count [2] long-match count [3] chain state P has reached [5] final chain state [6] final output bytes

    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // H and V: 0 .. NVR-1, E: NVR .. NVR+NE-1 in the order of the warp numbers 1,2,3,5,6,7,9,10,11,13,14,15; the table and parse warps get large numbers
    const u32 warp = (wid & 3u) ? (wid >> 2) * 3 + (wid & 3u) - 1 : 0xFFFFu;
    const u32 ltid = tid;                             // (warp 12 has no stage: it helps with the loading and clearing between chunks and takes part in the CTA barriers)
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
            for (u32 i = ltid; i < (2u << C1_HBITS) / 16; i += C3_LIVE) t4[i] = make_uint4(0, 0, 0, 0);
            if (bulk) { mbar_wait(s_bar, phase); phase ^= 1; }
        } else {                                      // source not on a 4-byte boundary: realigning word copy
            const u32 sh = (head & 3u) * 8u;
            const u32* sw = reinterpret_cast<const u32*>(src - (head & 3u));
            const u32 nw = n >= 8 ? n / 4 - 1 : 0;   // words whose two source words lie inside the chunk
            u32* dw = reinterpret_cast<u32*>(s_in);
            for (u32 k = ltid; k < nw; k += C3_LIVE) dw[k] = __funnelshift_r(__ldg(sw + k), __ldg(sw + k + 1), sh);
            for (u32 i = nw * 4 + ltid; i < n; i += C3_LIVE) s_in[i] = src[i];
            uint4* t4 = reinterpret_cast<uint4*>(s_tab);
            for (u32 i = ltid; i < (2u << C1_HBITS) / 16; i += C3_LIVE) t4[i] = make_uint4(0, 0, 0, 0);
        }
        const u32 hist = hist_arr ? (hist_arr[chunk] & 0x7FFFFFFFu) : 0u;     // the chain starts behind the history
        const bool inner = hist_arr && (hist_arr[chunk] >> 31);
        if (tid == 0) { s_misc[1] = 0; s_misc[3] = hist; s_misc[5] = hist; s_misc[6] = 0; }
        __syncthreads();
        if (prof && tid == 0) atomicAdd((unsigned long long*)&prof[5], (unsigned long long)(clock64() - pt0));

        if (n >= MIN_LENGTH) {
            const u32 mfl = n - (inner ? MINMATCH : MFLIMIT);     // last position a match may start at (lz4/lz4.c:883-884 bind only the real end of a block)
            const u32 mlimit = n - (inner ? 0u : LASTLITERALS);   // matches end at or before this
            const u32 n_tiles = mfl / C3_TILE + 1;
            const u32* const s_in32 = reinterpret_cast<const u32*>(s_in);
            u32 cur = 0;                              // P, E: chain state = end of the last match (= literal anchor)
            u32 op = 0;                               // E: bytes emitted so far

            for (u32 k = 0; k < n_tiles + 4; k++) {
                const long long q0 = prof ? clock64() : 0;
                if (warp < C3_NVR) {                 // (unsigned: not the table / parse warps)
                    // ================= H of tile k =================
                    if (k < n_tiles) {
                        u16* const hc = s_hc + (k & 3) * C3_TILE;
                        for (u32 s = warp; s < C3_TILE / C3_SLICE; s += C3_NVR) {
                            const u32 i0 = s * C3_SLICE + 4 * lane;
                            const u32* w = s_in32 + ((k * C3_TILE + i0) >> 2);
                            const u32 a = w[0], b = w[1];
                            const u32 h0 = hash13(a), h1 = hash13(__funnelshift_r(a, b, 8));
                            const u32 h2 = hash13(__funnelshift_r(a, b, 16)), h3 = hash13(__funnelshift_r(a, b, 24));
                            *reinterpret_cast<uint2*>(hc + i0) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
                        }
                    }
                    // ================= V of tile k - 2 =================
                    const u32 t = k - 2;
                    if (k >= 2 && t < n_tiles && s_misc[3] < (t + 1) * C3_TILE) {     // (a tile inside a match the chain has taken: nothing to find)
                        const u32 t0 = t * C3_TILE;
                        u16* const hc = s_hc + (t & 3) * C3_TILE;
                        u8* const mlen = s_mlen + (t & 1) * C3_TILE;
                        u32* const okb = s_okb + (t & 1) * (C3_TILE / 32);
                        u8* const list = s_list + warp * 2 * C3_SLICE;
                        u8* const extw = list + C3_SLICE;
                        for (u32 s = warp; s < C3_TILE / C3_SLICE; s += C3_NVR) {
                            const u32 i0 = s * C3_SLICE + 4 * lane, p0 = t0 + i0;
                            const u32* w = s_in32 + (p0 >> 2);
                            const u32 a = w[0], b = w[1];
                            const uint2 cc = *reinterpret_cast<const uint2*>(hc + i0);
                            const u32 v[4] = { a, __funnelshift_r(a, b, 8), __funnelshift_r(a, b, 16), __funnelshift_r(a, b, 24) };
                            const u32 c[4] = { cc.x & 0xFFFFu, cc.x >> 16, cc.y & 0xFFFFu, cc.y >> 16 };
                            u32 off[4];
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                const u32 there = c1_rd32(s_in + c[j]);               // (any candidate lies inside the buffer: no branch)
                                const bool ok = (p0 + j <= mfl) & (c[j] < p0 + j) & (there == v[j]);
                                off[j] = ok ? p0 + j - c[j] : 0u;
                            }
                            *reinterpret_cast<uint2*>(hc + i0) = make_uint2(off[0] | (off[1] << 16), off[2] | (off[3] << 16));
                            if (!__any_sync(FULL, (off[0] | off[1] | off[2] | off[3]) != 0)) {     // nothing matched in the slice (incompressible data)
                                if ((lane & 7) == 0) okb[i0 >> 5] = 0;
                                continue;
                            }
                            u32 prev = __shfl_up_sync(FULL, off[3], 1);
                            if (lane == 0) prev = 0;                    // a slice starts a new run
                            u32 okn = 0, cn = 0;
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                okn |= (u32)(off[j] != 0) << j;
                                cn |= (u32)(off[j] != 0 && off[j] == (j ? off[j - 1] : prev)) << j;
                            }
                            u32 nextc = __shfl_down_sync(FULL, cn & 1u, 1);
                            if (lane == 31) nextc = 0;
                            const u32 ren = okn & ~((cn >> 1) | (nextc << 3));   // run ends: matched, and the next position does not continue
                            u32 okw = okn << (4 * (lane & 7));                  // bit map: lanes 8j .. 8j+7 hold word j of the slice
                            #pragma unroll
                            for (u32 d = 1; d < 8; d <<= 1) okw |= __shfl_xor_sync(FULL, okw, d);
                            if ((lane & 7) == 0) okb[i0 >> 5] = okw;
                            // run ends, compacted (any order will do) and extended densely
                            u32 total = 0;
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                const unsigned bj = __ballot_sync(FULL, (ren >> j) & 1u);
                                if ((ren >> j) & 1u) list[total + __popc(bj & ((1u << lane) - 1u))] = (u8)(4 * lane + j);
                                total += __popc(bj);
                            }
                            __syncwarp();
                            for (u32 base = 0; base < total; base += 32) {
                                const u32 q = base + lane;
                                u32 x = 0, a4 = 0, o = 0, idx = 0;
                                if (q < total) {
                                    idx = list[q];
                                    o = hc[s * C3_SLICE + idx]; a4 = t0 + s * C3_SLICE + idx + MINMATCH;
                                    x = c1_count(s_in, a4, a4 - o, mlimit, C3_CAP1);
                                }
                                unsigned more = __ballot_sync(FULL, x >= C3_CAP1);
                                while (more) {                                          // 16 bytes and going: one step by the whole warp
                                    const int sl = __ffs(more) - 1; more &= more - 1;
                                    const u32 r = c3_warp_count(s_in, __shfl_sync(FULL, a4, sl), __shfl_sync(FULL, o, sl), C3_CAP1, mlimit, 1, lane);
                                    if ((int)lane == sl) x = r == 0xFFFFFFFFu ? C3_LONG : r;
                                }
                                if (q < total) extw[idx] = (u8)x;
                            }
                            __syncwarp();
                            // the end of a run, handed back to all its positions (ends are slice-relative; 0xFFFF: long)
                            const u32 ex4 = *reinterpret_cast<const u32*>(extw + 4 * lane);
                            u32 own[4];
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                const u32 x = (ex4 >> (8 * j)) & 0xFFu;
                                own[j] = x == C3_LONG ? 0xFFFFu : 4 * lane + j + MINMATCH + x;
                            }
                            const u32 first = ren ? (u32)__ffs(ren) - 1u : 0u;
                            // end of the run my position 0 is in, if it ends with me (bit 31 clear), else the same as for the next lane's position 0
                            u32 hv = ren == 0 ? 0x80000000u : first == 0 ? own[0] : first == 1 ? own[1] : first == 2 ? own[2] : own[3];
                            #pragma unroll
                            for (u32 d = 1; d < 32; d <<= 1) {
                                const u32 v2 = __shfl_down_sync(FULL, hv, d);
                                if ((hv >> 31) && lane + d < 32) hv = v2;
                            }
                            const u32 cin = __shfl_down_sync(FULL, hv, 1) & 0xFFFFu;
                            u32 e[4];
                            e[3] = ((ren >> 3) & 1u) ? own[3] : cin;
                            e[2] = ((ren >> 2) & 1u) ? own[2] : e[3];
                            e[1] = ((ren >> 1) & 1u) ? own[1] : e[2];
                            e[0] = (ren & 1u) ? own[0] : e[1];
                            u32 ml4 = 0;
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                u32 m = e[j] - (4 * lane + j);                          // (>= 4 where the position is matched; unused elsewhere)
                                if (m > C3_LONG) m = C3_LONG;
                                ml4 |= (m & 0xFFu) << (8 * j);
                            }
                            *reinterpret_cast<u32*>(mlen + i0) = ml4;
                            __syncwarp();
                        }
                    }
                    if (prof && warp == 0 && lane == 0) atomicAdd((unsigned long long*)&prof[0], (unsigned long long)(clock64() - q0));
                } else if (wid == C3_WT) {
                    // ================= T of tile k - 1 =================
                    const u32 t = k - 1;
                    if (k >= 1 && t < n_tiles) {
                        u16* const hc = s_hc + (t & 3) * C3_TILE;
                        const u32 t0 = t * C3_TILE;
                        for (u32 s = 0; s < C3_TILE / 32; s += 4) {
                            u32 h[4], c[4];
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) h[j] = hc[(s + j) * 32 + lane];
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                c[j] = s_tab[h[j]];
                                s_tab[h[j]] = (u16)(t0 + (s + j) * 32 + lane);
                            }
                            #pragma unroll
                            for (u32 j = 0; j < 4; j++) {
                                const u32 hp = __shfl_up_sync(FULL, h[j], 1);
                                if (lane && hp == h[j]) c[j] = t0 + (s + j) * 32 + lane - 1;   // the table is a step behind: a run would miss its own first step
                                hc[(s + j) * 32 + lane] = (u16)c[j];
                            }
                        }
                        if (prof && lane == 0) atomicAdd((unsigned long long*)&prof[1], (unsigned long long)(clock64() - q0));
                    }
                } else if (wid == C3_WP || wid == C3_WP2) {
                    // ================= P of tile k - 3: two warps, one half of the tile each =================
                    // The second warp starts from a guess (the chain enters its half at the half's first position) and is told the
                    // state the first warp ends in when both have settled; one more repair round then takes it to the true chain.
                    const u32 t = k - 3, half = wid == C3_WP2;
                    if (k >= 3 && t < n_tiles) {
                        const u32 t0 = t * C3_TILE, h0 = t0 + half * (C3_TILE / 2);
                        const u16* const hc = s_hc + (t & 3) * C3_TILE;
                        const u8* const mlen = s_mlen + (t & 1) * C3_TILE;
                        const u32* const okb = s_okb + (t & 1) * (C3_TILE / 32);
                        u16* const seq = s_seq + (t & 1) * 3 * C3_MAXSEQ;
                        u16* const longm = s_longm + half * 2 * C3_NLONGM;
                        volatile u32* const nlongm = s_misc + (half ? 7 : 2);
                        const u32 cur_tile = s_misc[3];                      // chain state before the tile (the second warp left it there)
                        u32 curh = cur_tile;                                 // ... before my half (second warp: a guess -- no match of the first half ends in mine)
                        u32 cur_out = cur_tile, nseq = 0, seq_base = 0;
                        if (lane == 0) *nlongm = 0;
                        __syncwarp();
                        // the end of a long match that starts at p: remembered, else measured by the whole warp (uniform p)
                        auto long_end = [&](u32 p) -> u32 {
                            const u32 nl = *nlongm;
                            for (u32 i = 0; i < nl; i++) if (longm[i] == p) return (u32)longm[C3_NLONGM + i] + 1u;
                            const u32 o = hc[p - t0];
                            const u32 e = p + c3_warp_count16(s_in, p, o, C3_KNOWN, mlimit, lane);
                            if (lane == 0 && nl < C3_NLONGM) { longm[nl] = (u16)p; longm[C3_NLONGM + nl] = (u16)(e - 1u); *nlongm = nl + 1; }
                            __syncwarp();
                            return e;
                        };
                        const bool live = cur_tile < t0 + C3_TILE;           // (else the tile lies inside a match: nothing to take)
                        const u32 ws = h0 + 32 * lane, we = ws + 32;
                        const u32 okw = live ? okb[half * 32 + lane] : 0u;   // matched positions of my window
                        const u8* const mlw = mlen + (ws - t0);
                        u32 sel = 0;                                         // starts of the matches the chain takes in my window
                        u32 exitv = 0;                                       // end of the last of them (0: none)
                        u32 from = curh > ws ? curh : ws;                    // where the chain enters my window, as far as known
                        bool need = live && from < we, have = false;         // have: sel / exitv are the chain from `from`
                        auto settle = [&](bool chase_first) {
                            for (bool chase = chase_first;; chase = true) {
                                if (chase) {
                                    u32 nsel = 0, pos = from, nexit = 0; bool act = need, dropped = false;
                                    if (prof && lane == 0) atomicAdd((unsigned long long*)&prof[7], 1ull);
                                    while (__any_sync(FULL, act)) {
                                        if (prof && lane == 0) atomicAdd((unsigned long long*)&prof[8], 1ull);
                                        bool lng = false; u32 p = 0;
                                        if (act) {
                                            const u32 q = pos - ws, m = okw >> q;      // the first matched position at or after pos
                                            if (!m) act = false;
                                            else {
                                                const u32 idx = q + (u32)__ffs(m) - 1u, bit = 1u << idx;
                                                p = ws + idx;
                                                if (have && (sel & bit)) { nsel |= sel & ~(bit - 1u); nexit = exitv; act = false; }   // joins the chain of the round before
                                                else {
                                                    const u32 ml = mlw[idx];
                                                    nsel |= bit;
                                                    if (ml == C3_LONG) lng = true;
                                                    else { pos = p + ml; nexit = pos; act = pos < we; }
                                                }
                                            }
                                        }
                                        // long matches, lowest window first.  A later window's long match that starts inside this one is not
                                        // the chain's: that window is parsed again once its true entry is known (on run-length data every
                                        // window would measure its own)
                                        unsigned lm = __ballot_sync(FULL, lng);
                                        while (lm) {
                                            const int sl = __ffs(lm) - 1;
                                            const u32 e = long_end(__shfl_sync(FULL, p, sl));
                                            if ((int)lane == sl) { pos = e; nexit = e; lng = false; act = pos < we; }
                                            if ((int)lane > sl && lng && p < e) { lng = false; act = false; nsel = 0; nexit = 0; dropped = true; }
                                            lm = __ballot_sync(FULL, lng);
                                        }
                                    }
                                    if (need) { sel = nsel; exitv = nexit; have = !dropped; if (dropped) from = 0xFFFFFFFFu; }
                                }
                                // repair: the chain enters a window at max(state before the half, ends of the matches of the earlier windows);
                                // a window that lies inside an earlier match takes nothing (and contributes nothing)
                                u32 incl, nfrom;
                                for (;;) {
                                    incl = exitv;
                                    #pragma unroll
                                    for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, d); if (lane >= d && y > incl) incl = y; }
                                    u32 ent = __shfl_up_sync(FULL, incl, 1);
                                    if (lane == 0 || ent < curh) ent = curh;
                                    nfrom = ent > ws ? ent : ws;
                                    const bool cov = nfrom >= we && (exitv | sel) != 0;
                                    if (!__any_sync(FULL, cov)) break;
                                    if (cov) { sel = 0; exitv = 0; have = false; from = 0xFFFFFFFFu; }
                                }
                                need = nfrom < we && nfrom != from;
                                if (need) from = nfrom;
                                if (!__any_sync(FULL, need)) {
                                    const u32 last = __shfl_sync(FULL, incl, 31);
                                    cur_out = last > curh ? last : curh;
                                    return;
                                }
                            }
                        };
                        if (live) settle(true);
                        // hand-over: the first warp's end state and number of sequences
                        u32 cnt = (u32)__popc(sel), base = cnt;
                        if (!half) {
                            #pragma unroll
                            for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, base, d); if (lane >= d) base += y; }
                            nseq = __shfl_sync(FULL, base, 31);
                            if (lane == 0) { s_misc[8] = cur_out; s_misc[9] = nseq; }
                        }
                        asm volatile("bar.sync 2, 64;" ::: "memory");
                        if (half) {
                            const u32 mid = s_misc[8];
                            seq_base = s_misc[9];
                            if (live && mid > curh) { curh = mid; settle(false); cnt = (u32)__popc(sel); }
                            else if (mid > cur_out) cur_out = mid;
                            base = cnt;
                            #pragma unroll
                            for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, base, d); if (lane >= d) base += y; }
                            nseq = __shfl_sync(FULL, base, 31);
                        }
                        // the sequences of my half, in order: start, end, offset
                        base += seq_base - cnt;
                        u32 myl = 0xFFFFFFFFu;                                     // the long one of my window (it leaves the window: at most one)
                        for (u32 sb = sel; sb; sb &= sb - 1u) {
                            const u32 idx = (u32)__ffs(sb) - 1u, p = ws + idx;
                            const u32 ml = mlw[idx];
                            seq[base] = (u16)p; seq[C3_MAXSEQ + base] = (u16)(p + ml - 1u); seq[2 * C3_MAXSEQ + base] = hc[p - t0];   // (ends are stored less 1: a job inside a chunk may match up to position 65 536)
                            if (ml == C3_LONG) myl = base;
                            base++;
                        }
                        __syncwarp();
                        unsigned lm = __ballot_sync(FULL, myl != 0xFFFFFFFFu);
                        while (lm) {                                                // the long ones: their ends, remembered from the chase
                            const int sl = __ffs(lm) - 1; lm &= lm - 1;
                            const u32 j = __shfl_sync(FULL, myl, sl);
                            const u32 e = long_end(seq[j]);
                            if (lane == 0) seq[C3_MAXSEQ + j] = (u16)(e - 1u);
                        }
                        if (half && lane == 0) {
                            s_hdr[(t & 1) * 2] = seq_base + nseq; s_hdr[(t & 1) * 2 + 1] = cur_tile;
                            s_misc[3] = cur_out;
                            if (t + 1 == n_tiles) s_misc[5] = cur_out;
                        }
                        if (prof && lane == 0) atomicAdd((unsigned long long*)&prof[half ? 9 : 2], (unsigned long long)(clock64() - q0));
                    }
                } else if (warp < C3_NVR + C3_NE) {
                    // ================= E of tile k - 4: one lane per sequence =================
                    const u32 t = k - 4, ew = warp - C3_NVR;
                    if (k >= 4 && t < n_tiles) {
                        const u16* const seq = s_seq + (t & 1) * 3 * C3_MAXSEQ;
                        const u32 nseq = s_hdr[(t & 1) * 2];
                        cur = s_hdr[(t & 1) * 2 + 1];
                        // long literal runs of the previous tile, whole warps
                        {
                            const u32 nlong = s_misc[1];
                            for (u32 e = ew; e < nlong; e += C3_NE) {
                                const uint4 q = s_longq[e];
                                for (u32 i = lane; i < q.z; i += 32) dst[q.y + i] = s_in[q.x + i];
                            }
                        }
                        // pass A: backward extension, encoded size
                        const u32 nblk = (nseq + 31) / 32;
                        for (u32 b = ew; b < nblk; b += C3_NE) {
                            const u32 j = b * 32 + lane;
                            u32 enc = 0;
                            if (j < nseq) {
                                u32 p = seq[j];
                                const u32 e = (u32)seq[C3_MAXSEQ + j] + 1u, o = seq[2 * C3_MAXSEQ + j];
                                const u32 anchor = j ? (u32)seq[C3_MAXSEQ + j - 1] + 1u : cur;
                                while (p > anchor && p > o && s_in[p - 1] == s_in[p - o - 1]) p--;      // lz4/lz4.c:1019
                                const u32 lit = p - anchor, ml = e - p - MINMATCH;
                                enc = 1 + c1_len_ext_bytes(lit) + lit + 2 + c1_len_ext_bytes(ml);
                                s_seqq[j] = (u16)p;
                            }
                            u32 incl = enc;
                            #pragma unroll
                            for (u32 d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += y; }
                            if (j < nseq) s_seqo[j] = incl - enc;
                            if (lane == 31) s_blk[b] = incl;
                        }
                        c3_bar_emit();
                        if (warp == C3_NVR && lane == 0) s_misc[1] = 0;
                        c3_bar_emit();
                        // pass B: the bytes
                        u32 tile_bytes = 0;
                        for (u32 b = ew; b < nblk; b += C3_NE) {
                            const u32 j = b * 32 + lane;
                            u32 before = 0;
                            for (u32 q = 0; q < b; q++) before += s_blk[q];
                            if (j < nseq) {
                                const u32 p = s_seqq[j], e = (u32)seq[C3_MAXSEQ + j] + 1u, o = seq[2 * C3_MAXSEQ + j];
                                const u32 anchor = j ? (u32)seq[C3_MAXSEQ + j - 1] + 1u : cur;
                                const u32 lit = p - anchor, ml = e - p - MINMATCH;
                                u8* q = dst + op + before + s_seqo[j];
                                *q++ = (u8)(((lit < RUN_MASK ? lit : RUN_MASK) << 4) | (ml < ML_MASK ? ml : ML_MASK));
                                if (lit >= RUN_MASK) q = c1_put_ext(q, lit - RUN_MASK);
                                if (lit >= C3_LONG_LIT) {
                                    const u32 slot = atomicAdd((u32*)&s_misc[1], 1u);
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
                        op += tile_bytes;
                        if (t + 1 == n_tiles && warp == C3_NVR && lane == 0) s_misc[6] = op;
                        if (prof && warp == C3_NVR && lane == 0) {
                            atomicAdd((unsigned long long*)&prof[3], (unsigned long long)(clock64() - q0));
                            atomicAdd((unsigned long long*)&prof[4], 1ull);
                        }
                    }
                }
                __syncthreads();
                if (prof && tid == 0) atomicAdd((unsigned long long*)&prof[6], (unsigned long long)(clock64() - q0));
            }
        }
        // ---- long literal runs of the last tile, then the last literals (lz4/lz4.c:1204-1231), all threads ----
        {
            const u32 nlong = s_misc[1];
            for (u32 e = ltid >> 5; e < nlong; e += C3_LIVE / 32) {
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
            for (u32 i = ltid; i < last; i += C3_LIVE) o[hdr + i] = s_in[cur + i];
            if (tid == 0 && tail_pos) { tail_pos[chunk] = op; tail_len[chunk] = last; }
            op += hdr + last;
            if (tid == 0) csz[chunk] = op;
        }
        __syncthreads();
    }
}

inline cudaError_t c3_launch_compress(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                                      u32 n, u32* ticket, int sm_count, cudaStream_t st, u64* prof = nullptr, const u64* n_dev = nullptr,
                                      u32* tail_pos = nullptr, u32* tail_len = nullptr, const u32* hist = nullptr)
{
    cudaError_t e = cudaFuncSetAttribute(compress_c3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C3Smem::BYTES);   // per device: every launch
    if (e == cudaSuccess && prof) e = cudaFuncSetAttribute(compress_c3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C3Smem::BYTES);
    if (e != cudaSuccess) return e;
    unsigned grid = (unsigned)(sm_count * 2);
    if (grid > n) grid = n;
    e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    if (prof) compress_c3_kernel<true><<<grid, C3_THREADS, C3Smem::BYTES, st>>>(in, in_off, usz, stage, stride, csz, n, ticket, prof, n_dev, tail_pos, tail_len, hist);
    else compress_c3_kernel<false><<<grid, C3_THREADS, C3Smem::BYTES, st>>>(in, in_off, usz, stage, stride, csz, n, ticket, nullptr, n_dev, tail_pos, tail_len, hist);
    return cudaGetLastError();
}

}  // namespace lz4b200
