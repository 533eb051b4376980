// kernels_c1.cuh -- generation-1 compressor: one CTA per chunk (<= 64 KB), the chunk resident in
// shared memory (TMA bulk load), match finding parallel over positions, greedy parse resolved by
// speculative windows.
//
// The reference compressor (lz4/lz4.c:851-1240) is one serial loop per block: hash the current
// position, probe a 16 KB table, compare 4 bytes, extend, emit, continue after the match.  Here the
// loop is taken apart, one 4096-position tile at a time:
//   M  match finding, every position in parallel.  Batches of 256 consecutive positions (one per
//      thread): 13-bit hash of 4 bytes (:698-702), read the u16 position table, prefer an earlier
//      lane of the same warp with the same hash (__match_any_sync), verify the 4 bytes, then insert
//      the position (last writer of a warp wins, :736-765).  One barrier per batch keeps the table a
//      batch behind at most, so a candidate is (nearly) always the latest earlier occurrence.  Unlike
//      the reference EVERY position is inserted, not only the ones the parser visits.
//      Result: s_moff[p] = distance to a verified 4-byte match (0 = none) + a bitmap.
//   P  greedy parse.  "First matched position at or after the end of the previous match, extend,
//      jump" is again a serial chain, but chains started at different positions merge as soon as
//      they hit the same match, so 256 lanes walk 256 windows of 16 positions from a guessed state
//      and entry states are repaired with the same cover / prefix-max rounds as the v1 decoder.
//      A lane compares at most CAP bytes; longer matches are finished by warp 0 in window order.
//   E  emit.  Per-window encoded sizes -> exclusive scan -> every lane writes its sequences (token,
//      literal length, literals, LE16 offset, match length: :1022-1135); literal runs >= 64 bytes are
//      copied by whole warps.
// End-of-block rules (:883-884,:921,:1204-1231): no match starts after n-12 or ends after n-5, inputs
// shorter than 13 bytes are one literal run, the block ends with a literal-only sequence.
#pragma once
#include "common.cuh"
#include "kernels_v1.cuh"      // mbarrier / bulk-copy wrappers

namespace lz4b200 {

constexpr u32 C1_THREADS = 256;
constexpr u32 C1_CHUNK_MAX = 65536;
constexpr u32 C1_TILE = 4096;                    // positions per tile
constexpr u32 C1_WP = 16;                        // positions per parse window
constexpr u32 C1_WP_SHIFT = 4;
constexpr u32 C1_NW = C1_TILE / C1_WP;           // 256 windows = one per thread
constexpr u32 C1_MAXREC = C1_WP / MINMATCH;      // matches that can start inside one window
constexpr u32 C1_CAP = 64;                       // bytes a lane compares before handing the match to warp 0
constexpr u32 C1_LONG_LIT = 64;                  // literal runs this long are copied by a whole warp
constexpr u32 C1_HBITS = 13;
constexpr u32 C1_LT = 512;                       // buckets of the batch-local table
constexpr u32 C1_NONE = 0xFFFFFFFFu;

struct C1Smem {
    static constexpr u32 OFF_IN = 0;                                  // chunk bytes (+ alignment slack)
    static constexpr u32 IN_CAP = C1_CHUNK_MAX + 48;
    static constexpr u32 OFF_TAB = OFF_IN + IN_CAP;                    // u16[1 << HBITS]
    static constexpr u32 OFF_MOFF = OFF_TAB + 2 * (1u << C1_HBITS);    // u16[TILE]
    static constexpr u32 OFF_BITS = OFF_MOFF + 2 * C1_TILE;            // u32[TILE / 32]
    static constexpr u32 OFF_EXIT = OFF_BITS + 4 * (C1_TILE / 32);     // u32[NW]
    static constexpr u32 OFF_NREC = OFF_EXIT + 4 * C1_NW;              // u32[NW]   records written by the last walk
    static constexpr u32 OFF_RSTART = OFF_NREC + 4 * C1_NW;            // u16[NW * MAXREC]
    static constexpr u32 OFF_RLEN = OFF_RSTART + 2 * C1_NW * C1_MAXREC;
    static constexpr u32 OFF_ROFF = OFF_RLEN + 2 * C1_NW * C1_MAXREC;
    static constexpr u32 OFF_CAPOFF = OFF_ROFF + 2 * C1_NW * C1_MAXREC; // u16[NW] offset of a window's capped match (0 none, 0xFFFF resolved)
    static constexpr u32 OFF_HEADQ = OFF_CAPOFF + 2 * C1_NW;           // u16[NW] windows whose long match gets finished this round
    static constexpr u32 OFF_LONGQ = (OFF_HEADQ + 2 * C1_NW + 15) & ~15u;   // uint4[TILE/64 + 2]
    static constexpr u32 NLONG = C1_TILE / C1_LONG_LIT + 2;
    static constexpr u32 OFF_WARP = OFF_LONGQ + 16 * NLONG;            // u32[64] cross-warp scan scratch
    static constexpr u32 OFF_LT = OFF_WARP + 4 * 64;                   // u32[C1_LT] batch-local first-occurrence table
    static constexpr u32 OFF_MISC = OFF_LT + 4 * C1_LT;                // mbarrier (8 B) + scalars
    static constexpr u32 BYTES = OFF_MISC + 64;
};

// 4 bytes at any shared-memory address: two aligned loads and a funnel shift.
__device__ __forceinline__ u32 c1_rd32(const u8* p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const u32* w = reinterpret_cast<const u32*>(a & ~uintptr_t(3));
    return __funnelshift_r(w[0], w[1], (u32)(a & 3) * 8);
}

// common prefix of in[a..] and in[b..] (b < a), at most `cap` bytes, never reading in[limit..] as data
__device__ __forceinline__ u32 c1_count(const u8* in, u32 a, u32 b, u32 limit, u32 cap)
{
    u32 len = 0;
    while (len < cap && a + len + 4 <= limit) {
        const u32 x = c1_rd32(in + a + len) ^ c1_rd32(in + b + len);
        if (x) return len + ((__ffs(x) - 1) >> 3);
        len += 4;
    }
    while (len < cap && a + len < limit && in[a + len] == in[b + len]) len++;
    return len;
}

__device__ __forceinline__ u32 c1_len_ext_bytes(u32 v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }

// write the 255-run tail of a length (value v >= 15 already reduced by the caller: r = v - 15)
__device__ __forceinline__ u8* c1_put_ext(u8* o, u32 r)
{
    for (; r >= 255; r -= 255) *o++ = 255;
    *o++ = (u8)r;
    return o;
}

// block-wide exclusive prefix (max or sum) over one value per thread; `warp_scratch` holds 2*32 u32
template <bool IS_MAX>
__device__ __forceinline__ u32 c1_block_excl(u32 v, u32* warp_scratch, u32& total)
{
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr u32 NWARP = C1_THREADS / 32;
    u32 incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 y = __shfl_up_sync(FULL, incl, d);
        if (lane >= (u32)d) incl = IS_MAX ? (y > incl ? y : incl) : incl + y;
    }
    if (lane == 31) warp_scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 m = lane < NWARP ? warp_scratch[lane] : 0;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 y = __shfl_up_sync(FULL, m, d);
            if (lane >= (u32)d) m = IS_MAX ? (y > m ? y : m) : m + y;
        }
        warp_scratch[32 + lane] = m;
    }
    __syncthreads();
    u32 excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 0;
    if (warp) {
        const u32 y = warp_scratch[32 + warp - 1];
        excl = IS_MAX ? (y > excl ? y : excl) : excl + y;
    }
    total = warp_scratch[32 + NWARP - 1];
    return excl;
}

template <int MODE>
__global__ void __launch_bounds__(C1_THREADS, 2)
compress_c1_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ usz,
                   u8* __restrict__ stage, u64 slot_stride, u64* __restrict__ csz, u32 n_chunks, u32* ticket, u64* prof)
{
    extern __shared__ __align__(128) u8 smem[];
    u8* const s_in_raw = smem + C1Smem::OFF_IN;
    u16* const s_tab = reinterpret_cast<u16*>(smem + C1Smem::OFF_TAB);
    u16* const s_moff = reinterpret_cast<u16*>(smem + C1Smem::OFF_MOFF);
    u32* const s_bits = reinterpret_cast<u32*>(smem + C1Smem::OFF_BITS);
    u32* const s_exit = reinterpret_cast<u32*>(smem + C1Smem::OFF_EXIT);
    u32* const s_nrec = reinterpret_cast<u32*>(smem + C1Smem::OFF_NREC);
    u16* const r_start = reinterpret_cast<u16*>(smem + C1Smem::OFF_RSTART);
    u16* const r_len = reinterpret_cast<u16*>(smem + C1Smem::OFF_RLEN);
    u16* const r_off = reinterpret_cast<u16*>(smem + C1Smem::OFF_ROFF);
    u16* const s_capoff = reinterpret_cast<u16*>(smem + C1Smem::OFF_CAPOFF);
    u16* const s_headq = reinterpret_cast<u16*>(smem + C1Smem::OFF_HEADQ);
    uint4* const s_longq = reinterpret_cast<uint4*>(smem + C1Smem::OFF_LONGQ);
    u32* const s_warp = reinterpret_cast<u32*>(smem + C1Smem::OFF_WARP);
    u32* const s_lt = reinterpret_cast<u32*>(smem + C1Smem::OFF_LT);
    u64* const s_bar = reinterpret_cast<u64*>(smem + C1Smem::OFF_MISC);
    u32* const s_misc = reinterpret_cast<u32*>(smem + C1Smem::OFF_MISC + 16);   // [0] chunk [1] long-literal count [3] head count

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr u32 NWARP = C1_THREADS / 32;
    u32 phase = 0;
    if (tid == 0) { mbar_init(s_bar, 1); s_misc[3] = 0; }
    __syncthreads();
    long long t_prev = prof ? clock64() : 0;
    u64 t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define C1_TICK(k) do { if (prof && tid == 0) { long long t_ = clock64(); t_acc[k] += (u64)(t_ - t_prev); t_prev = t_; } } while (0)

    for (;;) {
        if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 chunk = s_misc[0];
        if (chunk >= n_chunks) break;
        const u8* src = in + in_off[chunk];
        const u32 n = (u32)usz[chunk];
        u8* dst = stage + (u64)chunk * slot_stride;

        // ---- stage the chunk: bulk copy of the 16-byte aligned interior, edges by hand; clear the table ----
        const u32 head = (u32)(reinterpret_cast<uintptr_t>(src) & 15);
        u8* const s_in = s_in_raw + head;
        {
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
            for (u32 i = tid; i < (2u << C1_HBITS) / 16; i += C1_THREADS) t4[i] = make_uint4(0, 0, 0, 0);
            for (u32 i = tid; i < C1_LT; i += C1_THREADS) s_lt[i] = 0;
            if (bulk) { mbar_wait(s_bar, phase); phase ^= 1; }
        }
        __syncthreads();
        C1_TICK(0);

        u32 cur = 0;                                    // chain state: end of the last match (= literal anchor)
        u32 op = 0;                                     // bytes emitted so far
        if (n >= MIN_LENGTH) {
            const u32 mfl = n - MFLIMIT;                // last position a match may start at
            const u32 mlimit = n - LASTLITERALS;        // matches end at or before this
            for (u32 t0 = 0; t0 <= mfl; t0 += C1_TILE) {
                // ---- M: match finding over the tile, 256 positions per batch ----
                for (u32 b = 0; b < C1_TILE; b += C1_THREADS) {
                    const u32 p = t0 + b + tid;
                    const bool valid = p <= mfl;
                    u32 v = 0, h = 0x10000u + lane, cand = 0;
                    if (valid) { v = c1_rd32(s_in + p); h = hash13(v); cand = s_tab[h]; }
                    // batch-local first occurrence: the main table is a batch behind, so a repeat at distance < 256 inside this
                    // batch would be missed (measured on real binaries / source text: +6..7 % size).  Every position of the
                    // batch posts itself to a 512-bucket table with atomicMax(tag | ~pos): the tag makes entries of older
                    // batches lose without clearing, the inverted position makes the FIRST occurrence of the batch win.
                    const u32 tag = ((t0 + b) / C1_THREADS + 1u) << 16;
                    if (valid) atomicMax(&s_lt[h & (C1_LT - 1)], tag | (0xFFFFu - p));
                    __syncthreads();
                    bool ok = false;
                    {
                        u32 near = 0;                         // same 4 bytes d lanes earlier in this warp (short periods)
                        if (MODE == 0) {
                            const unsigned peers = __match_any_sync(FULL, h);
                            const unsigned lower = peers & ((1u << lane) - 1u);
                            if (lower) near = lane - (31 - __clz(lower));
                        } else if (MODE == 1) {
                            #pragma unroll
                            for (u32 d : {16u, 8u, 4u, 3u, 2u, 1u}) {
                                const u32 y = __shfl_up_sync(FULL, v, d);
                                if (lane >= d && y == v) near = d;
                            }
                        }
                        if (valid) {
                            const u32 lt = s_lt[h & (C1_LT - 1)];
                            const u32 q = 0xFFFFu - (lt & 0xFFFFu);               // first position of this batch in my bucket
                            const bool local = (lt >> 16) == (tag >> 16) && q < p;
                            const u32 main_cand = cand;
                            if (near) cand = p - near; else if (local) cand = q;
                            ok = cand < p && c1_rd32(s_in + cand) == v;
                            if (!ok && cand != main_cand) { cand = main_cand; ok = cand < p && c1_rd32(s_in + cand) == v; }
                            s_tab[h] = (u16)p;                                     // racing lanes: any winner is a valid candidate
                        }
                    }
                    s_moff[b + tid] = ok ? (u16)(p - cand) : (u16)0;
                    const unsigned hit = __ballot_sync(FULL, ok);
                    if (lane == 0) s_bits[(b + tid) >> 5] = hit;
                    __syncthreads();
                }

                C1_TICK(1);
                // ---- P: greedy parse by speculative windows ----
                const u32 wstart = t0 + tid * C1_WP, wend = wstart + C1_WP;
                const u32 wbits = (s_bits[tid >> 1] >> ((tid & 1) * 16)) & 0xFFFFu;
                u32 entry = cur, nrec = 0, exitv = cur;
                bool capped = false;
                bool need_walk = true;
                u32 round_count = 0;
                for (u32 round = 0;; round++) {
                    round_count++;
                    if (need_walk) {
                        u32 pos = entry > wstart ? entry : wstart;
                        nrec = 0; capped = false; exitv = entry;
                        while (pos < wend) {
                            const u32 m = wbits >> (pos - wstart);
                            if (!m) break;
                            const u32 p = pos + (__ffs(m) - 1);
                            const u32 off = s_moff[p - t0];
                            const u32 ext = c1_count(s_in, p + MINMATCH, p - off + MINMATCH, mlimit, C1_CAP);
                            const u32 len = MINMATCH + ext;
                            r_start[tid * C1_MAXREC + nrec] = (u16)p;
                            r_len[tid * C1_MAXREC + nrec] = (u16)len;
                            r_off[tid * C1_MAXREC + nrec] = (u16)off;
                            nrec++;
                            pos = p + len; exitv = pos;
                            if (ext >= C1_CAP) { capped = true; break; }
                        }
                        s_exit[tid] = exitv;
                        s_nrec[tid] = nrec;
                    }
                    // long matches.  A run makes EVERY window inside it report a capped match; only the first window of
                    // a stretch of capped windows (the head) is worth finishing, the others are skipped by the head's
                    // cover one round later -- or become heads themselves if they survive.
                    const u32 my_off = capped ? (u32)r_off[tid * C1_MAXREC + nrec - 1] : 0u;
                    s_capoff[tid] = (u16)my_off;
                    bool uncertain = false;
                    if (__syncthreads_or(capped)) {
                        if (capped && (tid == 0 || s_capoff[tid - 1] == 0)) s_headq[atomicAdd(&s_misc[3], 1u)] = (u16)tid;
                        __syncthreads();
                        const u32 nheads = s_misc[3];
                        for (u32 e = warp; e < nheads; e += NWARP) {           // one warp per head, 128 bytes per step
                            const u32 w = s_headq[e];
                            const u32 k = w * C1_MAXREC + s_nrec[w] - 1;         // the capped match is the walk's last record
                            const u32 p = r_start[k], off = r_off[k];
                            u32 len = r_len[k];
                            for (;;) {
                                const u32 a = p + len + lane * 4;
                                u32 c = 0;
                                if (a + 4 <= mlimit) {
                                    const u32 x = c1_rd32(s_in + a) ^ c1_rd32(s_in + a - off);
                                    c = x ? ((__ffs(x) - 1) >> 3) : 4;
                                } else {
                                    while (a + c < mlimit && s_in[a + c] == s_in[a + c - off]) c++;
                                }
                                const unsigned full = __ballot_sync(FULL, c == 4);
                                if (full == FULL) { len += 128; continue; }
                                const int fl = __ffs(~full) - 1;                // first lane that stopped
                                len += fl * 4 + __shfl_sync(FULL, c, fl);
                                break;
                            }
                            if (lane == 0) { r_len[k] = (u16)len; s_exit[w] = p + len; s_capoff[w] = 0xFFFFu; }   // resolved
                        }
                        __syncthreads();
                        if (tid == 0) s_misc[3] = 0;
                        if (capped) { if (s_capoff[tid] == 0xFFFFu) capped = false; else uncertain = true; }
                        exitv = s_exit[tid];
                    }
                    // producers claim the window their chain state lands in; prefix-max picks the furthest
                    u32 key = 0;
                    if (nrec) {
                        u32 cover = (exitv - t0) >> C1_WP_SHIFT;
                        if (cover > C1_NW) cover = C1_NW;
                        if (cover < tid) cover = tid;
                        key = ((cover + 1) << 16) | tid;         // +1: window 0 landing in window 0 must not look like "no producer"
                    }
                    u32 total_key;
                    const u32 best = c1_block_excl<true>(key, s_warp, total_key);
                    u32 e_new = cur;
                    if (best) { const u32 x = s_exit[best & 0xFFFFu]; if (x > e_new) e_new = x; }
                    const u32 eff_new = e_new > wstart ? e_new : wstart;
                    const u32 eff_old = entry > wstart ? entry : wstart;
                    need_walk = eff_new != eff_old || (uncertain && eff_new < wend) || (nrec && eff_new >= wend);
                    entry = e_new;
                    if (uncertain && eff_new >= wend) { capped = false; }
                    if (!__syncthreads_or(need_walk) || round > 2 * C1_NW + 8) {   // the round limit cannot trigger; keeps the loop finite
                        // converged: chain state after the tile = exit of the furthest producer
                        if (total_key) { const u32 x = s_exit[total_key & 0xFFFFu]; if (x > cur) cur = x; }
                        break;
                    }
                }

                C1_TICK(2);
                if (prof && tid == 0) t_acc[5] += round_count;
                // ---- E: emit the tile's sequences ----
                // pass 1: backward catch-up (lz4/lz4.c:1019) and encoded size per window
                u32 enc = 0;
                {
                    u32 anchor = entry;
                    for (u32 k = 0; k < nrec; k++) {
                        u32 p = r_start[tid * C1_MAXREC + k], len = r_len[tid * C1_MAXREC + k];
                        const u32 off = r_off[tid * C1_MAXREC + k];
                        while (p > anchor && p - off > 0 && s_in[p - 1] == s_in[p - off - 1]) { p--; len++; }
                        r_start[tid * C1_MAXREC + k] = (u16)p; r_len[tid * C1_MAXREC + k] = (u16)len;
                        const u32 lit = p - anchor;
                        enc += 1 + c1_len_ext_bytes(lit) + lit + 2 + c1_len_ext_bytes(len - MINMATCH);
                        anchor = p + len;
                    }
                }
                u32 tile_bytes;
                const u32 my_o = op + c1_block_excl<false>(enc, s_warp, tile_bytes);
                if (tid == 0) s_misc[1] = 0;
                __syncthreads();
                // pass 2: write
                {
                    u32 anchor = entry;
                    u8* o = dst + my_o;
                    for (u32 k = 0; k < nrec; k++) {
                        const u32 p = r_start[tid * C1_MAXREC + k], len = r_len[tid * C1_MAXREC + k];
                        const u32 off = r_off[tid * C1_MAXREC + k];
                        const u32 lit = p - anchor, ml = len - MINMATCH;
                        *o++ = (u8)(((lit < RUN_MASK ? lit : RUN_MASK) << 4) | (ml < ML_MASK ? ml : ML_MASK));
                        if (lit >= RUN_MASK) o = c1_put_ext(o, lit - RUN_MASK);
                        if (lit >= C1_LONG_LIT) {
                            const u32 slot = atomicAdd(&s_misc[1], 1u);
                            s_longq[slot] = make_uint4(anchor, (u32)(o - dst), lit, 0);
                        } else {
                            for (u32 i = 0; i < lit; i++) o[i] = s_in[anchor + i];
                        }
                        o += lit;
                        *o++ = (u8)off; *o++ = (u8)(off >> 8);
                        if (ml >= ML_MASK) o = c1_put_ext(o, ml - ML_MASK);
                        anchor = p + len;
                    }
                }
                __syncthreads();
                {
                    const u32 nlong = s_misc[1];
                    for (u32 e = warp; e < nlong; e += NWARP) {
                        const uint4 q = s_longq[e];
                        for (u32 i = lane; i < q.z; i += 32) dst[q.y + i] = s_in[q.x + i];
                    }
                }
                op += tile_bytes;
                __syncthreads();
                C1_TICK(3);
            }
        }
        // ---- last literals (lz4/lz4.c:1204-1231) ----
        {
            const u32 last = n - cur;
            u8* o = dst + op;
            const u32 hdr = 1 + c1_len_ext_bytes(last);
            if (tid == 0) {
                *o = (u8)((last < RUN_MASK ? last : RUN_MASK) << 4);
                if (last >= RUN_MASK) c1_put_ext(o + 1, last - RUN_MASK);
            }
            for (u32 i = tid; i < last; i += C1_THREADS) o[hdr + i] = s_in[cur + i];
            op += hdr + last;
        }
        if (tid == 0) csz[chunk] = op;
        __syncthreads();
        C1_TICK(4);
    }
    if (prof && tid == 0)
        for (int k = 0; k < 8; k++) atomicAdd(reinterpret_cast<unsigned long long*>(prof) + 8 + k, (unsigned long long)t_acc[k]);
#undef C1_TICK
}

inline bool c1_compress_supported(size_t max_chunk) { return max_chunk <= C1_CHUNK_MAX; }

template <int MODE>
inline cudaError_t c1_launch_mode(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                                      u32 n, u32* ticket, u64* prof, int sm_count, size_t, cudaStream_t st)
{
    // the opt-in is per device (a process may hold contexts on several GPUs): set it at every launch, it is cheap
    {
        cudaError_t e = cudaFuncSetAttribute(compress_c1_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)C1Smem::BYTES);
        if (e != cudaSuccess) return e;
    }
    unsigned grid = (unsigned)(sm_count * 2);
    if (grid > n) grid = n;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    compress_c1_kernel<MODE><<<grid, C1_THREADS, C1Smem::BYTES, st>>>(in, in_off, usz, stage, stride, csz, n, ticket, prof);
    return cudaGetLastError();
}

// mode = intra-warp candidates: 0 match_any, 1 shuffle distances, 2 none (default: the batch-local table covers them); a per-context option
inline cudaError_t c1_launch_compress(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                                      u32 n, u32* ticket, u64* prof, int sm_count, size_t mc, cudaStream_t st, int mode = 2)
{
    switch (mode) {
    case 1:  return c1_launch_mode<1>(in, in_off, usz, stage, stride, csz, n, ticket, prof, sm_count, mc, st);
    case 2:  return c1_launch_mode<2>(in, in_off, usz, stage, stride, csz, n, ticket, prof, sm_count, mc, st);
    default: return c1_launch_mode<0>(in, in_off, usz, stage, stride, csz, n, ticket, prof, sm_count, mc, st);
    }
}

}  // namespace lz4b200
