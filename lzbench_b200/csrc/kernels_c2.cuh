// kernels_c2.cuh -- generation-2 compressor: kernels_c1.cuh with the two halves of a tile overlapped.
//
// The c1 profile (profiles/r01_compress_c1_ncu_*) shows a barrier-bound kernel: 35 % of stall samples are
// barrier waits, 16 warps per SM, IPC 1.7 -- match finding (M) and parse + emit (P, E) alternate, each phase a
// chain of short steps separated by CTA barriers.  The phases of DIFFERENT tiles are independent except through
// the hash table (owned by M) and the match arrays of a tile (written by M, read by P/E), so c2 runs them on
// two warp groups of one 512-thread CTA: group A produces tile t+1 (M) while group B consumes tile t (P, E),
// hand-over through double-buffered match arrays and four named barriers (full / empty per buffer); barriers
// inside a phase are 256-thread named barriers.  Same algorithm, same output as c1.
#pragma once
#include "common.cuh"
#include "kernels_c1.cuh"

namespace lz4b200 {

constexpr u32 C2_THREADS = 512;
constexpr u32 C2_GROUP = 256;                    // threads per warp group

struct C2Smem {
    static constexpr u32 OFF_IN = 0;
    static constexpr u32 IN_CAP = C1_CHUNK_MAX + 48;
    static constexpr u32 OFF_TAB = OFF_IN + IN_CAP;                    // u16[1 << HBITS]
    static constexpr u32 OFF_MOFF = OFF_TAB + 2 * (1u << C1_HBITS);    // u16[2][TILE]
    static constexpr u32 OFF_BITS = OFF_MOFF + 2 * 2 * C1_TILE;        // u32[2][TILE / 32]
    static constexpr u32 OFF_EXIT = OFF_BITS + 2 * 4 * (C1_TILE / 32); // u32[NW]
    static constexpr u32 OFF_NREC = OFF_EXIT + 4 * C1_NW;              // u32[NW]
    static constexpr u32 OFF_RSTART = OFF_NREC + 4 * C1_NW;            // u16[NW * MAXREC]
    static constexpr u32 OFF_RLEN = OFF_RSTART + 2 * C1_NW * C1_MAXREC;
    static constexpr u32 OFF_ROFF = OFF_RLEN + 2 * C1_NW * C1_MAXREC;
    static constexpr u32 OFF_CAPOFF = OFF_ROFF + 2 * C1_NW * C1_MAXREC;
    static constexpr u32 OFF_HEADQ = OFF_CAPOFF + 2 * C1_NW;
    static constexpr u32 OFF_LONGQ = (OFF_HEADQ + 2 * C1_NW + 15) & ~15u;
    static constexpr u32 NLONG = C1_TILE / C1_LONG_LIT + 2;
    static constexpr u32 OFF_WARP = OFF_LONGQ + 16 * NLONG;            // u32[64]
    static constexpr u32 OFF_LT = OFF_WARP + 4 * 64;                   // u32[C1_LT] batch-local first-occurrence table
    static constexpr u32 OFF_MISC = OFF_LT + 4 * C1_LT;                // mbarrier (8 B) + scalars
    static constexpr u32 BYTES = OFF_MISC + 64;
};

// named barriers: 0 is __syncthreads; 1,2 full[buffer]; 3,4 empty[buffer]; 5 inside group A; 6,7 inside group B
template <u32 ID, u32 COUNT> __device__ __forceinline__ void c2_bar_sync_c() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <u32 ID, u32 COUNT> __device__ __forceinline__ void c2_bar_arrive_c() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
// buffer-indexed forms with immediate ids (a register id would make ptxas reserve all 16 barriers)
__device__ __forceinline__ void c2_bar_sync(u32 base, u32 buf)
{
    if (base == 1) { if (buf) c2_bar_sync_c<2, 512>(); else c2_bar_sync_c<1, 512>(); }
    else           { if (buf) c2_bar_sync_c<4, 512>(); else c2_bar_sync_c<3, 512>(); }
}
__device__ __forceinline__ void c2_bar_arrive(u32 base, u32 buf)
{
    if (base == 1) { if (buf) c2_bar_arrive_c<2, 512>(); else c2_bar_arrive_c<1, 512>(); }
    else           { if (buf) c2_bar_arrive_c<4, 512>(); else c2_bar_arrive_c<3, 512>(); }
}
__device__ __forceinline__ bool c2_bar_or(u32 id, u32 count, bool pred)
{
    u32 r;
    (void)id; (void)count;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, 7, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r) : "r"((u32)pred) : "memory");
    return r != 0;
}
#define C2_BAR_A() c2_bar_sync_c<5, 256>()
#define C2_BAR_B() c2_bar_sync_c<6, 256>()
#define C2_OR_B(x) c2_bar_or(7, C2_GROUP, (x))

// exclusive prefix (max or sum) over the 256 threads of group B
template <bool IS_MAX>
__device__ __forceinline__ u32 c2_block_excl(u32 v, u32* warp_scratch, u32& total)
{
    const u32 lane = threadIdx.x & 31, warp = (threadIdx.x & (C2_GROUP - 1)) >> 5;
    constexpr u32 NWARP = C2_GROUP / 32;
    u32 incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 y = __shfl_up_sync(FULL, incl, d);
        if (lane >= (u32)d) incl = IS_MAX ? (y > incl ? y : incl) : incl + y;
    }
    if (lane == 31) warp_scratch[warp] = incl;
    C2_BAR_B();
    if (warp == 0) {
        u32 m = lane < NWARP ? warp_scratch[lane] : 0;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 y = __shfl_up_sync(FULL, m, d);
            if (lane >= (u32)d) m = IS_MAX ? (y > m ? y : m) : m + y;
        }
        warp_scratch[32 + lane] = m;
    }
    C2_BAR_B();
    u32 excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 0;
    if (warp) {
        const u32 y = warp_scratch[32 + warp - 1];
        excl = IS_MAX ? (y > excl ? y : excl) : excl + y;
    }
    total = warp_scratch[32 + NWARP - 1];
    return excl;
}

template <int MODE, int ACC>
__global__ void __launch_bounds__(C2_THREADS, 2)
compress_c2_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ usz,
                   u8* __restrict__ stage, u64 slot_stride, u64* __restrict__ csz, u32 n_chunks_arg, u32* ticket,
                   const u64* __restrict__ n_dev, u32* __restrict__ tail_pos, u32* __restrict__ tail_len,
                   const u32* __restrict__ hist_arr, u64* __restrict__ prof)
{
    // segmented mode (segmented.cuh): the job count lives in device memory and every job also reports where its
    // last_literals record starts and how many literals it holds
    const u32 n_chunks = n_dev ? (u32)*n_dev : n_chunks_arg;
    extern __shared__ __align__(128) u8 smem[];
    u8* const s_in_raw = smem + C2Smem::OFF_IN;
    u16* const s_tab = reinterpret_cast<u16*>(smem + C2Smem::OFF_TAB);
    u16* const s_moff2 = reinterpret_cast<u16*>(smem + C2Smem::OFF_MOFF);
    u32* const s_bits2 = reinterpret_cast<u32*>(smem + C2Smem::OFF_BITS);
    u32* const s_exit = reinterpret_cast<u32*>(smem + C2Smem::OFF_EXIT);
    u32* const s_nrec = reinterpret_cast<u32*>(smem + C2Smem::OFF_NREC);
    u16* const r_start = reinterpret_cast<u16*>(smem + C2Smem::OFF_RSTART);
    u16* const r_len = reinterpret_cast<u16*>(smem + C2Smem::OFF_RLEN);
    u16* const r_off = reinterpret_cast<u16*>(smem + C2Smem::OFF_ROFF);
    u16* const s_capoff = reinterpret_cast<u16*>(smem + C2Smem::OFF_CAPOFF);
    u16* const s_headq = reinterpret_cast<u16*>(smem + C2Smem::OFF_HEADQ);
    uint4* const s_longq = reinterpret_cast<uint4*>(smem + C2Smem::OFF_LONGQ);
    u32* const s_warp = reinterpret_cast<u32*>(smem + C2Smem::OFF_WARP);
    u32* const s_lt = reinterpret_cast<u32*>(smem + C2Smem::OFF_LT);
    u64* const s_bar = reinterpret_cast<u64*>(smem + C2Smem::OFF_MISC);
    u32* const s_misc = reinterpret_cast<u32*>(smem + C2Smem::OFF_MISC + 16);   // [0] chunk [1] long-literal count [3] head count [4] cur [5] op

    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 gtid = tid & (C2_GROUP - 1), grp = tid / C2_GROUP, warp = gtid >> 5;
    constexpr u32 NWARP = C2_GROUP / 32;
    u32 phase = 0;
    if (tid == 0) { mbar_init(s_bar, 1); s_misc[3] = 0; }
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
        // ---- stage the chunk (TMA bulk copy of the aligned interior, edges by hand), clear the table ----
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
            for (u32 i = tid; i < (2u << C1_HBITS) / 16; i += C2_THREADS) t4[i] = make_uint4(0, 0, 0, 0);
            for (u32 i = tid; i < C1_LT; i += C2_THREADS) s_lt[i] = 0;
            if (tid == 0) { s_misc[4] = hist_arr ? (hist_arr[chunk] & 0x7FFFFFFFu) : 0u; s_misc[5] = 0; }   // chain starts after the history
            if (bulk) { mbar_wait(s_bar, phase); phase ^= 1; }
        }
        __syncthreads();
        if (prof && tid == 0) atomicAdd((unsigned long long*)&prof[5], (unsigned long long)(clock64() - pt0));

        if (n >= MIN_LENGTH) {
            // a job that is not the end of its chunk (segmented mode, bit 31 of hist_arr) may match right up to its
            // last byte: the end-of-block rules (lz4/lz4.c:883-884) bind only the real end of the block
            const bool inner = hist_arr && (hist_arr[chunk] >> 31);
            const u32 mfl = n - (inner ? MINMATCH : MFLIMIT);      // last position a match may start at
            const u32 mlimit = n - (inner ? 0u : LASTLITERALS);    // matches end at or before this
            const u32 n_tiles = mfl / C1_TILE + 1;
            if (grp == 0) {
                // ======== group A: match finding, one tile ahead of group B ========
                for (u32 t = 0; t < n_tiles; t++) {
                    const u32 t0 = t * C1_TILE, buf = t & 1;
                    u16* const mo = s_moff2 + buf * C1_TILE;
                    u32* const bi = s_bits2 + buf * (C1_TILE / 32);
                    const long long a0 = prof ? clock64() : 0;
                    if (t >= 2) c2_bar_sync(3, buf);          // group B has released this buffer
                    const long long a1 = prof ? clock64() : 0;
                    if (ACC == 1) {
                    // ---- M: match finding over the tile, 256 positions per batch ----
                        for (u32 b = 0; b < C1_TILE; b += C2_GROUP) {
                            const u32 p = t0 + b + gtid;
                            const bool valid = p <= mfl;
                            u32 v = 0, h = 0x10000u + lane, cand = 0;
                            if (valid) { v = c1_rd32(s_in + p); h = hash13(v); cand = s_tab[h]; }
                            // batch-local first occurrence: the main table is a batch behind, so a repeat at distance < 256 inside this
                            // batch would be missed (measured on real binaries / source text: +6..7 % size).  Every position of the
                            // batch posts itself to a 512-bucket table with atomicMax(tag | ~pos): the tag makes entries of older
                            // batches lose without clearing, the inverted position makes the FIRST occurrence of the batch win.
                            const u32 tag = ((t0 + b) / C2_GROUP + 1u) << 16;
                            if (valid) atomicMax(&s_lt[h & (C1_LT - 1)], tag | (0xFFFFu - p));
                            C2_BAR_A();
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
                            mo[b + gtid] = ok ? (u16)(p - cand) : (u16)0;
                            const unsigned hit = __ballot_sync(FULL, ok);
                            if (lane == 0) bi[(b + gtid) >> 5] = hit;
                            C2_BAR_A();
                        }
    
                    } else {
                        // ---- M, lz4fast levels (LZ4_compress_fast's acceleration, lz4/lz4.c:1352-1375): every position is still
                        //      posted and inserted (cheap), but only every ACC-th one is PROBED (the expensive part: candidate
                        //      choice, verification) -- the same density of match attempts as the reference's constant skip
                        //      step, on a fixed grid.  A batch is 256 * ACC positions, one probe per thread.
                        for (u32 b = 0; b < C1_TILE; b += C2_GROUP * ACC) {
                            const u32 p = t0 + b + gtid * ACC;              // the probed position
                            const bool valid = p <= mfl;
                            const u32 tag = ((t0 + b) / (C2_GROUP * ACC) + 1u) << 16;
                            u32 v = 0, h = 0, cand = 0;
                            #pragma unroll
                            for (u32 i = 0; i < (u32)ACC; i++) {
                                if (p + i <= mfl) {
                                    const u32 vi = c1_rd32(s_in + p + i), hi = hash13(vi);
                                    if (i == 0) { v = vi; h = hi; cand = s_tab[hi]; }
                                    atomicMax(&s_lt[hi & (C1_LT - 1)], tag | (0xFFFFu - (p + i)));
                                }
                            }
                            C2_BAR_A();
                            bool ok = false;
                            if (valid) {
                                const u32 lt = s_lt[h & (C1_LT - 1)];
                                const u32 q = 0xFFFFu - (lt & 0xFFFFu);               // first position of this batch in my bucket
                                const bool local = (lt >> 16) == (tag >> 16) && q < p;
                                const u32 main_cand = cand;
                                if (local) cand = q;
                                ok = cand < p && c1_rd32(s_in + cand) == v;
                                if (!ok && cand != main_cand) { cand = main_cand; ok = cand < p && c1_rd32(s_in + cand) == v; }
                            }
                            #pragma unroll
                            for (u32 i = 0; i < (u32)ACC; i++)
                                if (p + i <= mfl) s_tab[hash13(c1_rd32(s_in + p + i))] = (u16)(p + i);
                            mo[b + gtid * ACC] = ok ? (u16)(p - cand) : (u16)0;
                            const unsigned hit = __ballot_sync(FULL, ok);
                            if (lane < (u32)ACC) {                               // the warp's 32 probes cover ACC words of the bit map
                                constexpr u32 PER = 32 / ACC;
                                const u32 sub = (hit >> (lane * PER)) & ((1u << PER) - 1u);
                                u32 word = 0;
                                #pragma unroll
                                for (u32 k = 0; k < PER; k++) word |= ((sub >> k) & 1u) << (k * ACC);
                                bi[(b >> 5) + warp * ACC + lane] = word;
                            }
                            C2_BAR_A();
                        }
                    }

                    __threadfence_block();
                    if (prof && gtid == 0) { atomicAdd((unsigned long long*)&prof[1], (unsigned long long)(a1 - a0)); atomicAdd((unsigned long long*)&prof[0], (unsigned long long)(clock64() - a1)); }
                    c2_bar_arrive(1, buf);                    // tile t is ready
                }
            } else {
                // ======== group B: parse + emit ========
                u32 cur = hist_arr ? (hist_arr[chunk] & 0x7FFFFFFFu) : 0u;  // chain state: end of the last match (= literal anchor)
                u32 op = 0;                                 // bytes emitted so far
                for (u32 t = 0; t < n_tiles; t++) {
                    const u32 t0 = t * C1_TILE, buf = t & 1;
                    const u16* const mo = s_moff2 + buf * C1_TILE;
                    const u32* const bi = s_bits2 + buf * (C1_TILE / 32);
                    const long long b0 = prof ? clock64() : 0;
                    c2_bar_sync(1, buf);                      // wait for tile t
                    const long long b1 = prof ? clock64() : 0;
                    // ---- P: greedy parse by speculative windows ----
                    const u32 wstart = t0 + gtid * C1_WP, wend = wstart + C1_WP;
                    const u32 wbits = (bi[gtid >> 1] >> ((gtid & 1) * 16)) & 0xFFFFu;
                    u32 entry = cur, nrec = 0, exitv = cur;
                    bool capped = false;
                    bool need_walk = true;
                
                    for (u32 round = 0;; round++) {
                    
                        if (need_walk) {
                            u32 pos = entry > wstart ? entry : wstart;
                            nrec = 0; capped = false; exitv = entry;
                            while (pos < wend) {
                                const u32 m = wbits >> (pos - wstart);
                                if (!m) break;
                                const u32 p = pos + (__ffs(m) - 1);
                                const u32 off = mo[p - t0];
                                const u32 ext = c1_count(s_in, p + MINMATCH, p - off + MINMATCH, mlimit, C1_CAP);
                                const u32 len = MINMATCH + ext;
                                r_start[gtid * C1_MAXREC + nrec] = (u16)p;
                                r_len[gtid * C1_MAXREC + nrec] = (u16)len;
                                r_off[gtid * C1_MAXREC + nrec] = (u16)off;
                                nrec++;
                                pos = p + len; exitv = pos;
                                if (ext >= C1_CAP) { capped = true; break; }
                            }
                            s_exit[gtid] = exitv;
                            s_nrec[gtid] = nrec;
                        }
                        // long matches.  A run makes EVERY window inside it report a capped match; only the first window of
                        // a stretch of capped windows (the head) is worth finishing, the others are skipped by the head's
                        // cover one round later -- or become heads themselves if they survive.
                        const u32 my_off = capped ? (u32)r_off[gtid * C1_MAXREC + nrec - 1] : 0u;
                        s_capoff[gtid] = (u16)my_off;
                        bool uncertain = false;
                        if (C2_OR_B(capped)) {
                            if (capped && (gtid == 0 || s_capoff[gtid - 1] == 0)) s_headq[atomicAdd(&s_misc[3], 1u)] = (u16)gtid;
                            C2_BAR_B();
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
                            C2_BAR_B();
                            if (gtid == 0) s_misc[3] = 0;
                            if (capped) { if (s_capoff[gtid] == 0xFFFFu) capped = false; else uncertain = true; }
                            exitv = s_exit[gtid];
                        }
                        // producers claim the window their chain state lands in; prefix-max picks the furthest
                        u32 key = 0;
                        if (nrec) {
                            u32 cover = (exitv - t0) >> C1_WP_SHIFT;
                            if (cover > C1_NW) cover = C1_NW;
                            if (cover < gtid) cover = gtid;
                            key = ((cover + 1) << 16) | gtid;         // +1: window 0 landing in window 0 must not look like "no producer"
                        }
                        u32 total_key;
                        const u32 best = c2_block_excl<true>(key, s_warp, total_key);
                        u32 e_new = cur;
                        if (best) { const u32 x = s_exit[best & 0xFFFFu]; if (x > e_new) e_new = x; }
                        const u32 eff_new = e_new > wstart ? e_new : wstart;
                        const u32 eff_old = entry > wstart ? entry : wstart;
                        need_walk = eff_new != eff_old || (uncertain && eff_new < wend) || (nrec && eff_new >= wend);
                        entry = e_new;
                        if (uncertain && eff_new >= wend) { capped = false; }
                        if (!C2_OR_B(need_walk) || round > 2 * C1_NW + 8) {   // the round limit cannot trigger; keeps the loop finite
                            // converged: chain state after the tile = exit of the furthest producer
                            if (total_key) { const u32 x = s_exit[total_key & 0xFFFFu]; if (x > cur) cur = x; }
                            if (prof && gtid == 0) atomicAdd((unsigned long long*)&prof[7], (unsigned long long)(round + 1));
                            break;
                        }
                    }

                    const long long b2 = prof ? clock64() : 0;
                    // ---- E: emit the tile's sequences ----
                    // pass 1: backward catch-up (lz4/lz4.c:1019) and encoded size per window
                    u32 enc = 0;
                    {
                        u32 anchor = entry;
                        for (u32 k = 0; k < nrec; k++) {
                            u32 p = r_start[gtid * C1_MAXREC + k], len = r_len[gtid * C1_MAXREC + k];
                            const u32 off = r_off[gtid * C1_MAXREC + k];
                            while (p > anchor && p - off > 0 && s_in[p - 1] == s_in[p - off - 1]) { p--; len++; }
                            r_start[gtid * C1_MAXREC + k] = (u16)p; r_len[gtid * C1_MAXREC + k] = (u16)len;
                            const u32 lit = p - anchor;
                            enc += 1 + c1_len_ext_bytes(lit) + lit + 2 + c1_len_ext_bytes(len - MINMATCH);
                            anchor = p + len;
                        }
                    }
                    u32 tile_bytes;
                    const u32 my_o = op + c2_block_excl<false>(enc, s_warp, tile_bytes);
                    if (gtid == 0) s_misc[1] = 0;
                    C2_BAR_B();
                    // pass 2: write
                    {
                        u32 anchor = entry;
                        u8* o = dst + my_o;
                        for (u32 k = 0; k < nrec; k++) {
                            const u32 p = r_start[gtid * C1_MAXREC + k], len = r_len[gtid * C1_MAXREC + k];
                            const u32 off = r_off[gtid * C1_MAXREC + k];
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
                    C2_BAR_B();
                    {
                        const u32 nlong = s_misc[1];
                        for (u32 e = warp; e < nlong; e += NWARP) {
                            const uint4 q = s_longq[e];
                            for (u32 i = lane; i < q.z; i += 32) dst[q.y + i] = s_in[q.x + i];
                        }
                    }
                    op += tile_bytes;
                    C2_BAR_B();
                    if (prof && gtid == 0) { atomicAdd((unsigned long long*)&prof[2], (unsigned long long)(b1 - b0)); atomicAdd((unsigned long long*)&prof[3], (unsigned long long)(b2 - b1)); atomicAdd((unsigned long long*)&prof[4], (unsigned long long)(clock64() - b2)); }
                    if (t + 2 < n_tiles) c2_bar_arrive(3, buf);   // buffer free for tile t + 2
                }
                if (gtid == 0) { s_misc[4] = cur; s_misc[5] = op; }
            }
        }
        __syncthreads();
        // ---- last literals (lz4/lz4.c:1204-1231), all 512 threads ----
        {
            const u32 cur = s_misc[4];
            u32 op = s_misc[5];
            const u32 last = n - cur;
            u8* o = dst + op;
            const u32 hdr = 1 + c1_len_ext_bytes(last);
            if (tid == 0) {
                *o = (u8)((last < RUN_MASK ? last : RUN_MASK) << 4);
                if (last >= RUN_MASK) c1_put_ext(o + 1, last - RUN_MASK);
            }
            for (u32 i = tid; i < last; i += C2_THREADS) o[hdr + i] = s_in[cur + i];
            if (tid == 0 && tail_pos) { tail_pos[chunk] = op; tail_len[chunk] = last; }
            op += hdr + last;
            if (tid == 0) csz[chunk] = op;
        }
        __syncthreads();
    }
}

template <int MODE, int ACC>
inline cudaError_t c2_launch_mode(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                                  u32 n, u32* ticket, int sm_count, cudaStream_t st, const u64* n_dev, u32* tail_pos, u32* tail_len, const u32* hist, u64* prof)
{
    // the opt-in is per device (a process may hold contexts on several GPUs): set it at every launch, it is cheap
    {
        cudaError_t e = cudaFuncSetAttribute(compress_c2_kernel<MODE, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)C2Smem::BYTES);
        if (e != cudaSuccess) return e;
    }
    unsigned grid = (unsigned)(sm_count * 2);
    if (grid > n) grid = n;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    compress_c2_kernel<MODE, ACC><<<grid, C2_THREADS, C2Smem::BYTES, st>>>(in, in_off, usz, stage, stride, csz, n, ticket, n_dev, tail_pos, tail_len, hist, prof);
    return cudaGetLastError();
}

inline cudaError_t c2_launch_compress(const u8* in, const u64* in_off, const u64* usz, u8* stage, u64 stride, u64* csz,
                                      u32 n, u32* ticket, int sm_count, cudaStream_t st, const u64* n_dev = nullptr,
                                      u32* tail_pos = nullptr, u32* tail_len = nullptr, const u32* hist = nullptr, u64* prof = nullptr,
                                      u32 accel = 1, int mode = 2)
{
    // acceleration (lz4fast levels): probe grid of 2 (accel 2-3) or 4 (accel >= 4) positions.  A grid of 8 was measured:
    // 133 GB/s but 6.8 % larger than the reference at accel 8, while the grid of 4 is SMALLER than the reference from
    // accel 8 on (66.0 % vs 70.4 %) at 127 GB/s.  The match-finding modes 0 / 1 exist only at acceleration 1.
    if (accel >= 4) return c2_launch_mode<2, 4>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, n_dev, tail_pos, tail_len, hist, prof);
    if (accel >= 2) return c2_launch_mode<2, 2>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, n_dev, tail_pos, tail_len, hist, prof);
    switch (mode) {
    case 0:  return c2_launch_mode<0, 1>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, n_dev, tail_pos, tail_len, hist, prof);
    case 2:  return c2_launch_mode<2, 1>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, n_dev, tail_pos, tail_len, hist, prof);
    default: return c2_launch_mode<1, 1>(in, in_off, usz, stage, stride, csz, n, ticket, sm_count, st, n_dev, tail_pos, tail_len, hist, prof);
    }
}

}  // namespace lz4b200
