// kernels_v1.cuh -- generation-1 kernels: one CTA per chunk, the whole chunk resident in shared
// memory (compressed image staged by a TMA bulk copy, 64 KB output tile written back with a
// TMA bulk store), persistent grid with a work ticket.
//
// DECODE.  The reference decoder is one serial chain per block: token -> lengths -> next token
// (lz4/lz4.c:1929-2151).  Here the chain is broken up front:
//   P1  speculative parse.  The compressed image is cut into W windows of WS bytes; one lane per
//       window walks the token chain from a guessed start (the window start) to the first token
//       at or past the window end.  LZ4 token chains started at a wrong byte re-synchronise with
//       the true chain within a few sequences, so after the walk most EXITS are already right.
//       Then entry[w] := exit[w-1] and only windows whose entry changed walk again, until
//       nothing changes.  Window 0 starts at byte 0, so by induction every entry is then the
//       true one (at most W rounds in an adversarial stream; 2-3 in practice).
//   P2  exclusive scan of the bytes each window produces -> every window knows where its output
//       starts; total must equal the chunk size.
//   P3  literal pass: each lane walks its window's sequences and copies the literal runs into
//       the output tile; runs >= LONG bytes are queued and copied by whole warps afterwards.
//   P4  match pass: each lane executes its window's matches in order.  A match may read bytes
//       that another window has not produced yet, so every window publishes its progress and a
//       lane only executes a match whose source range is complete (rounds, no blocking waits:
//       the earliest unfinished match of the chunk is always ready, so every round progresses).
//       Matches >= LONG bytes are executed by the whole warp.
//   P5  the tile goes back to HBM with one bulk store.
// Chunks with very few sequences (incompressible or highly repetitive data: one or two
// sequences of tens of KB) take a cooperative path instead: the CTA executes the sequences one
// after another, every copy spread over all threads.
// All format checks of LZ4_decompress_safe are kept (status codes in common.cuh).
#pragma once
#include "common.cuh"

namespace lz4b200 {

// ---- PTX wrappers: mbarrier + bulk async copies (TMA engine, SASS UBLKCP) ---------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity)
{
    u32 ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, u32 bytes, u64* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, u32 bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- decode ------------------------------------------------------------------------------

constexpr u32 D1_END_OK = 0xFFFFFFF0u;           // chain ended with a well-formed last sequence
constexpr u32 D1_END_BAD = 0xFFFFFFF1u;          // chain ran into a malformed sequence
constexpr u32 D1_LONG = 48;                      // copies this long are done by a whole warp
constexpr u32 D1_FEW_SEQ = 48;                   // chunks with at most this many sequences: cooperative path

template <u32 CHUNK_MAX, u32 THREADS, u32 WS_MIN>
struct D1Cfg {
    static constexpr u32 IN_CAP = ((CHUNK_MAX + CHUNK_MAX / 255 + 16 + 15) & ~15u) + 32;   // compress bound + alignment slack
    static constexpr u32 OUT_CAP = (CHUNK_MAX + 15) & ~15u;
    static constexpr u32 MAXW = THREADS;                              // parse windows <= threads
    static constexpr u32 LONGQ = IN_CAP / D1_LONG + 2;                // every long literal run eats >= D1_LONG input bytes
    static constexpr u32 GRAN_SHIFT = 5;
    static constexpr u32 NGRAN = (CHUNK_MAX >> GRAN_SHIFT) + 2;
    // byte layout of dynamic shared memory
    static constexpr u32 OFF_IN = 0;
    static constexpr u32 OFF_OUT = OFF_IN + IN_CAP;
    static constexpr u32 OFF_ENTRY = OFF_OUT + OUT_CAP;               // u32[MAXW+1]
    static constexpr u32 OFF_EXIT = OFF_ENTRY + 4 * (MAXW + 1);       // u32[MAXW]
    static constexpr u32 OFF_OBASE = OFF_EXIT + 4 * MAXW;             // u32[MAXW+1]  output start per window
    static constexpr u32 OFF_DONE = OFF_OBASE + 4 * (MAXW + 1);       // u32[MAXW]    output progress per window
    static constexpr u32 OFF_WINOF = OFF_DONE + 4 * MAXW;             // u16[NGRAN]
    static constexpr u32 OFF_LONGQ = (OFF_WINOF + 2 * NGRAN + 15) & ~15u;   // uint4[LONGQ] {src, dst, len, -}
    static constexpr u32 OFF_MISC = OFF_LONGQ + 16 * LONGQ;            // mbarrier + scalars
    static constexpr u32 SMEM_BYTES = OFF_MISC + 64;
};

struct D1Seq {           // one parsed sequence
    u32 lit_pos;         // first literal byte (position in the compressed image)
    u32 lit_len;
    u32 off;             // 0 for the last sequence
    u32 mlen;            // 0 for the last sequence
    u32 next;            // position of the next token, D1_END_OK after a good last sequence, D1_END_BAD on error
    u32 err;
};

// Parse the sequence whose token is at `pos`.  Only input-side rules are checked here
// (lz4/lz4.c:1976-2036, :2055-2067); rules that need the output position are checked by the caller.
__device__ __forceinline__ D1Seq d1_parse(const u8* in, u32 clen, u32 pos)
{
    D1Seq s; s.off = 0; s.mlen = 0; s.err = ST_OK; s.lit_pos = 0; s.lit_len = 0; s.next = D1_END_BAD;
    if (pos >= clen) { s.err = ST_TOKEN_OOB; return s; }
    const u32 token = in[pos++];
    u32 len = token >> 4;
    if (len == RUN_MASK) {
        if (pos + RUN_MASK >= clen) { s.err = ST_LITLEN_OOB; return s; }
        u32 b;
        do {
            b = in[pos++]; len += b;
            if (b == 255 && pos + RUN_MASK >= clen) { s.err = ST_LITLEN_OOB; return s; }
        } while (b == 255);
    }
    s.lit_pos = pos; s.lit_len = len;
    const u32 after = pos + len;                       // len <= 255 * clen: no 32-bit wrap for clen < 2^23
    if (after + (2 + 1 + LASTLITERALS) > clen) {       // input says: last sequence
        if (after != clen) { s.err = ST_LAST_LITERALS; return s; }
        s.next = D1_END_OK;
        return s;
    }
    s.off = (u32)in[after] | ((u32)in[after + 1] << 8);
    pos = after + 2;
    u32 ml = token & ML_MASK;
    if (ml == ML_MASK) {
        u32 b;
        do {
            b = in[pos++]; ml += b;
            if (pos + LASTLITERALS - 1 >= clen) { s.err = ST_MATCHLEN_OOB; return s; }
        } while (b == 255);
    }
    s.mlen = ml + MINMATCH;
    if (s.off == 0) { s.err = ST_OFFSET; return s; }
    s.next = pos;
    return s;
}

// Walk one window: from `pos` to the first token at or past `wend`.  Returns the exit and the
// number of output bytes / sequences produced by the sequences that start inside the window.
__device__ __forceinline__ u32 d1_walk(const u8* in, u32 clen, u32 pos, u32 wend, u32& out_bytes, u32& n_seq)
{
    out_bytes = 0; n_seq = 0;
    while (pos < wend) {
        const D1Seq s = d1_parse(in, clen, pos);
        if (s.err) return D1_END_BAD;
        out_bytes += s.lit_len + s.mlen; n_seq++;
        pos = s.next;                                  // >= wend terminates; END markers are > any wend
    }
    return pos;
}

// smem -> smem copy by one warp; src/dst/len arbitrary, non-overlapping.
__device__ __forceinline__ void d1_warp_copy(u8* dst, const u8* src, u32 len, u32 lane)
{
    for (u32 i = lane; i < len; i += 32) dst[i] = src[i];
}

// Overlap-safe match copy by `nthr` cooperating threads (tid in [0,nthr)); all threads of the
// group must call it, followed by the group's own barrier.  off < len: periodic extension.
__device__ __forceinline__ void d1_group_match(u8* out, u32 dst, u32 off, u32 len, u32 tid, u32 nthr)
{
    const u8* src = out + dst - off;
    if (off >= len) {
        for (u32 i = tid; i < len; i += nthr) out[dst + i] = src[i];
    } else {
        u32 r = tid % off;
        const u32 step = nthr % off;
        for (u32 i = tid; i < len; i += nthr) {
            out[dst + i] = src[r];
            r += step; if (r >= off) r -= off;
        }
    }
}

template <u32 CHUNK_MAX, u32 THREADS, u32 WS_MIN>
__global__ void __launch_bounds__(THREADS, 1)
decompress_v1_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ csz,
                     const u64* __restrict__ out_off, const u64* __restrict__ usz, u8* __restrict__ out,
                     u32* __restrict__ status, u32* __restrict__ bad_chunks, u32 n_chunks, u32* ticket, u64* prof)
{
    using Cfg = D1Cfg<CHUNK_MAX, THREADS, WS_MIN>;
    extern __shared__ __align__(128) u8 smem[];
    u8* const s_in_raw = smem + Cfg::OFF_IN;
    u8* const s_out = smem + Cfg::OFF_OUT;
    u32* const s_entry = reinterpret_cast<u32*>(smem + Cfg::OFF_ENTRY);
    u32* const s_exit = reinterpret_cast<u32*>(smem + Cfg::OFF_EXIT);
    u32* const s_obase = reinterpret_cast<u32*>(smem + Cfg::OFF_OBASE);
    volatile u32* const s_done = reinterpret_cast<volatile u32*>(smem + Cfg::OFF_DONE);
    u16* const s_winof = reinterpret_cast<u16*>(smem + Cfg::OFF_WINOF);
    uint4* const s_longq = reinterpret_cast<uint4*>(smem + Cfg::OFF_LONGQ);
    u64* const s_bar = reinterpret_cast<u64*>(smem + Cfg::OFF_MISC);
    u32* const s_misc = reinterpret_cast<u32*>(smem + Cfg::OFF_MISC + 16);    // [0] chunk, [1] long count, [2] err, [3] scratch

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr u32 NWARPS = THREADS / 32;
    u32 phase = 0;
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();
    // optional phase profile (tid 0 only): cycles spent in load / parse / scan / literals / matches / store
    long long t_prev = prof ? clock64() : 0;
    u64 t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define D1_TICK(k) do { if (prof && tid == 0) { long long t_ = clock64(); t_acc[k] += (u64)(t_ - t_prev); t_prev = t_; } } while (0)

    for (;;) {
        if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 chunk = s_misc[0];
        if (chunk >= n_chunks) break;
        D1_TICK(0);
        const u8* src = in + in_off[chunk];
        const u32 clen = (u32)csz[chunk];
        const u32 n = (u32)usz[chunk];
        u8* dst = out + out_off[chunk];
        u32 err = ST_OK;

        if (clen == n) {                                   // raw chunk (_lzbench/lzbench.cpp:311-315)
            coop_copy(dst, src, n, tid, THREADS);
            if (tid == 0) status[chunk] = ST_OK;
            __syncthreads();
            continue;
        }
        if (n == 0 || clen == 0 || clen > Cfg::IN_CAP - 32 || n > CHUNK_MAX) {
            if (tid == 0) {
                const bool ok = n == 0 && clen == 1 && src[0] == 0;
                status[chunk] = ok ? ST_OK : ST_EMPTY;
                if (!ok && bad_chunks) atomicAdd(bad_chunks, 1u);
            }
            __syncthreads();
            continue;
        }

        // ---- stage the compressed image: bulk copy of the 16-byte aligned interior, edges by hand ----
        const u32 head = (u32)(reinterpret_cast<uintptr_t>(src) & 15);
        u8* const s_in = s_in_raw + head;                  // byte j of the chunk lives at s_in[j]
        {
            const u32 lead = head ? 16 - head : 0;          // bytes before the first aligned address
            const u32 lead_n = lead < clen ? lead : clen;
            const u32 bulk = (clen - lead_n) & ~15u;
            if (tid == 0) {
                s_misc[1] = 0; s_misc[2] = ST_OK;
                if (bulk) {
                    fence_async_smem();
                    mbar_expect_tx(s_bar, bulk);
                    bulk_g2s(s_in + lead_n, src + lead_n, bulk, s_bar);
                }
            }
            if (tid < lead_n) s_in[tid] = src[tid];
            const u32 tail0 = lead_n + bulk;
            if (tid >= 32 && tid - 32 < clen - tail0) s_in[tail0 + tid - 32] = src[tail0 + tid - 32];
            if (bulk) { mbar_wait(s_bar, phase); phase ^= 1; }
        }
        __syncthreads();
        D1_TICK(1);

        // ---- P1: speculative parse ----
        // window size: power of two, at least WS_MIN, at most MAXW windows
        u32 ws_shift = 31 - __clz(WS_MIN);
        while (((clen + (1u << ws_shift) - 1) >> ws_shift) > Cfg::MAXW) ws_shift++;
        const u32 W = (clen + (1u << ws_shift) - 1) >> ws_shift;     // 1 <= W <= MAXW
        const u32 wend = (tid + 1) << ws_shift;
        u32 my_out = 0, my_nseq = 0;
        if (tid < W) {
            s_entry[tid] = tid << ws_shift;
            s_exit[tid] = d1_walk(s_in, clen, tid << ws_shift, wend, my_out, my_nseq);
        }
        __syncthreads();
        // Fix-up rounds.  A window that walked ("producer") claims the window its exit lands in and
        // covers every window it jumps over (long literal runs).  Window w takes its entry from the
        // producer before it with the furthest cover: exclusive prefix-max of (cover << 16 | index).
        // Once every window before w holds its true entry that producer is the last true one, so w is
        // settled one round later (induction from window 0); typical streams settle in 2-3 rounds.
        for (;;) {
            u32 key = 0;
            if (tid < W && s_entry[tid] < wend) {
                const u32 x = s_exit[tid];
                const u32 cover = x >= D1_END_OK ? W : (x >> ws_shift);
                key = (cover << 16) | tid;
            }
            u32 incl = key;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 y = __shfl_up_sync(FULL, incl, d);
                if (lane >= (u32)d && y > incl) incl = y;
            }
            u32* s_wmax = reinterpret_cast<u32*>(s_longq);            // scratch: the long queue is idle in P1/P2
            if (lane == 31) s_wmax[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                u32 m = lane < NWARPS ? s_wmax[lane] : 0;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 y = __shfl_up_sync(FULL, m, d);
                    if (lane >= (u32)d && y > m) m = y;
                }
                s_wmax[32 + lane] = m;
            }
            __syncthreads();
            u32 best = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) best = 0;
            if (warp) { const u32 y = s_wmax[32 + warp - 1]; if (y > best) best = y; }
            bool dirty = false;
            u32 e_new = 0;
            if (tid > 0 && tid < W) {
                const u32 cw = best >> 16, cv = best & 0xFFFFu;
                e_new = (best && cw >= tid) ? s_exit[cv] : s_exit[tid - 1];
                dirty = e_new != s_entry[tid];
            }
            if (!__syncthreads_or(dirty)) break;           // also orders the s_exit reads before the rewrites below
            if (dirty) {
                s_entry[tid] = e_new;
                s_exit[tid] = d1_walk(s_in, clen, e_new, wend, my_out, my_nseq);
            }
            __syncthreads();
        }
        D1_TICK(2);
        // chain state now: window w starts at s_entry[w]; END markers pass through windows untouched
        // ---- P2: scan of output bytes (and sequence counts) over windows ----
        {
            u32 x = tid < W ? my_out : 0, q = tid < W ? my_nseq : 0;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 y = __shfl_up_sync(FULL, x, d), z = __shfl_up_sync(FULL, q, d);
                if (lane >= (u32)d) { x += y; q += z; }
            }
            u32* s_wsum = reinterpret_cast<u32*>(s_longq);         // scratch: long queue is not in use yet
            if (lane == 31) { s_wsum[warp] = x; s_wsum[32 + warp] = q; }
            __syncthreads();
            if (warp == 0) {
                u32 a = lane < NWARPS ? s_wsum[lane] : 0, b = lane < NWARPS ? s_wsum[32 + lane] : 0;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    u32 y = __shfl_up_sync(FULL, a, d), z = __shfl_up_sync(FULL, b, d);
                    if (lane >= (u32)d) { a += y; b += z; }
                }
                s_wsum[lane] = a; s_wsum[32 + lane] = b;
            }
            __syncthreads();
            const u32 base = x - (tid < W ? my_out : 0) + (warp ? s_wsum[warp - 1] : 0);
            const u32 total = s_wsum[NWARPS - 1];
            const u32 total_seq = s_wsum[32 + NWARPS - 1];
            if (tid < W) { s_obase[tid] = base; s_done[tid] = base; }
            if (tid == 0) s_obase[W] = total;
            const u32 last_exit = s_exit[W - 1];
            __syncthreads();                                        // s_wsum (aliasing the long queue) is dead from here
            if (last_exit != D1_END_OK) err = last_exit == D1_END_BAD ? ST_LAST_LITERALS : ST_TOKEN_OOB;
            else if (total != n) err = ST_SIZE;

            D1_TICK(3);
            if (!err && total_seq <= D1_FEW_SEQ) {
                // ---- cooperative path: few, long sequences; the CTA executes them one by one ----
                u32 pos = 0, op = 0;
                for (;;) {
                    const D1Seq s = d1_parse(s_in, clen, pos);      // uniform: every thread parses the same token
                    if (s.err) { err = s.err; break; }
                    const u32 cpy = op + s.lit_len;
                    if (s.next == D1_END_OK) {
                        if (cpy != n) { err = cpy > n ? ST_LAST_LITERALS : ST_SIZE; break; }
                    } else if (cpy + MFLIMIT > n) { err = ST_LAST_LITERALS; break; }
                    for (u32 i = tid; i < s.lit_len; i += THREADS) s_out[op + i] = s_in[s.lit_pos + i];
                    op = cpy;
                    if (s.next == D1_END_OK) break;
                    if (s.off > op) { err = ST_OFFSET; break; }
                    if (op + s.mlen + LASTLITERALS > n) { err = ST_MATCH_END; break; }
                    __syncthreads();
                    d1_group_match(s_out, op, s.off, s.mlen, tid, THREADS);
                    __syncthreads();
                    op += s.mlen;
                    pos = s.next;
                }
                __syncthreads();
            } else if (!err) {
                // ---- P3: literal pass (one lane per window) ----
                const u32 my_base = tid < W ? s_obase[tid] : 0;
                if (tid < W) {
                    // granule -> window map: window tid owns every 32-byte granule whose first byte it produces
                    const u32 lo = my_base, hi = s_obase[tid + 1];
                    if (hi > lo || tid == 0) {
                        u32 g = (lo + 31) >> Cfg::GRAN_SHIFT;
                        if (tid == 0) g = 0;
                        for (; (g << Cfg::GRAN_SHIFT) < hi; g++) s_winof[g] = (u16)tid;
                    }
                    u32 pos = s_entry[tid], op = my_base;
                    while (pos < wend) {
                        const D1Seq s = d1_parse(s_in, clen, pos);
                        if (s.err) { s_misc[2] = s.err; break; }
                        const u32 cpy = op + s.lit_len;
                        if (s.next != D1_END_OK && cpy + MFLIMIT > n) { s_misc[2] = ST_LAST_LITERALS; break; }
                        if (s.lit_len >= D1_LONG) {
                            const u32 slot = atomicAdd(&s_misc[1], 1u);
                            s_longq[slot] = make_uint4(s.lit_pos, op, s.lit_len, 0);
                        } else {
                            for (u32 i = 0; i < s.lit_len; i++) s_out[op + i] = s_in[s.lit_pos + i];
                        }
                        op = cpy + s.mlen;
                        pos = s.next;
                    }
                }
                __syncthreads();
                {   // long literal runs: one warp per run
                    const u32 nlong = s_misc[1];
                    for (u32 e = warp; e < nlong; e += NWARPS) {
                        const uint4 q = s_longq[e];
                        d1_warp_copy(s_out + q.y, s_in + q.x, q.z, lane);
                    }
                }
                __syncthreads();
                err = s_misc[2];
                D1_TICK(4);

                // ---- P4: match pass ----
                if (!err) {
                    u32 pos = tid < W ? s_entry[tid] : D1_END_OK;
                    u32 op = my_base;
                    bool active = tid < W && pos < wend;
                    bool have = false;
                    D1Seq s;
                    u32 mdst = 0;
                    u32 rounds = 0;
                    while (__any_sync(FULL, active)) {
                        if (++rounds > (1u << 22)) { s_misc[2] = ST_SIZE; break; }   // watchdog: cannot trigger on a consistent chain
                        bool do_long = false;
                        if (active) {
                            if (!have) {
                                s = d1_parse(s_in, clen, pos);
                                mdst = op + s.lit_len;
                                have = true;
                                if (s.next == D1_END_OK) {                    // last sequence: literals only
                                    op = mdst; have = false; active = false;
                                    __threadfence_block();
                                    s_done[tid] = op;
                                } else if (s.off > mdst) { s_misc[2] = ST_OFFSET; s.off = mdst ? mdst : 1; s.mlen = 0; }
                                else if (mdst + s.mlen + LASTLITERALS > n) { s_misc[2] = ST_MATCH_END; s.mlen = 0; }
                            }
                            if (have) {
                                // source range [a, b): for an overlapping match only the first `off` bytes are foreign
                                const u32 a = mdst - s.off;
                                const u32 b = s.off < s.mlen ? mdst : a + s.mlen;
                                bool ready = true;
                                if (a < my_base && s.mlen) {
                                    u32 w = s_winof[a >> Cfg::GRAN_SHIFT];
                                    for (;;) {
                                        if (w >= tid) break;                  // the rest lies in my own window
                                        const u32 wnext = s_obase[w + 1];
                                        if (wnext <= a) { w++; continue; }    // that window ends before my source starts
                                        const u32 need = b < wnext ? b : wnext;
                                        if (s_done[w] < need) { ready = false; break; }
                                        if (wnext >= b) break;
                                        w++;
                                    }
                                }
                                if (ready) {
                                    __threadfence_block();                     // order the progress reads before the data reads
                                    if (s.mlen >= D1_LONG) do_long = true;
                                    else {
                                        const u8* m = s_out + a;
                                        for (u32 i = 0; i < s.mlen; i++) s_out[mdst + i] = m[i];
                                    }
                                }
                                if (ready && !do_long) {
                                    op = mdst + s.mlen; pos = s.next; have = false;
                                    __threadfence_block();
                                    s_done[tid] = op;
                                    if (pos >= wend) active = false;
                                }
                            }
                        }
                        // long matches: the warp executes them one after another, all lanes copying
                        unsigned lm = __ballot_sync(FULL, do_long);
                        while (lm) {
                            const int l = __ffs(lm) - 1; lm &= lm - 1;
                            const u32 ldst = __shfl_sync(FULL, mdst, l), loff = __shfl_sync(FULL, s.off, l),
                                      llen = __shfl_sync(FULL, s.mlen, l);
                            d1_group_match(s_out, ldst, loff, llen, lane, 32);
                            __syncwarp();
                            if ((int)lane == l) {
                                op = mdst + s.mlen; pos = s.next; have = false;
                                __threadfence_block();
                                s_done[tid] = op;
                                if (pos >= wend) active = false;
                            }
                        }
                    }
                }
                __syncthreads();
                if (!err) err = s_misc[2];
            }
        }

        D1_TICK(5);
        // ---- P5: tile -> HBM ----
        if (!err) {
            fence_async_smem();                            // generic-proxy writes of the tile -> visible to the bulk engine
            __syncthreads();
            const u32 bulk = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (n & ~15u) : 0;
            if (bulk) {
                if (tid == 0) { bulk_s2g(dst, s_out, bulk); bulk_commit(); }
                if (tid >= 32 && tid - 32 < n - bulk) dst[bulk + tid - 32] = s_out[bulk + tid - 32];
                if (tid == 0) bulk_wait_read0();           // the tile may be overwritten once it has been read
            } else {
                for (u32 i = tid; i < n; i += THREADS) dst[i] = s_out[i];
            }
        }
        if (tid == 0) {
            status[chunk] = err;
            if (err && bad_chunks) atomicAdd(bad_chunks, 1u);
        }
        __syncthreads();
        D1_TICK(6);
    }
    if (prof && tid == 0) {
        for (int k = 0; k < 8; k++) atomicAdd(reinterpret_cast<unsigned long long*>(prof) + k, (unsigned long long)t_acc[k]);
    }
#undef D1_TICK
}

// ---- host-side launchers ---------------------------------------------------------------------

template <u32 CHUNK_MAX, u32 THREADS, u32 WS_MIN>
inline cudaError_t d1_launch(const u8* in, const u64* in_off, const u64* csz, const u64* out_off, const u64* usz,
                             u8* out, u32* status, u32* bad, u32 n, u32* ticket, u64* prof, int sm_count, cudaStream_t st)
{
    using Cfg = D1Cfg<CHUNK_MAX, THREADS, WS_MIN>;
    auto kern = decompress_v1_kernel<CHUNK_MAX, THREADS, WS_MIN>;
    // the opt-in is per device (a process may hold contexts on several GPUs): set it at every launch, it is cheap
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    int per_sm = (int)(227u * 1024u / (Cfg::SMEM_BYTES + 1024u));
    if (per_sm < 1) per_sm = 1;
    const int by_threads = 2048 / (int)THREADS;
    if (per_sm > by_threads) per_sm = by_threads;
    unsigned grid = (unsigned)(sm_count * per_sm);
    if (grid > n) grid = n;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    kern<<<grid, THREADS, Cfg::SMEM_BYTES, st>>>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, prof);
    return cudaGetLastError();
}

inline bool v1_decompress_supported(size_t max_chunk) { return max_chunk <= 65536; }

inline cudaError_t v1_launch_decompress(const u8* in, const u64* in_off, const u64* csz, const u64* out_off,
                                        const u64* usz, u8* out, u32* status, u32* bad, u32 n, u32* ticket,
                                        u64* prof, int sm_count, size_t max_chunk, cudaStream_t st)
{
    if (max_chunk <= 8192)
        return d1_launch<8192, 128, 128>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, prof, sm_count, st);
    if (max_chunk <= 32768)
        return d1_launch<32768, 256, 128>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, prof, sm_count, st);
    return d1_launch<65536, 512, 128>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, prof, sm_count, st);
}

}  // namespace lz4b200
