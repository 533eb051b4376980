// kernels_v7.cuh -- generation-7 decode: a sequence table first, then one LANE per sequence.
//
// Every earlier generation walks the token chain of a block (lz4/lz4.c:1797-2151) with a whole team of lanes:
// v5 spends 155 warp instructions per iteration on 4 sequences of ~16 bytes and is issue-bound (profiles/r01_*).
// The chain itself is cheap -- token -> literal length -> next token -- it is the copies that need the lanes.  So:
//
//  1. parse_v7_kernel: one lane per CHUNK walks the chain, checks every rule of the reference decoder that concerns
//     lengths and positions (lz4/lz4.c:1797-2151; the same checks as the general path of v5) and writes one 8-byte
//     entry per sequence: output position, literal count, position of the literals in the stream, match length.
//     32 chunks per warp instruction; the kernel is latency-bound (a few warps per SM).  The offsets are not on the
//     chain: they are read and checked in pass 2, one lane per sequence.
//  2. decompress_v7_kernel: one warp per chunk, 32 consecutive sequences per step, one lane each.  All literal
//     runs of a group are independent; matches depend only on earlier output.  Output goes to a per-warp linear
//     window in shared memory (the group's span plus ~1 KiB of history) and is flushed to global memory in
//     aligned 16-byte stores; sources below the window are fetched from global memory with two aligned 8-byte
//     loads.  Dependencies INSIDE a group (31 % of the matches of the bench corpus: record data where every
//     record copies from the previous one) are first shortened by source forwarding -- if the whole source of
//     match i lies inside the output of an earlier match j of the group, i can read from j's source instead
//     (shift_i += shift_j, pointer jumping over lanes with shuffles, 2.1 steps per group on the bench corpus,
//     scripts/sim/decode_sim.c) -- and what is left (2 % of the matches) is executed in dependency rounds.
//  Sequences with more than V7_CAP literals or match bytes end the group and are executed by the whole warp
//  (coalesced copies through global memory).  Chunks that are mostly such copies (incompressible or run-length
//  data: < 1 sequence per 256 bytes), chunks above 64 KiB and anything the table cannot represent are put on a
//  list for the team decoder (v5), which is at memcpy speed on them.
#pragma once
#include "common.cuh"
#include "kernels_v2.cuh"

namespace lz4b200 {

constexpr u32 V7_CAP = 64;                 // literal runs / matches up to this many bytes are copied by their own lane
constexpr u32 V7_SPAN = 2048;              // output bytes of one group (32 lanes x 2 x V7_CAP = 4096 would be the worst case: groups are cut)
constexpr u32 V7_WIN = 4096;               // bytes of the per-warp output window
constexpr u32 V7_KEEP = 1008;              // history kept below a group when the window slides
constexpr u32 V7_LPIECE = 8191;            // literal runs are cut into entries of at most this many bytes (13-bit field)
constexpr u32 V7_MAX_CHUNK = 65536;
constexpr u32 V7_BIG_MAX = 16u << 20;      // BIG variant (chunks above 64 KiB): the entries keep their 16-bit position fields, which then hold the low 16 bits --
                                           // positions only grow along a chunk and by less than 65 536 per entry (matches are cut into pieces of V7_MPIECE),
                                           // so pass 2 counts the wrap-arounds
constexpr u32 V7_MPIECE = 32768;
constexpr int V7_WARPS = 4;                // warps per CTA of the lane-per-sequence kernel
constexpr u32 V7_STAGE = 512;              // stream bytes staged per group (one 16-byte block per lane; a group reads ~330)
constexpr u32 V7_WARP_SMEM = 4096 + 32 + 2 * 512;    // V7_WIN + 32 + 2 * V7_STAGE
constexpr u32 V7_ROUTE_NONE = 0, V7_ROUTE_SEQ = 1, V7_ROUTE_TEAM = 2;
constexpr u32 V7_ZERO_WORDS = 2 * 64 + 4;  // one block per slice, cleared by ONE memset per pass: bucket counters, team_count, the two work tickets
constexpr u32 V7_BUCKETS = 64;             // chunks are ordered by sequence count / 128 (64 KiB chunks have up to ~8000)

// entries of the table a chunk may need: sequences (>= 4 output bytes each but the last), literal pieces, terminator
__host__ __device__ inline u64 v7_entries_for(u64 out_bytes, u64 chunks)
{
    // per chunk of n bytes: <= n/4 sequences with a match, the last one, <= n/8191 literal pieces, the terminator
    return out_bytes / 4 + out_bytes / 4096 + 4 * chunks;
}

struct V7Params {
    const u8* in; const u64* in_off; const u64* csz; const u64* out_off; const u64* usz; u8* out;
    u32* status; u32* bad_chunks; u32* ticket; u32 n_chunks;
    u32 cap_all;                           // see V5Params::cap_all
    u64* table; u64 table_cap;             // sequence table of this slice and its capacity in entries
    u32* nseq;                             // entries per chunk (without the terminator)
    u32* route;                            // V7_ROUTE_* per chunk
    u32* team_list; u32* team_count;       // chunks left to the team decoder
    u32* order; u32* bucket;               // pass 2 takes the chunks with the most sequences first: order[ticket], bucket counters u32[2][V7_BUCKETS]
    const u32* skip;                       // optional: chunks with skip[c] != 0 are somebody else's (kernels_coarse.cuh has decoded them)
};

// entry, low word: bits 0-16 output position of the sequence, 17-29 literal count; high word: bits 0-15 stream position
// of its literals (chunks whose stream is 64 KiB or longer go to the team decoder), 16-31 match length (0: no match; the
// two offset bytes follow the literals)
__device__ __forceinline__ u64 v7_pack(u32 op, u32 lsrc, u32 M, u32 L)
{
    return (u64)(op | (L << 17)) | ((u64)(lsrc | (M << 16)) << 32);
}

// chunks with more sequences take longer in pass 2; bucket 0 = nothing to do there
__device__ __forceinline__ u32 v7_bucket(u32 nseq) { const u32 b = 1u + (nseq >> 7); return b < V7_BUCKETS ? b : V7_BUCKETS - 1; }

// order[] = the chunks sorted by bucket, largest first (a warp of pass 2 works through ~5 chunks of ~0.5 ms each: with the
// long ones first the last warps finish closer together)
__global__ void __launch_bounds__(256)
order_v7_kernel(const __grid_constant__ V7Params P)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_chunks) return;
    const u32 b = P.route[c] == V7_ROUTE_SEQ ? v7_bucket(P.nseq[c]) : 0u;
    u32 base = 0;
    for (u32 i = V7_BUCKETS - 1; i > b; i--) base += P.bucket[i];
    P.order[base + atomicAdd(P.bucket + V7_BUCKETS + b, 1u)] = c;
}

// ---- pass 1: the token chain, one lane per chunk ---------------------------------------------------------------
// A warp step waits for its slowest lane, so nothing in the loop may be slow for ANY lane:
//  * 32 independent streams miss L1 somewhere in almost every step (first build: 2300 cycles per step,
//    profiles/r02_parse_*): every lane reads its stream through its own ring in shared memory, filled by cp.async
//    16 bytes at a time ahead of the read position (LDGSTS, as v5 does per team);
//  * nothing waits for memory: a lane whose bytes have not landed sits the step out, and a literal run that leaves
//    the ring restarts the ring behind it and waits for it the same way (V7M_WAITQ);
//  * the common sequence (length nibbles complete or one extension byte below 255, not near the end of the block,
//    every length rule satisfied) is decoded without a branch: token -> literal count -> match length byte -> next
//    token; everything else goes through the general code, which repeats the checks of the reference decoder in
//    its order and names the violated rule.
constexpr u32 V7_RING = 512;               // bytes of stream per lane in shared memory (+ 16: the first block again, so that short reads never wrap)
constexpr u32 V7_LOOK = 448;               // blocks are requested up to this far ahead of the read position
constexpr u32 V7_G = 1;                    // steps between a request and the first read that may rely on it
constexpr u32 V7_UNROLL = 16;               // common-case sequences per step
enum : u32 { V7M_TOKEN = 0, V7M_WAITQ = 1, V7M_IDLE = 2 };

__device__ __forceinline__ void v7_cp16(u8* smem_dst, const u8* gsrc)
{
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}

template <u32 LANES, bool BIG>            // LANES chunks per warp: 32 leaves one warp per scheduler on a 16 384-chunk batch; fewer lanes = more warps to interleave
__global__ void __launch_bounds__(32)
parse_v7_kernel(const __grid_constant__ V7Params P)
{
    __shared__ __align__(16) u8 ring_all[LANES][V7_RING + 16];
    const u32 c = blockIdx.x * LANES + threadIdx.x;
    const bool have = threadIdx.x < LANES && c < P.n_chunks;
    u32 err = ST_OK, route = V7_ROUTE_NONE, k = 0;
    u32 mode = V7M_IDLE;
    const u8* srcA = nullptr;              // the stream, from its start rounded down to 16 bytes: positions below count from here
    u64* tab = nullptr;
    u32 clen = 0, n = 0, capv = 0, mis = 0, ip = 0, op = 0;
    if (have) {
        const u8* src = P.in + P.in_off[c];
        const u64 clen64 = P.csz[c], n64 = P.usz[c];
        const u64 base = v7_entries_for(P.out_off[c] - P.out_off[0], c);
        if (BIG && P.skip && P.skip[c]) {
            route = V7_ROUTE_NONE;                                        // decoded elsewhere; its status is set there
        } else if (clen64 == n64) {
            route = n64 >= 4096 ? V7_ROUTE_TEAM : V7_ROUTE_SEQ;        // raw chunk (_lzbench/lzbench.cpp:311-315): a plain copy
        } else if (n64 > (BIG ? V7_BIG_MAX : V7_MAX_CHUNK) || clen64 > (BIG ? (u64)V7_BIG_MAX + (V7_BIG_MAX >> 7) : 0xFFFFull) || base + v7_entries_for(n64, 1) > P.table_cap ||
                   clen64 * 85 < n64 ||                                   // fewer than n / 256 sequences (>= 3 stream bytes each): the team decoder's case, no need to walk
                   clen64 + (n64 >> 5) >= n64) {                           // hardly compressed: long literal runs, the team decoder's case as well
            route = V7_ROUTE_TEAM;
        } else if (n64 == 0 || clen64 == 0) {                            // lz4/lz4.c:1776-1782
            err = (n64 == 0 && clen64 == 1 && src[0] == 0) ? ST_OK : ST_EMPTY;
        } else {
            mis = (u32)(reinterpret_cast<uintptr_t>(src) & 15u);
            srcA = src - mis; clen = (u32)clen64 + mis; n = (u32)n64; ip = mis;
            capv = P.cap_all > n ? P.cap_all : n;
            tab = P.table + base;
            route = V7_ROUTE_SEQ; mode = V7M_TOKEN;
        }
    }
    u8* const ring = ring_all[threadIdx.x & (LANES - 1)];
    u32 req = 0;                           // blocks below this position have been requested (multiple of 16)
    u32 landed = 0, pend = 0;              // positions below `landed` have arrived
    u32 rlo = 0;                           // ... and are in the ring if not below its last restart (and not yet overwritten: see V7_RD)
    u32 wtok = 0, wL = 0, wlsrc = 0;       // V7M_WAITQ: token, literal count and literal position of the sequence that waits for what follows its literals
#define V7_RING_AT(POS) ring[(POS) & (V7_RING - 1)]
#define V7_RD(POS) ((POS) < landed && (POS) >= rlo && (POS) + V7_RING >= req ? (u32)V7_RING_AT(POS) : (u32)__ldg(srcA + (POS)))
#define V7_REQUEST() do { v7_cp16(&V7_RING_AT(req), srcA + req); if ((req & (V7_RING - 1)) == 0) v7_cp16(ring + V7_RING, srcA + req); req += 16; } while (0)
    if (mode == V7M_TOKEN) {
        while (req < V7_RING && req < clen) V7_REQUEST();
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    landed = pend = req;
    u32 tok = V7_RING_AT(ip), te1 = V7_RING_AT(ip + 1);                  // token at ip and the byte behind it
    bool tv = mode == V7M_TOKEN && ip + 2 <= landed;                    // ... if they had landed when they were read
    // limits of the common case: the offset position q must satisfy q + 3 <= landed and q + 8 <= clen (not the last sequence,
    // lz4/lz4.c:1990), the literal end cpy + MFLIMIT <= capv, the match end <= n and + LASTLITERALS <= capv
    const bool tiny = capv < 32 || clen < 32;                             // everything through the general code
    const u32 lim_cpy = tiny ? 0u : capv - MFLIMIT;
    const u32 lim_nop = tiny ? 0u : (n < capv - LASTLITERALS ? n : capv - LASTLITERALS);
    u32 qlim = tiny ? 0u : (clen < landed + 5 ? clen : landed + 5) - 8;
    for (u32 it = 1; __any_sync(FULL, mode != V7M_IDLE); it++) {
        // ---- ring upkeep every V7_G steps: up to six blocks (a step consumes ~10 bytes), readable V7_G .. 2 V7_G steps later ----
        if (it % V7_G == 0) {
            const u32 cons = mode == V7M_WAITQ ? wlsrc + wL : ip;
            if (mode != V7M_IDLE && req < clen && req < cons + V7_LOOK) {
                if (cons >= req + 16) {                                   // the chain jumped past everything requested: restart there,
                    if (landed == req) req = rlo = cons & ~15u;           // once nothing is in flight (a slot must not have two copies pending)
                } else {                                                  // (req + 16 <= (cons & ~15) + V7_RING follows from req < cons + V7_LOOK)
                    const u32 lim = clen < cons + V7_LOOK ? clen : cons + V7_LOOK;
                    for (u32 j = 0; j < 20 && req < lim; j++) V7_REQUEST();
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            landed = pend; pend = req;
            qlim = tiny ? 0u : (clen < landed + 5 ? clen : landed + 5) - 8;
        }
        // ---- common case, no branch, V7_UNROLL sequences per step (what a step costs beside them -- the vote, the ring, the test
        //      for the general code -- is paid once).  The token (and the byte behind it) of the CURRENT position were read with the
        //      previous sequence (tok, te1, valid iff tv): a sequence reads its match length byte and, at the position the chain
        //      moves to if it is a common one, the next token. ----
        bool fast = false;
        #pragma unroll
        for (u32 u = 0; u < V7_UNROLL; u++) {
            const bool lx = (tok >> 4) == RUN_MASK;
            const u32 Lf = lx ? RUN_MASK + te1 : tok >> 4;
            const u32 lsrc_f = ip + (lx ? 2u : 1u);
            const u32 q = lsrc_f + Lf;                                    // position of the offset
            const bool mx = (tok & ML_MASK) == ML_MASK;
            const u32 ipn = q + (mx ? 3u : 2u);
            const u32 m1 = V7_RING_AT(q + 2);
            const u8* rn = ring + (ipn & (V7_RING - 1));
            const u32 tokn = rn[0], e1n = rn[1];                          // (whatever is there if it has not landed: tv says)
            const u32 ml_f = (mx ? ML_MASK + m1 : tok & ML_MASK) + MINMATCH;
            const u32 cpy_f = op + Lf, nop_f = cpy_f + ml_f;
            fast = tv && q <= qlim && cpy_f <= lim_cpy && nop_f <= lim_nop && !(lx && te1 == 255) && !(mx && m1 == 255);
            if (fast) {                                                   // (tv is only ever set in V7M_TOKEN)
                tab[k++] = BIG ? v7_pack(op & 0xFFFFu, (lsrc_f - mis) & 0xFFFFu, ml_f, Lf) : v7_pack(op, lsrc_f - mis, ml_f, Lf);
                op = nop_f;
                ip = ipn; tok = tokn; te1 = e1n; tv = ipn + 2 <= landed;
            }
        }
        // ---- everything else: one sequence through the general code, when its bytes are there (else the lane sits this step out) ----
        // TOKEN: 32 bytes from the token on must have landed (token, length bytes; longer runs of length bytes read global memory)
        // unless the stream ends first.  What follows a long literal run is asked for through WAITQ.
        if (!fast && mode != V7M_IDLE) {
        const bool go = (mode == V7M_TOKEN && (ip + 32 <= landed || (ip + 32 > clen && landed >= clen))) ||
                        (mode == V7M_WAITQ && (wlsrc + wL + 32 <= landed || landed >= clen));
        if (go) do {
            u32 tk, len, lsrc;
            if (mode == V7M_TOKEN) {
                if (ip >= clen) { err = ST_TOKEN_OOB; mode = V7M_IDLE; break; }
                tk = V7_RD(ip); ip++;
                len = tk >> 4;
                if (len == RUN_MASK) {                                   // lz4/lz4.c:1707-1729 read_variable_length
                    if (ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; mode = V7M_IDLE; break; }
                    u32 s;
                    // a run of 255s, four at a time while they are in the ring and far from the end (the check below holds for all four)
                    while (ip + 4 <= landed && ip >= rlo && ip + 4 + RUN_MASK < clen &&
                           (V7_RING_AT(ip) & V7_RING_AT(ip + 1) & V7_RING_AT(ip + 2) & V7_RING_AT(ip + 3)) == 255) { len += 1020; ip += 4; }
                    do {
                        s = V7_RD(ip); ip++; len += s;
                        if (s == 255 && ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                    } while (s == 255);
                    if (err) { mode = V7M_IDLE; break; }
                }
                lsrc = ip;
                const u32 qq = ip + len;
                if (qq + 32 > landed && landed < clen && qq + (2 + 1 + LASTLITERALS) <= clen && op + len + MFLIMIT <= capv) {
                    wtok = tk; wL = len; wlsrc = lsrc; mode = V7M_WAITQ;  // what follows the literals is not there yet: ask for it, come back
                    break;
                }
            } else { tk = wtok; len = wL; lsrc = wlsrc; ip = lsrc; }
            mode = V7M_IDLE;                                              // set back when the sequence is complete and not the last
            const u32 cpy = op + len;                                    // len < 255 * clen: no wrap
            const bool last = cpy + MFLIMIT > capv || ip + len + (2 + 1 + LASTLITERALS) > clen;
            if (last && (ip + len != clen || cpy > n)) { err = ST_LAST_LITERALS; break; }
            if (cpy > n) { err = ST_MATCH_END; break; }                  // (frame capacity > n) rejected below anyway; keeps the entry count bounded by n
            u32 ls = lsrc - mis, Lr = len, lop = op;
            while (Lr > V7_LPIECE) {                                     // literal-only pieces
                tab[k++] = BIG ? v7_pack(lop & 0xFFFFu, ls & 0xFFFFu, 0, V7_LPIECE) : v7_pack(lop, ls, 0, V7_LPIECE);
                lop += V7_LPIECE; ls += V7_LPIECE; Lr -= V7_LPIECE;
            }
            if (last) {
                tab[k++] = BIG ? v7_pack(lop & 0xFFFFu, ls & 0xFFFFu, 0, Lr) : v7_pack(lop, ls, 0, Lr);
                op = cpy;
                if (op != n) err = ST_SIZE;
                break;
            }
            ip += len + 2; op = cpy;                                     // the offset itself is read and checked in pass 2
            u32 m = tk & ML_MASK;
            if (m == ML_MASK) {
                u32 s;
                do {
                    s = V7_RD(ip); ip++; m += s;
                    if (ip + LASTLITERALS - 1 >= clen) { err = ST_MATCHLEN_OOB; break; }
                } while (s == 255);
                if (err) break;
            }
            m += MINMATCH;
            if (op + m + LASTLITERALS > capv || op + m > n) { err = ST_MATCH_END; break; }
            if (BIG) {                                                    // pieces of a long match: the same offset again (the two offset bytes follow the literals)
                u32 mr = m;
                while (mr > V7_MPIECE) {
                    tab[k++] = v7_pack(lop & 0xFFFFu, ls & 0xFFFFu, V7_MPIECE, Lr);
                    lop += Lr + V7_MPIECE; ls += Lr; Lr = 0; mr -= V7_MPIECE;
                }
                tab[k++] = v7_pack(lop & 0xFFFFu, ls & 0xFFFFu, mr, Lr);
            } else
            tab[k++] = v7_pack(lop, ls, m, Lr);
            op += m;
            mode = V7M_TOKEN;
        } while (0);
        {                                                                 // the general code moved ip, or the token had not landed: read it (again)
            const u8* ra = ring + (ip & (V7_RING - 1));
            tok = ra[0]; te1 = ra[1]; tv = mode == V7M_TOKEN && ip + 2 <= landed && ip >= rlo;
        }
        }
    }
#undef V7_RD
#undef V7_RING_AT
#undef V7_REQUEST
    if (!have) return;
    if (BIG && P.skip && P.skip[c]) { P.nseq[c] = 0; P.route[c] = V7_ROUTE_NONE; atomicAdd(P.bucket, 1u); return; }
    if (!err && route == V7_ROUTE_SEQ && (u64)k * 256 < n) route = V7_ROUTE_TEAM;     // mostly long copies: the team decoder's case
    if (err) route = V7_ROUTE_NONE;
    P.status[c] = err;
    P.nseq[c] = k;
    P.route[c] = route;
    if (err && P.bad_chunks) atomicAdd(P.bad_chunks, 1u);
    if (route == V7_ROUTE_TEAM) P.team_list[atomicAdd(P.team_count, 1u)] = c;
    atomicAdd(P.bucket + (route == V7_ROUTE_SEQ ? v7_bucket(k) : 0u), 1u);
}

// ---- pass 2: one warp per chunk, one lane per sequence ----------------------------------------------------------

__device__ __forceinline__ u64 ldcg_u64(const void* p)
{
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldcg_u128(const void* p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// window -> global for the completed positions [from, to): byte stores up to the first 16-byte boundary, 16-byte
// stores for whole blocks, and (all = true) byte stores for the rest.  Positions are relative to the 16-aligned dstv.
__device__ __forceinline__ void v7_flush(u8* dstv, const u8* win, u32 win_base, u32& from, u32 to, bool all, u32 lane)
{
    if (from & 15u) {
        u32 hend = (from + 15u) & ~15u;
        if (hend > to) hend = to;
        if (from + lane < hend) dstv[from + lane] = win[from - win_base + lane];
        from = hend;
    }
    const u32 to16 = to & ~15u;
    for (u32 x = from + lane * 16u; x < to16; x += 512u)
        *reinterpret_cast<uint4*>(dstv + x) = *reinterpret_cast<const uint4*>(win + (x - win_base));
    if (to16 > from) from = to16;
    if (all && from < to) {
        if (from + lane < to) dstv[from + lane] = win[from - win_base + lane];
        from = to;
    }
}

template <bool BIG>
__global__ void __launch_bounds__(V7_WARPS * 32, 7)
decompress_v7_kernel(const __grid_constant__ V7Params P)
{
    // per warp: the output window, then two buffers for the stream bytes of the next group (on their way while this one runs)
    extern __shared__ __align__(16) u8 v7_smem[];
    const u32 lane = lane_id();
    u8* const win = v7_smem + (threadIdx.x >> 5) * V7_WARP_SMEM;
    u8* const stage0 = win + V7_WIN + 32;
    const unsigned lt_mask = (1u << lane) - 1u;

    for (;;) {
        u32 c = 0;
        if (lane == 0) c = atomicAdd(P.ticket, 1u);
        c = __shfl_sync(FULL, c, 0);
        if (c >= P.n_chunks) break;
        c = P.order[c];
        if (P.route[c] != V7_ROUTE_SEQ) break;                            // the rest of the order is not for this kernel
        const u8* __restrict__ src = P.in + P.in_off[c];
        u8* const dst = P.out + P.out_off[c];
        const u32 clen = (u32)P.csz[c], n = (u32)P.usz[c];
        if (clen == n) { coop_copy(dst, src, n, lane, 32); continue; }
        const u32 ns = P.nseq[c];
        const u64* __restrict__ tab = P.table + v7_entries_for(P.out_off[c] - P.out_off[0], c);
        // positions below are relative to dstv, the chunk start rounded down to 16 bytes
        const u32 A = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
        u8* const dstv = dst - A;
        u32 win_base = 0;                  // position of win[0]; multiple of 16
        u32 flushed = A;                   // positions below are in global memory
        u32 cur = 0;
        bool bad = false;                  // an offset of this chunk points before its start (or is 0)
        u64 e_next = lane < ns ? tab[lane] : 0ull;
        // stream positions below count from srcA, the stream start rounded down to 16 bytes (cp.async moves aligned blocks)
        const u32 mis = (u32)(reinterpret_cast<uintptr_t>(src) & 15u);
        const u8* const srcA = src - mis;
        const u32 clenA = clen + mis;
        u32 sb = 0, sbuf = 0;              // the staged bytes are positions [sb, sb + V7_STAGE) in buffer sbuf
        u32 obase = 0, lbase = 0, oprev = 0, lprev = 0;   // BIG: multiples of 65 536 the 16-bit fields have passed, and the fields of the last entry consumed
        {
            u8* st = stage0;
            if (lane * 16u < clenA) v7_cp16(st + lane * 16u, srcA + lane * 16u);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        while (cur < ns) {
            const bool valid = cur + lane < ns;
            const u64 e = e_next;
            u32 op = ((u32)e & 0x1FFFFu) + A;
            const u32 L = (u32)e >> 17;
            u32 lsrc = (u32)(e >> 32) & 0xFFFFu;
            const u32 M = (u32)(e >> 48);
            unsigned owrap = 0, lwrap = 0;
            u32 o16 = 0, l16 = 0;
            if (BIG) {                                                    // the fields hold the low 16 bits: count the wrap-arounds up to my entry
                o16 = (u32)e & 0xFFFFu; l16 = lsrc;
                u32 po = __shfl_up_sync(FULL, o16, 1), pl = __shfl_up_sync(FULL, l16, 1);
                if (lane == 0) { po = oprev; pl = lprev; }
                owrap = __ballot_sync(FULL, valid && o16 < po);
                lwrap = __ballot_sync(FULL, valid && l16 < pl);
                const unsigned upto = (2u << lane) - 1u;
                op = obase + ((u32)__popc(owrap & upto) << 16) + o16 + A;
                lsrc = lbase + ((u32)__popc(lwrap & upto) << 16) + l16;
            }
            const u32 opn = op + L + M;
            // the literals and the offset of this lane: from the staged bytes if they are all there, else from global memory
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const u32 lpos = lsrc + mis;
            const bool staged = lpos >= sb && lpos + L + 2 <= sb + V7_STAGE;
            const u8* const stg = stage0 + sbuf * V7_STAGE + (lpos - sb);
            const u8* const lit = staged ? stg : srcA + lpos;           // generic pointer: shared or global
            u32 off = 0;
            if (valid && M) off = (u32)lit[L] | ((u32)lit[L + 1] << 8);
            const u32 op0 = __shfl_sync(FULL, op, 0);
            const bool lng = valid && (L > V7_CAP || M > V7_CAP);
            const unsigned stop = __ballot_sync(FULL, !valid || lng || opn - op0 > V7_SPAN);
            const u32 k = stop ? (u32)__ffs(stop) - 1u : 32u;          // lanes [0, k) form the group
            const bool lng_k = k < 32 && __shfl_sync(FULL, (u32)lng, k & 31);       // ... and a long sequence follows it
            if (BIG) {                                                    // the entries this iteration consumes move the bases on
                const u32 used = k + (lng_k ? 1u : 0u);                  // (at least 1)
                const unsigned um = used >= 32 ? FULL : ((1u << used) - 1u);
                obase += (u32)__popc(owrap & um) << 16; lbase += (u32)__popc(lwrap & um) << 16;
                oprev = __shfl_sync(FULL, o16, used - 1); lprev = __shfl_sync(FULL, l16, used - 1);
            }
            {   // the next group's entries and stream bytes: on their way while this one is executed
                const u32 nx = cur + k + (lng_k ? 1u : 0u) + lane;
                e_next = nx < ns ? tab[nx] : 0ull;
                const u32 lastl = lng_k ? k : (k ? k - 1 : 0u);           // the stream continues behind this lane's literals
                const u32 nsb = __shfl_sync(FULL, lpos + L, lastl) & ~15u;
                sbuf ^= 1u; sb = nsb;
                u8* st = stage0 + sbuf * V7_STAGE;
                if (nsb + lane * 16u < clenA) v7_cp16(st + lane * 16u, srcA + nsb + lane * 16u);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // the offset rule (lz4/lz4.c:1764 checkOffset; offset 0 rejected as in the oracle): a bad one makes the chunk bad
            const bool off_bad = valid && M && (off == 0 || off > op - A + L);
            if (__any_sync(FULL, off_bad && lane <= k)) bad = true;
            if (k) {
                const bool act = lane < k;
                const u32 opend = __shfl_sync(FULL, opn, k - 1);
                // ---- window: slide when the group might not fit ----
                if (op0 - win_base > V7_WIN - V7_SPAN) {
                    const u32 nb = (op0 > V7_KEEP ? op0 - V7_KEEP : 0u) & ~15u;   // <= flushed: flushes leave less than 16 bytes
                    const u32 sh = nb - win_base;
                    const uint4 b0 = *reinterpret_cast<const uint4*>(win + sh + lane * 16u);
                    const uint4 b1 = *reinterpret_cast<const uint4*>(win + sh + 512u + lane * 16u);
                    __syncwarp();
                    *reinterpret_cast<uint4*>(win + lane * 16u) = b0;
                    *reinterpret_cast<uint4*>(win + 512u + lane * 16u) = b1;
                    win_base = nb;
                    __syncwarp();
                }
                // ---- literals: every run is independent ----
                {
                    const u32 Lc = act ? L : 0u;
                    const u32 maxL = __reduce_max_sync(FULL, Lc);
                    u8* w = win + (op - win_base);
                    const u8* s = lit;
                    for (u32 i = 0; i < maxL; i += 4) {
                        u32 v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                        if (i + 0 < Lc) v0 = s[i + 0];
                        if (i + 1 < Lc) v1 = s[i + 1];
                        if (i + 2 < Lc) v2 = s[i + 2];
                        if (i + 3 < Lc) v3 = s[i + 3];
                        if (i + 0 < Lc) w[i + 0] = (u8)v0;
                        if (i + 1 < Lc) w[i + 1] = (u8)v1;
                        if (i + 2 < Lc) w[i + 2] = (u8)v2;
                        if (i + 3 < Lc) w[i + 3] = (u8)v3;
                    }
                }
                // ---- matches ----
                const bool hasm = act && M != 0 && !off_bad;
                const bool ovl = off < M;                               // overlapping: periodic extension, copied serially by its lane
                const u32 dstm = op + L;
                const u32 key = act ? op : 0xFFFFFFFFu;                 // sorted over the lanes of the group
                u32 shift = off;
                u32 ja = 0;                                             // lane whose sequence holds the first source byte
                const bool fwdable = hasm && !ovl;
                // source forwarding (pointer jumping): see the header
                const bool reach = hasm && dstm - off + (ovl ? off : M) > op0;   // the source reaches into the group
                unsigned depmask = 0;
                if (__any_sync(FULL, reach)) {
                    const u32 pk = act ? (dstm - op0) | ((opn - op0) << 12) | ((u32)fwdable << 24) : 0u;     // spans <= 2048: 12 bits each
                    for (u32 it = 0; it < 6; it++) {
                        const u32 s = dstm - shift;
                        ja = 0;
                        #pragma unroll
                        for (u32 st = 16; st; st >>= 1) {
                            const u32 t = __shfl_sync(FULL, key, ja | st);
                            if (t <= s) ja |= st;
                        }
                        const u32 pj = __shfl_sync(FULL, pk, ja);
                        const u32 sj = __shfl_sync(FULL, shift, ja);
                        const u32 dj = op0 + (pj & 0xFFFu), ej = op0 + ((pj >> 12) & 0xFFFu);
                        const bool can = fwdable && s >= op0 && (pj >> 24) && s >= dj && s + M <= ej;
                        if (!__any_sync(FULL, can)) break;
                        if (can) shift += sj;
                    }
                    // dependencies: earlier matches of the group whose output intersects [s, e)
                    const u32 s = dstm - shift, e2 = s + (ovl ? off : M);
                    ja = 0; u32 jb = 0;
                    #pragma unroll
                    for (u32 st = 16; st; st >>= 1) {
                        const u32 ta = __shfl_sync(FULL, key, ja | st);
                        const u32 tb = __shfl_sync(FULL, key, jb | st);
                        if (ta <= s) ja |= st;
                        if (tb <= e2 - 1) jb |= st;
                    }
                    const unsigned mb = __ballot_sync(FULL, hasm);
                    const u32 da = __shfl_sync(FULL, dstm, ja), db = __shfl_sync(FULL, dstm, jb);
                    if (hasm && e2 > op0) {
                        unsigned range = (jb >= 31 ? FULL : ((2u << jb) - 1u)) & ~((1u << ja) - 1u);
                        if (da >= e2) range &= ~(1u << ja);
                        if (db >= e2) range &= ~(1u << jb);
                        depmask = range & mb & lt_mask;
                    }
                }
                __syncwarp();                                           // literals are in the window
                const u32 src_pos = dstm - shift;
                bool pend = hasm;
                for (;;) {
                    const unsigned pb = __ballot_sync(FULL, pend);
                    if (!pb) break;
                    const bool rdy = pend && !(depmask & pb);
                    u32 rem = (rdy && !ovl) ? M : 0u;
                    u32 a = src_pos;
                    u8* w = win + (dstm - win_base);
                    while (__any_sync(FULL, rem != 0)) {
                        if (rem) {
                            u32 w0, w1, w2, w3;
                            if (a >= win_base) {
                                const uint2* q = reinterpret_cast<const uint2*>(win + ((a - win_base) & ~7u));
                                const uint2 x = q[0], y = q[1];
                                w0 = x.x; w1 = x.y; w2 = y.x; w3 = y.y;
                            } else {
                                const u8* q = dstv + (a & ~7u);
                                const u64 x = ldcg_u64(q), y = ldcg_u64(q + 8);
                                w0 = (u32)x; w1 = (u32)(x >> 32); w2 = (u32)y; w3 = (u32)(y >> 32);
                            }
                            const bool hi = (a & 4u) != 0;
                            const u32 x0 = hi ? w1 : w0, x1 = hi ? w2 : w1, x2 = hi ? w3 : w2;
                            const u32 shb = (a & 3u) * 8u;
                            const u32 r0 = __funnelshift_r(x0, x1, shb), r1 = __funnelshift_r(x1, x2, shb);
                            if (rem > 0) w[0] = (u8)r0;
                            if (rem > 1) w[1] = (u8)(r0 >> 8);
                            if (rem > 2) w[2] = (u8)(r0 >> 16);
                            if (rem > 3) w[3] = (u8)(r0 >> 24);
                            if (rem > 4) w[4] = (u8)r1;
                            if (rem > 5) w[5] = (u8)(r1 >> 8);
                            if (rem > 6) w[6] = (u8)(r1 >> 16);
                            if (rem > 7) w[7] = (u8)(r1 >> 24);
                            rem = rem > 8 ? rem - 8 : 0u;
                            a += 8; w += 8;
                        }
                    }
                    if (rdy && ovl) {                                   // source and target overlap: byte by byte (same lane: ordered)
                        volatile u8* wv = win + (dstm - win_base);
                        for (u32 i = 0; i < M; i++) wv[i] = wv[(int)i - (int)off];
                    }
                    pend = pend && !rdy;
                    __syncwarp();
                }
                v7_flush(dstv, win, win_base, flushed, opend, false, lane);
                cur += k;
            }
            // ---- a long sequence: whole warp, through global memory ----
            if (lng_k) {
                const u32 l_op = __shfl_sync(FULL, op, k), l_src = __shfl_sync(FULL, lsrc, k), l_L = __shfl_sync(FULL, L, k);
                const u32 l_off = __shfl_sync(FULL, off, k), l_end = __shfl_sync(FULL, opn, k);
                const u32 l_M = __shfl_sync(FULL, off_bad ? 0u : M, k);
                v7_flush(dstv, win, win_base, flushed, l_op, true, lane);
                __syncwarp();
                if (l_L) { coop_copy(dstv + l_op, src + l_src, l_L, lane, 32); __syncwarp(); }
                if (l_M) { v2_match_copy<32>(dstv, l_op + l_L, l_off, l_M, lane, FULL); __syncwarp(); }
                flushed = l_end;
                cur += 1;
                if (cur < ns) {                                         // history for what follows
                    win_base = (l_end > V7_KEEP ? l_end - V7_KEEP : 0u) & ~15u;
                    for (u32 x = win_base + lane * 16u; x < l_end; x += 512u) {
                        if (x + 16u <= l_end) *reinterpret_cast<uint4*>(win + (x - win_base)) = ldcg_u128(dstv + x);
                        else for (u32 b = x; b < l_end; b++) win[b - win_base] = *(volatile u8*)(dstv + b);
                    }
                    __syncwarp();
                }
            }
        }
        v7_flush(dstv, win, win_base, flushed, n + A, true, lane);
        __syncwarp();
        if (bad && lane == 0) {
            P.status[c] = ST_OFFSET;
            if (P.bad_chunks) atomicAdd(P.bad_chunks, 1u);
        }
    }
}

inline cudaError_t v7_launch_parse(const V7Params& P, cudaStream_t st, int lanes = 32, bool big = false)
{
    // P.bucket points at the slice's zero block: [0, 2 V7_BUCKETS) counters, then team_count, ticket of pass 2, ticket of the team decoder
    cudaError_t e = cudaMemsetAsync(P.bucket, 0, V7_ZERO_WORDS * sizeof(u32), st);
    if (e != cudaSuccess) return e;
    if (big) parse_v7_kernel<32, true><<<(P.n_chunks + 31) / 32, 32, 0, st>>>(P);
    else if (lanes == 8) parse_v7_kernel<8, false><<<(P.n_chunks + 7) / 8, 32, 0, st>>>(P);
    else if (lanes == 16) parse_v7_kernel<16, false><<<(P.n_chunks + 15) / 16, 32, 0, st>>>(P);
    else parse_v7_kernel<32, false><<<(P.n_chunks + 31) / 32, 32, 0, st>>>(P);
    order_v7_kernel<<<(P.n_chunks + 255) / 256, 256, 0, st>>>(P);
    return cudaGetLastError();
}

template <bool BIG>
inline cudaError_t v7_launch_decompress_t(const V7Params& P, int sm_count, cudaStream_t st)
{
    int per_sm = 0;
    const size_t smem = (size_t)V7_WARPS * V7_WARP_SMEM;
    cudaError_t e = cudaFuncSetAttribute(decompress_v7_kernel<BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: every launch
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v7_kernel<BIG>, V7_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (P.n_chunks + V7_WARPS - 1) / V7_WARPS;
    if (grid > need) grid = need;
    decompress_v7_kernel<BIG><<<grid, V7_WARPS * 32, smem, st>>>(P);               // (P.ticket was cleared with the zero block)
    return cudaGetLastError();
}

inline cudaError_t v7_launch_decompress(const V7Params& P, int sm_count, cudaStream_t st, bool big = false)
{
    return big ? v7_launch_decompress_t<true>(P, sm_count, st) : v7_launch_decompress_t<false>(P, sm_count, st);
}

}  // namespace lz4b200
