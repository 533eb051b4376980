// kernels_v7.cuh -- generation-7 decode: a sequence table first, then one LANE per sequence.
//
// Every earlier generation walks the token chain of a block (lz4/lz4.c:1797-2151) with a whole team of lanes:
// v5 spends 155 warp instructions per iteration on 4 sequences of ~16 bytes and is issue-bound (profiles/r01_*).
// The chain itself is cheap -- token -> literal length -> next token -- it is the copies that need the lanes.  So:
//
//  1. parse_v7_kernel: one lane per CHUNK walks the chain, validates every rule the reference decoder enforces
//     (the same checks, in the same order, as the general path of v5) and writes one 8-byte entry per sequence:
//     output position, position of the literals in the stream, offset, literal count.  32 chunks per warp
//     instruction; the kernel is latency-bound (a few warps per SM) and leaves the machine to its neighbour.
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
constexpr int V7_WARPS = 8;                // warps per CTA of the lane-per-sequence kernel
constexpr u32 V7_ROUTE_NONE = 0, V7_ROUTE_SEQ = 1, V7_ROUTE_TEAM = 2;

// entries of the table a chunk may need: sequences (>= 4 output bytes each but the last), literal pieces, terminator
__host__ __device__ inline u64 v7_entries_for(u64 out_bytes, u64 chunks)
{
    return out_bytes / 4 + out_bytes / V7_LPIECE + 3 * chunks;
}

struct V7Params {
    const u8* in; const u64* in_off; const u64* csz; const u64* out_off; const u64* usz; u8* out;
    u32* status; u32* bad_chunks; u32* ticket; u32 n_chunks;
    u32 cap_all;                           // see V5Params::cap_all
    u64* table; u64 table_cap;             // sequence table of this slice and its capacity in entries
    u32* nseq;                             // entries per chunk (without the terminator)
    u32* route;                            // V7_ROUTE_* per chunk
    u32* team_list; u32* team_count;       // chunks left to the team decoder
};

// entry: bits 0-16 output position of the sequence, 17-33 stream position of its literals, 34-49 offset (0: no match),
//        50-62 literal count
__device__ __forceinline__ u64 v7_pack(u32 op, u32 lsrc, u32 off, u32 L)
{
    return (u64)op | ((u64)lsrc << 17) | ((u64)off << 34) | ((u64)L << 50);
}

// ---- pass 1: the token chain, one lane per chunk ---------------------------------------------------------------
__global__ void __launch_bounds__(32)
parse_v7_kernel(const __grid_constant__ V7Params P)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_chunks) return;
    const u8* __restrict__ src = P.in + P.in_off[c];
    const u64 clen64 = P.csz[c], n64 = P.usz[c];
    const u64 rel = P.out_off[c] - P.out_off[0];
    const u64 base = v7_entries_for(rel, c);
    u32 err = ST_OK, route = V7_ROUTE_SEQ, k = 0;
    if (clen64 == n64) {
        route = n64 >= 4096 ? V7_ROUTE_TEAM : V7_ROUTE_SEQ;            // raw chunk (_lzbench/lzbench.cpp:311-315): a plain copy
    } else if (n64 > V7_MAX_CHUNK || clen64 > 0x1FFFFu || base + v7_entries_for(n64, 1) > P.table_cap) {
        route = V7_ROUTE_TEAM;
    } else {
        const u32 clen = (u32)clen64, n = (u32)n64;
        const u32 capv = P.cap_all > n ? P.cap_all : n;
        u64* __restrict__ tab = P.table + base;
        u32 ip = 0, op = 0;
        if (n == 0 || clen == 0) {                                       // lz4/lz4.c:1776-1782
            err = (n == 0 && clen == 1 && src[0] == 0) ? ST_OK : ST_EMPTY;
            route = V7_ROUTE_NONE;
        } else for (;;) {
            if (ip >= clen) { err = ST_TOKEN_OOB; break; }
            const u32 token = src[ip++];
            u32 len = token >> 4;
            if (len == RUN_MASK) {                                       // lz4/lz4.c:1707-1729 read_variable_length
                if (ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                u32 s;
                do {
                    s = src[ip++]; len += s;
                    if (s == 255 && ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                } while (s == 255);
                if (err) break;
            }
            const u32 cpy = op + len;                                    // len < 255 * clen: no wrap
            const bool last = cpy + MFLIMIT > capv || ip + len + (2 + 1 + LASTLITERALS) > clen;
            if (last && (ip + len != clen || cpy > n)) { err = ST_LAST_LITERALS; break; }
            if (cpy > n) { err = ST_MATCH_END; break; }                  // (frame capacity > n) rejected below anyway; keeps the entry count bounded by n
            u32 lsrc = ip, L = len, lop = op;
            while (L > V7_LPIECE) {                                      // literal-only pieces
                tab[k++] = v7_pack(lop, lsrc, 0, V7_LPIECE);
                lop += V7_LPIECE; lsrc += V7_LPIECE; L -= V7_LPIECE;
            }
            if (last) {
                tab[k++] = v7_pack(lop, lsrc, 0, L);
                op = cpy;
                if (op != n) err = ST_SIZE;
                break;
            }
            const u32 off = (u32)src[ip + len] | ((u32)src[ip + len + 1] << 8);
            ip += len + 2; op = cpy;
            u32 ml = token & ML_MASK;
            if (ml == ML_MASK) {
                u32 s;
                do {
                    s = src[ip++]; ml += s;
                    if (ip + LASTLITERALS - 1 >= clen) { err = ST_MATCHLEN_OOB; break; }
                } while (s == 255);
                if (err) break;
            }
            ml += MINMATCH;
            if (off == 0 || off > op) { err = ST_OFFSET; break; }
            if (op + ml + LASTLITERALS > capv || op + ml > n) { err = ST_MATCH_END; break; }
            tab[k++] = v7_pack(lop, lsrc, off, L);
            op += ml;
        }
        if (!err && route == V7_ROUTE_SEQ) {
            tab[k] = v7_pack(n, 0, 0, 0);                                // terminator: carries the end position
            if ((u64)k * 256 < n) route = V7_ROUTE_TEAM;                 // mostly long copies: the team decoder's case
        }
    }
    if (err) route = V7_ROUTE_NONE;
    P.status[c] = err;
    P.nseq[c] = k;
    P.route[c] = route;
    if (err && P.bad_chunks) atomicAdd(P.bad_chunks, 1u);
    if (route == V7_ROUTE_TEAM) P.team_list[atomicAdd(P.team_count, 1u)] = c;
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

__global__ void __launch_bounds__(V7_WARPS * 32, 3)
decompress_v7_kernel(const __grid_constant__ V7Params P)
{
    __shared__ __align__(16) u8 win_all[V7_WARPS][V7_WIN + 32];
    const u32 lane = lane_id();
    u8* const win = win_all[threadIdx.x >> 5];
    const unsigned lt_mask = (1u << lane) - 1u;

    for (;;) {
        u32 c = 0;
        if (lane == 0) c = atomicAdd(P.ticket, 1u);
        c = __shfl_sync(FULL, c, 0);
        if (c >= P.n_chunks) break;
        if (P.route[c] != V7_ROUTE_SEQ) continue;
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

        while (cur < ns) {
            const u32 idx = cur + lane;
            const bool valid = idx < ns;
            const u64 e0 = valid ? tab[idx] : 0ull, e1 = valid ? tab[idx + 1] : 0ull;
            const u32 op = ((u32)e0 & 0x1FFFFu) + A;
            const u32 lsrc = (u32)(e0 >> 17) & 0x1FFFFu;
            const u32 off = (u32)(e0 >> 34) & 0xFFFFu;
            const u32 L = (u32)(e0 >> 50) & 0x1FFFu;
            const u32 opn = ((u32)e1 & 0x1FFFFu) + A;
            const u32 M = valid ? opn - op - L : 0u;
            const u32 op0 = __shfl_sync(FULL, op, 0);
            const bool lng = valid && (L > V7_CAP || M > V7_CAP);
            const unsigned stop = __ballot_sync(FULL, !valid || lng || opn - op0 > V7_SPAN);
            const u32 k = stop ? (u32)__ffs(stop) - 1u : 32u;          // lanes [0, k) form the group
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
                    const u8* s = src + lsrc;
                    for (u32 i = 0; i < maxL; i += 4) {
                        u32 v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                        if (i + 0 < Lc) v0 = __ldg(s + i + 0);
                        if (i + 1 < Lc) v1 = __ldg(s + i + 1);
                        if (i + 2 < Lc) v2 = __ldg(s + i + 2);
                        if (i + 3 < Lc) v3 = __ldg(s + i + 3);
                        if (i + 0 < Lc) w[i + 0] = (u8)v0;
                        if (i + 1 < Lc) w[i + 1] = (u8)v1;
                        if (i + 2 < Lc) w[i + 2] = (u8)v2;
                        if (i + 3 < Lc) w[i + 3] = (u8)v3;
                    }
                }
                // ---- matches ----
                const bool hasm = act && M != 0;
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
                    const u32 s = dstm - shift, e = s + (ovl ? off : M);
                    ja = 0; u32 jb = 0;
                    #pragma unroll
                    for (u32 st = 16; st; st >>= 1) {
                        const u32 ta = __shfl_sync(FULL, key, ja | st);
                        const u32 tb = __shfl_sync(FULL, key, jb | st);
                        if (ta <= s) ja |= st;
                        if (tb <= e - 1) jb |= st;
                    }
                    const unsigned mb = __ballot_sync(FULL, hasm);
                    const u32 da = __shfl_sync(FULL, dstm, ja), db = __shfl_sync(FULL, dstm, jb);
                    if (hasm && e > op0) {
                        const u32 lo = s >= op0 ? ja : 0u;
                        unsigned range = (jb >= 31 ? FULL : ((2u << jb) - 1u)) & ~((1u << lo) - 1u);
                        if (da >= e) range &= ~(1u << lo);
                        if (db >= e) range &= ~(1u << jb);
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
            if (k < 32 && (stop >> k & 1u) && __shfl_sync(FULL, (u32)lng, k)) {
                const u32 l_op = __shfl_sync(FULL, op, k), l_src = __shfl_sync(FULL, lsrc, k), l_L = __shfl_sync(FULL, L, k);
                const u32 l_M = __shfl_sync(FULL, M, k), l_off = __shfl_sync(FULL, off, k), l_end = __shfl_sync(FULL, opn, k);
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
    }
}

inline cudaError_t v7_launch_parse(const V7Params& P, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(P.team_count, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    parse_v7_kernel<<<(P.n_chunks + 31) / 32, 32, 0, st>>>(P);
    return cudaGetLastError();
}

inline cudaError_t v7_launch_decompress(const V7Params& P, int sm_count, cudaStream_t st)
{
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v7_kernel, V7_WARPS * 32, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (P.n_chunks + V7_WARPS - 1) / V7_WARPS;
    if (grid > need) grid = need;
    e = cudaMemsetAsync(P.ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    decompress_v7_kernel<<<grid, V7_WARPS * 32, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace lz4b200
