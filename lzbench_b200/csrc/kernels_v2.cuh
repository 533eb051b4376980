// kernels_v2.cuh -- generation-2 decode: sub-warp teams.
//
// Measurements that led here (profiles/r01_*): the serial token chain of a block
// (lz4/lz4.c:1929-2151) makes a warp-per-chunk decoder (v0) issue-bound -- ~114 warp
// instructions per sequence, every one of them spent on ONE sequence of 10-20 bytes -- while the
// whole-CTA tile decoder (v1) is bound by true match-dependency chains (60-290 deep per 64 KB
// chunk on text / binaries / record data).  The cheapest way to spend fewer issue slots per
// sequence without needing parallelism inside a chunk is to put several chunks in one warp:
// a TEAM of T lanes (T = 4/8/16/32) owns a chunk and runs its token chain serially, the 32/T
// teams of a warp advance in lock step, one sequence per loop iteration each, so every warp
// instruction now works on 32/T sequences.  T lanes are enough for the copies because the
// typical literal run and match are shorter than 16 bytes; copies of >= LONG bytes are handed
// to the whole warp (16-byte vector path) so incompressible and run-length data still move at
// memcpy speed.  Teams take chunks from a global ticket, so ragged chunk lists balance.
//
// All bounds and end-of-block checks of LZ4_decompress_safe are kept; status codes as in v0.
#pragma once
#include "common.cuh"

namespace lz4b200 {

constexpr u32 V2_LONG = 96;                      // copies this long are executed by the whole warp
constexpr int V2_BLOCK = 128;

// inverse table for j % off with off < 32:  q = (j * inv) >> 16 is exact for j < 32
__constant__ u16 v2_inv[32] = {0, 65535, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042,
                               4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428,
                               2341, 2260, 2185, 2115};

__device__ __forceinline__ u32 v2_mod_small(u32 j, u32 off)     // j < 64, 1 <= off < 32
{
    u32 r = j - ((j * (u32)v2_inv[off]) >> 16) * off;
    if (r >= off) r -= off;                                        // the 16-bit reciprocal may be one short
    return r;
}

// Overlap-aware match copy by a group of `G` lanes (g = lane index in group, `mask` = its lanes).
// dst may alias src; every lane of the group calls it with identical arguments.
template <u32 G>
__device__ __forceinline__ void v2_match_copy(u8* out, u32 op, u32 off, u32 ml, u32 g, unsigned mask)
{
    const u8* m = out + op - off;
    if (off >= ml) {                                   // no overlap
        for (u32 i = g; i < ml; i += G) out[op + i] = m[i];
    } else if (off >= G) {                             // a G-byte step never reads its own writes
        for (u32 i = 0; i < ml; i += G) {
            if (i + g < ml) out[op + i + g] = m[i + g];
            __syncwarp(mask);
        }
    } else {                                           // periodic extension of the last `off` bytes
        u32 r = v2_mod_small(g, off);
        const u32 step = v2_mod_small(G, off);
        for (u32 i = g; i < ml; i += G) {
            out[op + i] = m[r];
            r += step; if (r >= off) r -= off;
        }
    }
}

template <u32 T>
__global__ void __launch_bounds__(V2_BLOCK)
decompress_v2_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ csz,
                     const u64* __restrict__ out_off, const u64* __restrict__ usz, u8* out,
                     u32* __restrict__ status, u32* __restrict__ bad_chunks, u32 n_chunks, u32* ticket)
{
    const u32 lane = lane_id();
    const u32 j = lane & (T - 1);                      // lane in team
    const u32 tbase = lane & ~(T - 1);                 // first lane of my team
    const unsigned tmask = T == 32 ? FULL : (((1u << (T & 31)) - 1u) << tbase);

    // team state (identical in every lane of the team)
    bool busy = false, exhausted = false;
    const u8* src = nullptr;
    u8* dst = nullptr;
    u32 clen = 0, n = 0, ip = 0, op = 0, chunk = 0;
    u32 tok = 0;                                       // token at src[ip], loaded one sequence ahead

    for (;;) {
        // ---- between chunks: take the next one from the ticket ----
        if (!busy && !exhausted) {
            u32 c = 0;
            if (j == 0) c = atomicAdd(ticket, 1u);
            c = __shfl_sync(tmask, c, tbase);
            if (c >= n_chunks) exhausted = true;
            else {
                chunk = c;
                src = in + in_off[c]; clen = (u32)csz[c];
                dst = out + out_off[c]; n = (u32)usz[c];
                ip = 0; op = 0; busy = true;
                tok = clen ? __ldg(src) : 0;
            }
        }
        if (__all_sync(FULL, exhausted)) break;

        // copies this team hands to the whole warp this iteration: a literal run (or raw chunk), then a match
        u32 wl_len = 0, wl_ip = 0, wl_op = 0;          // wide literal copy: src[wl_ip..] -> dst[wl_op..]
        u32 wm_len = 0, wm_op = 0, wm_off = 0;         // wide match copy at dst[wm_op..]
        u32 err = ST_OK;
        bool finished = false;

        if (busy) {
            if (clen == n) {                           // raw chunk (_lzbench/lzbench.cpp:311-315)
                wl_len = n; finished = true;
            } else if (n == 0 || clen == 0) {          // lz4/lz4.c:1776-1782
                err = (n == 0 && clen == 1 && __ldg(src) == 0) ? ST_OK : ST_EMPTY; finished = true;
            } else if (ip >= clen) {
                err = ST_TOKEN_OOB; finished = true;
            } else if ((tok >> 4) < RUN_MASK && (tok & ML_MASK) < ML_MASK &&
                       op + (tok >> 4) + (tok & ML_MASK) + MINMATCH + MFLIMIT <= n &&
                       ip + 1 + (tok >> 4) + (2 + 1 + LASTLITERALS) <= clen) {
                // ---- fast path: short literal run, short match, far from both ends.  Straight-line and
                // predicated so that the teams of a warp stay converged and their loads overlap. ----
                constexpr u32 NL = (16 + T - 1) / T;       // literal run (<= 14) plus the 2 offset bytes
                constexpr u32 NM = (18 + T - 1) / T;       // match length <= 18
                const u32 L = tok >> 4, ml = (tok & ML_MASK) + MINMATCH;
                const u32 lit0 = ip + 1, cpy = op + L, nip = lit0 + L + 2;
                u32 x = 0;
                #pragma unroll
                for (u32 k = 0; k < NL; k++) {
                    const u32 idx = j + k * T;
                    u32 b = 0;
                    if (idx < L + 2) b = __ldg(src + lit0 + idx);
                    if (idx < L) dst[op + idx] = (u8)b;
                    x |= b << (8 * k);
                }
                const u32 ntok = __ldg(src + nip);         // next token (or first length byte), one sequence ahead
                const u32 lo = (__shfl_sync(tmask, x, tbase + (L & (T - 1))) >> (8 * (L / T))) & 0xFFu;
                const u32 hi = (__shfl_sync(tmask, x, tbase + ((L + 1) & (T - 1))) >> (8 * ((L + 1) / T))) & 0xFFu;
                const u32 off = lo | (hi << 8);
                if (off == 0 || off > cpy) { err = ST_OFFSET; finished = true; }
                else {
                    __syncwarp(tmask);                     // literals visible to the team
                    if (off >= ml) {
                        const u8* m = dst + cpy - off;
                        u32 v[NM];
                        #pragma unroll
                        for (u32 k = 0; k < NM; k++) { const u32 i = j + k * T; v[k] = i < ml ? m[i] : 0; }
                        #pragma unroll
                        for (u32 k = 0; k < NM; k++) { const u32 i = j + k * T; if (i < ml) dst[cpy + i] = (u8)v[k]; }
                    } else {
                        v2_match_copy<T>(dst, cpy, off, ml, j, tmask);
                    }
                    __syncwarp(tmask);
                    ip = nip; op = cpy + ml; tok = ntok;
                }
            } else {
                // ---- one sequence, general path ----
                const u32 token = tok; ip++;
                u64 len = token >> 4;                  // 64-bit: a hostile 255-run cannot wrap it
                if (len == RUN_MASK) {
                    if (ip + RUN_MASK >= clen) err = ST_LITLEN_OOB;
                    else {
                        u32 s;
                        do {
                            s = __ldg(src + ip); ip++; len += s;
                            if (s == 255 && ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                        } while (s == 255);
                    }
                }
                const u64 cpy = op + len;
                if (err) finished = true;
                else if (cpy + MFLIMIT > n || ip + len + (2 + 1 + LASTLITERALS) > clen) {     // last sequence
                    if (ip + len != clen || cpy > n) err = ST_LAST_LITERALS;
                    else {
                        const u32 L = (u32)len;
                        if (L >= V2_LONG) { wl_len = L; wl_ip = ip; wl_op = op; }
                        else for (u32 i = j; i < L; i += T) dst[op + i] = __ldg(src + ip + i);
                        op = (u32)cpy;
                        if (op != n) err = ST_SIZE;
                    }
                    finished = true;
                } else {
                    const u32 L = (u32)len;
                    u32 off;
                    if (L >= V2_LONG) {
                        wl_len = L; wl_ip = ip; wl_op = op;
                        off = __ldg(src + ip + L) | (__ldg(src + ip + L + 1) << 8);
                    } else if (L + 2 <= T) {           // literals and the offset in one load
                        u32 b = 0;
                        if (j < L + 2) b = __ldg(src + ip + j);
                        if (j < L) dst[op + j] = (u8)b;
                        off = __shfl_sync(tmask, b, tbase + L) | (__shfl_sync(tmask, b, tbase + L + 1) << 8);
                    } else {
                        for (u32 i = j; i < L; i += T) dst[op + i] = __ldg(src + ip + i);
                        off = __ldg(src + ip + L) | (__ldg(src + ip + L + 1) << 8);
                    }
                    ip += L + 2; op = (u32)cpy;
                    u64 ml64 = token & ML_MASK;
                    if (ml64 == ML_MASK) {
                        u32 s;
                        do {
                            s = __ldg(src + ip); ip++; ml64 += s;
                            if (ip + LASTLITERALS - 1 >= clen) { err = ST_MATCHLEN_OOB; break; }
                        } while (s == 255);
                    }
                    ml64 += MINMATCH;
                    if (!err && (off == 0 || off > op)) err = ST_OFFSET;
                    if (!err && op + ml64 + LASTLITERALS > n) err = ST_MATCH_END;
                    if (err) { finished = true; wl_len = 0; }
                    else {
                        const u32 ml = (u32)ml64;
                        if (wl_len == 0 && ml < V2_LONG) {
                            __syncwarp(tmask);         // the literals above are visible to the team
                            v2_match_copy<T>(dst, op, off, ml, j, tmask);
                            __syncwarp(tmask);
                        } else {                        // long match, or a long literal run must land first
                            wm_len = ml; wm_op = op; wm_off = off;
                        }
                        op += ml;
                        tok = ip < clen ? __ldg(src + ip) : 0;
                    }
                }
            }
        }

        // ---- wide phase: the whole warp executes the long copies, team by team ----
        if (__any_sync(FULL, (wl_len | wm_len) != 0)) {
            unsigned todo = __ballot_sync(FULL, j == 0 && (wl_len | wm_len) != 0);
            while (todo) {
                const int l = __ffs(todo) - 1; todo &= todo - 1;
                const u64 sp = __shfl_sync(FULL, (u64)reinterpret_cast<uintptr_t>(src), l);
                const u64 dp = __shfl_sync(FULL, (u64)reinterpret_cast<uintptr_t>(dst), l);
                const u32 a_len = __shfl_sync(FULL, wl_len, l), a_ip = __shfl_sync(FULL, wl_ip, l),
                          a_op = __shfl_sync(FULL, wl_op, l);
                const u32 b_len = __shfl_sync(FULL, wm_len, l), b_op = __shfl_sync(FULL, wm_op, l),
                          b_off = __shfl_sync(FULL, wm_off, l);
                u8* d = reinterpret_cast<u8*>(dp);
                if (a_len) {
                    coop_copy(d + a_op, reinterpret_cast<const u8*>(sp) + a_ip, a_len, lane, 32);
                    __syncwarp();
                }
                if (b_len) {
                    v2_match_copy<32>(d, b_op, b_off, b_len, lane, FULL);
                    __syncwarp();
                }
            }
        }

        if (finished) {
            if (j == 0) {
                status[chunk] = err;
                if (err && bad_chunks) atomicAdd(bad_chunks, 1u);
            }
            busy = false;
        }
    }
}

template <u32 T>
inline cudaError_t v2_launch_t(const u8* in, const u64* in_off, const u64* csz, const u64* out_off, const u64* usz,
                               u8* out, u32* status, u32* bad, u32 n, u32* ticket, int sm_count, cudaStream_t st)
{
    const u32 teams_per_block = V2_BLOCK / T;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v2_kernel<T>, V2_BLOCK, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (n + teams_per_block - 1) / teams_per_block;
    if (grid > need) grid = need;
    e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    decompress_v2_kernel<T><<<grid, V2_BLOCK, 0, st>>>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket);
    return cudaGetLastError();
}

inline cudaError_t v2_launch_decompress(int team, const u8* in, const u64* in_off, const u64* csz, const u64* out_off,
                                        const u64* usz, u8* out, u32* status, u32* bad, u32 n, u32* ticket,
                                        int sm_count, cudaStream_t st)
{
    switch (team) {
    case 4:  return v2_launch_t<4>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    case 16: return v2_launch_t<16>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    case 32: return v2_launch_t<32>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    default: return v2_launch_t<8>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    }
}

}  // namespace lz4b200
