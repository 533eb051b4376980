// kernels_v3.cuh -- generation-3 decode: sub-warp teams (as v2) with the per-sequence critical path cut down.
//
// What the v2 profile showed (profiles/r01_v2t8_*): with every chunk of a 1 GiB batch already in flight the run
// time is the latency of one chunk's serial chain, ~1.5 us per sequence, of which ~850 cycles are the store of
// a match waiting for its own source load (L2 round trip), ~290 the literal load, the rest instruction issue.
// v3 keeps the team structure and removes those waits from the chain:
//   * the 16 bytes after the current token (literal run + offset) are loaded one sequence AHEAD together with
//     the token, so decoding a sequence needs no load result at all (offset by shuffle from the window);
//   * the match is software-pipelined: sequence k issues its source loads, the stores happen during sequence
//     k+1 -- unless k+1's match would read what k has not stored yet (or its own literals / itself), a cheap
//     team-uniform range test, in which case the pending stores are flushed first;
//   * no memory barrier on the fast path.  __syncwarp() drains the long scoreboard (52 % of stall samples sat on
//     it), so nothing could stay in flight across it.  Instead every output byte belongs to a fixed lane of the
//     team -- position mod T -- which both stores it and later loads it (a thread always sees its own stores);
//     bytes move between lanes in registers, by shuffle: literals rotate from window order to position order,
//     match bytes rotate by (offset mod T).  Only the rare unmapped copies (overlapping or long matches, the
//     general path) are bracketed by __syncwarp;
//   * chunk hand-out, warp-wide long copies and chunk completion sit behind ONE vote per iteration.
// Format checks are unchanged (LZ4_decompress_safe rules, lz4/lz4.c:1929-2151); the general path is v2's.
#pragma once
#include "common.cuh"
#include "kernels_v2.cuh"

namespace lz4b200 {

constexpr u32 V3_AHEAD = 1024;                   // input prefetch distance in bytes

template <u32 T>
__global__ void __launch_bounds__(V2_BLOCK)
decompress_v3_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ csz,
                     const u64* __restrict__ out_off, const u64* __restrict__ usz, u8* out,
                     u32* __restrict__ status, u32* __restrict__ bad_chunks, u32 n_chunks, u32* ticket)
{
    const u32 lane = lane_id();
    const u32 j = lane & (T - 1);
    const u32 tbase = lane & ~(T - 1);
    const unsigned tmask = T == 32 ? FULL : (((1u << (T & 31)) - 1u) << tbase);
    constexpr u32 NX = (16 + T - 1) / T;               // window bytes held per lane (16 bytes after the token)
    constexpr u32 NM = (18 + T - 1) / T;               // bytes of a short match (<= 18) per lane

    bool busy = false, exhausted = false;
    const u8* src = nullptr;
    u8* dst = nullptr;
    u32 clen = 0, n = 0, ip = 0, op = 0, chunk = 0;
    // Loaded bytes are kept one per register and are NOT touched in the iteration that loads them: any ALU op
    // on a load result (even packing two bytes) would stall the warp on the load and undo the pipelining.
    u32 tok = 0;                                       // token at src[ip]
    u32 xb[NX];                                        // window: xb[k] = src[ip + 1 + j + k*T]
    u32 pm_len = 0, pm_dst = 0, pm_off = 0;            // pending match: loaded source bytes pvb[], stored one sequence later
    u32 pvb[NM];
    #pragma unroll
    for (u32 k = 0; k < NX; k++) xb[k] = 0;
    #pragma unroll
    for (u32 k = 0; k < NM; k++) pvb[k] = 0;

#define V3_PREFETCH(POS)                                                                     \
    do {                                                                                     \
        const u32 p_ = (POS);                                                                \
        tok = p_ < clen ? __ldg(src + p_) : 0;                                               \
        _Pragma("unroll") for (u32 k_ = 0; k_ < NX; k_++) {                                  \
            const u32 i_ = p_ + 1 + j + k_ * T;                                              \
            xb[k_] = i_ < clen ? (u32)__ldg(src + i_) : 0u;                                  \
        }                                                                                    \
    } while (0)
#define V3_FLUSH()                                                                           \
    do {                                                                                     \
        if (pm_len) {                                                                        \
            const u32 i0_ = (j - pm_dst) & (T - 1);      /* first match index this lane owns */  \
            const u32 sl_ = tbase + ((j - pm_off) & (T - 1));   /* lane that loaded it */    \
            _Pragma("unroll") for (u32 k_ = 0; k_ < NM; k_++) {                              \
                const u32 val_ = __shfl_sync(tmask, pvb[k_], sl_);                           \
                const u32 i_ = i0_ + k_ * T;                                                 \
                if (i_ < pm_len) dst[pm_dst + i_] = (u8)val_;                                \
            }                                                                                \
            pm_len = 0;                                                                      \
        }                                                                                    \
    } while (0)
// literal run of the current sequence, window order -> position order (lane = output position mod T)
#define V3_LITERALS()                                                                        \
    do {                                                                                     \
        const u32 i0_ = (j - op) & (T - 1);                                                  \
        _Pragma("unroll") for (u32 k_ = 0; k_ < NX; k_++) {                                  \
            const u32 val_ = __shfl_sync(tmask, xb[k_], tbase + i0_);                        \
            const u32 i_ = i0_ + k_ * T;                                                     \
            if (i_ < L) dst[op + i_] = (u8)val_;                                             \
        }                                                                                    \
    } while (0)

    for (;;) {
        u32 wl_len = 0, wl_ip = 0, wl_op = 0;          // warp-wide literal copy request
        u32 wm_len = 0, wm_op = 0, wm_off = 0;         // warp-wide match copy request
        u32 err = ST_OK;
        bool finished = false;

        if (busy) {
            const u32 L = tok >> 4, M = tok & ML_MASK;
            if (L < RUN_MASK && M < ML_MASK && clen != n &&
                op + L + M + MINMATCH + MFLIMIT <= n && ip + 1 + L + (2 + 1 + LASTLITERALS) <= clen) {
                // ---- fast path ----
                const u32 ml = M + MINMATCH, cpy = op + L, nip = ip + 1 + L + 2;
                // offset = window bytes L and L+1: pick the register (team-uniform index), then the lane
                u32 sel_lo = xb[0], sel_hi = xb[0];
                #pragma unroll
                for (u32 k = 1; k < NX; k++) { if (L / T == k) sel_lo = xb[k]; if ((L + 1) / T == k) sel_hi = xb[k]; }
                const u32 lo = __shfl_sync(tmask, sel_lo, tbase + (L & (T - 1)));
                const u32 hi = __shfl_sync(tmask, sel_hi, tbase + ((L + 1) & (T - 1)));
                const u32 off = lo | (hi << 8);
                if (off == 0 || off > cpy) { V3_FLUSH(); err = ST_OFFSET; finished = true; }
                else {
                    // the compressed stream is read strictly forward: pull it towards the SM well ahead of the
                    // token chain (one request per 128-byte line, by one lane)
                    if (j == 0 && ((nip + V3_AHEAD) ^ (ip + V3_AHEAD)) >= 128u && nip + V3_AHEAD < clen)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ((nip + V3_AHEAD) & ~127u)));
                    u32 ntok, nxb[NX];
                    ntok = __ldg(src + nip);            // next token and window, one sequence ahead (nip <= clen - 6)
                    #pragma unroll
                    for (u32 k = 0; k < NX; k++) {
                        const u32 i = nip + 1 + j + k * T;
                        nxb[k] = i < clen ? (u32)__ldg(src + i) : 0u;
                    }
                    // Previous match and this literal run are stored first (each byte by the lane that owns its
                    // position); then this match's source is loaded by the owning lanes -- same-thread read after
                    // write, so no barrier -- and stored during the NEXT sequence.  The only wait for a match load is
                    // the flush one sequence later, i.e. a dependent next sequence stalls, an independent one does not.
                    V3_FLUSH();
                    V3_LITERALS();
                    if (off >= ml) {
                        const u32 sbase = cpy - off;
                        const u32 t0 = (j - sbase) & (T - 1);
                        #pragma unroll
                        for (u32 k = 0; k < NM; k++) { const u32 t = t0 + k * T; pvb[k] = t < ml ? (u32)dst[sbase + t] : 0u; }
                        pm_len = ml; pm_dst = cpy; pm_off = off;
                    } else {                            // the match overlaps itself: periodic copy, unmapped
                        __syncwarp(tmask);
                        v2_match_copy<T>(dst, cpy, off, ml, j, tmask);
                        __syncwarp(tmask);
                    }
                    ip = nip; op = cpy + ml; tok = ntok;
                    #pragma unroll
                    for (u32 k = 0; k < NX; k++) xb[k] = nxb[k];
                }
            } else {
                // ---- general path: one sequence, everything re-read from memory ----
                V3_FLUSH();
                __syncwarp(tmask);
                if (clen == n) {                           // raw chunk (_lzbench/lzbench.cpp:311-315)
                    wl_len = n; finished = true;
                } else if (n == 0 || clen == 0) {          // lz4/lz4.c:1776-1782
                    err = (n == 0 && clen == 1 && __ldg(src) == 0) ? ST_OK : ST_EMPTY; finished = true;
                } else if (ip >= clen) {
                    err = ST_TOKEN_OOB; finished = true;
                } else {
                    const u32 token = tok; ip++;
                    u64 len = token >> 4;
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
                            const u32 LL = (u32)len;
                            if (LL >= V2_LONG) { wl_len = LL; wl_ip = ip; wl_op = op; }
                            else for (u32 i = j; i < LL; i += T) dst[op + i] = __ldg(src + ip + i);
                            op = (u32)cpy;
                            if (op != n) err = ST_SIZE;
                        }
                        finished = true;
                    } else {
                        const u32 LL = (u32)len;
                        if (LL >= V2_LONG) { wl_len = LL; wl_ip = ip; wl_op = op; }
                        else for (u32 i = j; i < LL; i += T) dst[op + i] = __ldg(src + ip + i);
                        const u32 off = __ldg(src + ip + LL) | (__ldg(src + ip + LL + 1) << 8);
                        ip += LL + 2; op = (u32)cpy;
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
                                __syncwarp(tmask);
                                v2_match_copy<T>(dst, op, off, ml, j, tmask);
                                __syncwarp(tmask);
                            } else {
                                wm_len = ml; wm_op = op; wm_off = off;
                            }
                            op += ml;
                            V3_PREFETCH(ip);
                        }
                    }
                }
            }
        }

        // ---- management, behind one vote: long copies by the whole warp, chunk completion, next chunk ----
        if (__any_sync(FULL, !busy || finished || (wl_len | wm_len) != 0)) {
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
            if (finished) {
                if (j == 0) {
                    status[chunk] = err;
                    if (err && bad_chunks) atomicAdd(bad_chunks, 1u);
                }
                busy = false;
            }
            if (!busy && !exhausted) {
                u32 c = 0;
                if (j == 0) c = atomicAdd(ticket, 1u);
                c = __shfl_sync(tmask, c, tbase);
                if (c >= n_chunks) exhausted = true;
                else {
                    chunk = c;
                    src = in + in_off[c]; clen = (u32)csz[c];
                    dst = out + out_off[c]; n = (u32)usz[c];
                    ip = 0; op = 0; busy = true; pm_len = 0;
                    if (j * 128u < clen && j * 128u < V3_AHEAD + 128u)         // first lines of the compressed chunk
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + j * 128u));
                    V3_PREFETCH(0);
                }
            }
            if (__all_sync(FULL, exhausted)) break;
        }
    }
#undef V3_PREFETCH
#undef V3_FLUSH
#undef V3_LITERALS
}

template <u32 T>
inline cudaError_t v3_launch_t(const u8* in, const u64* in_off, const u64* csz, const u64* out_off, const u64* usz,
                               u8* out, u32* status, u32* bad, u32 n, u32* ticket, int sm_count, cudaStream_t st)
{
    const u32 teams_per_block = V2_BLOCK / T;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v3_kernel<T>, V2_BLOCK, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (n + teams_per_block - 1) / teams_per_block;
    if (grid > need) grid = need;
    e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    decompress_v3_kernel<T><<<grid, V2_BLOCK, 0, st>>>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket);
    return cudaGetLastError();
}

inline cudaError_t v3_launch_decompress(int team, const u8* in, const u64* in_off, const u64* csz, const u64* out_off,
                                        const u64* usz, u8* out, u32* status, u32* bad, u32 n, u32* ticket,
                                        int sm_count, cudaStream_t st)
{
    switch (team) {
    case 16: return v3_launch_t<16>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    default: return v3_launch_t<8>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    }
}

}  // namespace lz4b200
