// kernels_v4.cuh -- generation-4 decode: v3's ideas with a warp-uniform instruction stream.
//
// The SASS of v2/v3 (profiles/sass) showed what the 280-340 instructions per 4-sequence iteration were: every
// __shfl_sync / __syncwarp with the per-team mask compiles to MATCH.ANY + REDUX + VOTEU + BRA.DIV (the compiler
// must prove the non-constant mask consistent), and every team-divergent branch re-executes its code once per
// direction.  v4 therefore
//   * runs ALL shuffles of the common path at warp level with the constant full mask -- every lane executes
//     them, a team that is not on the fast path this iteration just moves garbage between its own lanes;
//   * predicates the common path with one team-uniform flag instead of branching on it;
//   * keeps v3's memory discipline: each output byte is stored and later loaded by the same lane (position mod
//     T), the match source is loaded one sequence before it is stored, no barrier on the common path;
//   * handles everything else (long or overlapping matches, length bytes, first / last sequence, errors) in a
//     slow block entered by a warp-uniform vote, bracketed by full-mask barriers.
#pragma once
#include "common.cuh"
#include "kernels_v2.cuh"
#include "kernels_v3.cuh"

namespace lz4b200 {

template <u32 T>
__global__ void __launch_bounds__(V2_BLOCK, 7)     // <= 72 registers: 28 warps = 112 teams per SM, one wave for 16384 chunks
decompress_v4_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off, const u64* __restrict__ csz,
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

    u32 op_safe = 0, ip_safe = 0;                      // fast path allowed while op + L <= op_safe and ip + L <= ip_safe
    bool fastable = false;                             // busy, not a raw chunk, big enough for the fast path at all
    u32 rem = 0, roff = 0;                             // long match in progress: bytes still owed, its offset

    for (;;) {
        // ---- A: decode the current token of every team (warp level, no branches) ----
        // A match longer than 18 bytes (token nibble 15 + one length byte) is executed 18 bytes per iteration:
        // `rem` bytes are still owed at offset `roff`; while rem != 0 the (already prefetched) next token waits.
        const bool cont = rem != 0;
        const u32 Lt = tok >> 4, M = tok & ML_MASK;
        const u32 L = cont ? 0u : Lt;
        const u32 cpy = op + L;
        u32 sel_lo = xb[0], sel_hi = xb[0], sel_e = xb[0];
        #pragma unroll
        for (u32 k = 1; k < NX; k++) {
            if (Lt / T == k) sel_lo = xb[k];
            if ((Lt + 1) / T == k) sel_hi = xb[k];
            if ((Lt + 2) / T == k) sel_e = xb[k];
        }
        const u32 lo = __shfl_sync(FULL, sel_lo, tbase + (Lt & (T - 1)));
        const u32 hi = __shfl_sync(FULL, sel_hi, tbase + ((Lt + 1) & (T - 1)));
        const u32 e = __shfl_sync(FULL, sel_e, tbase + ((Lt + 2) & (T - 1)));     // length byte, if the nibble is 15
        const bool ext = M == ML_MASK;
        const u32 off = cont ? roff : (lo | (hi << 8));
        const u32 ml_total = cont ? rem : (ext ? 19u + e : M + MINMATCH);
        const u32 ml = ml_total < 18u ? ml_total : 18u;                           // this iteration's part
        const u32 nip = ip + Lt + 3 + (ext ? 1u : 0u);
        // fast: short literal run, non-overlapping match of <= 273 bytes, valid offset, far enough from both ends
        // of the block that (a) this is not the last sequence (lz4/lz4.c:1990), (b) the match ends 12 bytes
        // before the end and (c) the next window can be read unchecked
        const bool f = fastable && (cont || (tok < 0xF0u && (!ext || (Lt <= 13u && e != 255u)) &&
                                            cpy + ml_total <= op_safe && ip + Lt <= ip_safe && off <= cpy)) &&
                       off >= ml;

        // ---- B: store the previous part's match bytes (loaded one iteration ago), every lane its own positions ----
        {
            const u32 i0 = (j - pm_dst) & (T - 1);
            const u32 sl = tbase + ((j - pm_off) & (T - 1));
            u8* p = dst + (pm_dst + i0);
            asm volatile("" : "+l"(p));              // one address computation for the predicated stores below
            #pragma unroll
            for (u32 k = 0; k < NM; k++) {
                const u32 val = __shfl_sync(FULL, pvb[k], sl);
                if (i0 + k * T < pm_len) p[k * T] = (u8)val;
            }
            pm_len = 0;
        }
        // ---- C: literal run, window order -> position order ----
        {
            const u32 i0 = (j - op) & (T - 1);
            u8* p = dst + (op + i0);
            asm volatile("" : "+l"(p));
            const u32 lim = f ? L : 0u;
            #pragma unroll
            for (u32 k = 0; k < NX; k++) {
                const u32 val = __shfl_sync(FULL, xb[k], tbase + i0);
                if (i0 + k * T < lim) p[k * T] = (u8)val;
            }
        }
        // ---- D: this part's source (stored next iteration); at the first part also the next token + window ----
        if (f) {
            const u32 sbase = cpy - off;
            const u32 t0 = (j - sbase) & (T - 1);
            const u8* q = dst + (sbase + t0);
            asm volatile("" : "+l"(q));
            #pragma unroll
            for (u32 k = 0; k < NM; k++) pvb[k] = t0 + k * T < ml ? (u32)q[k * T] : 0u;
            pm_len = ml; pm_dst = cpy; pm_off = off;
            if (!cont) {
                const u8* s_ = src + nip;               // nip + 17 <= clen: unchecked reads
                tok = __ldg(s_);
                #pragma unroll
                for (u32 k = 0; k < NX; k++) xb[k] = (u32)__ldg(s_ + 1 + j + k * T);
                ip = nip; roff = off;
            }
            rem = ml_total - ml;
            op = cpy + ml;
        }

        // ---- E + F: everything else, behind one vote ----
        if (!__any_sync(FULL, (busy && !f) || !busy)) continue;
        u32 wl_len = 0, wl_ip = 0, wl_op = 0;          // warp-wide literal copy request
        u32 wm_len = 0, wm_op = 0, wm_off = 0;         // warp-wide match copy request
        u32 err = ST_OK;
        bool finished = false;
        // E: slow block -- one sequence through the general path for the teams that need it
        if (__any_sync(FULL, busy && !f)) {
            __syncwarp();                               // mapped stores above -> visible to the unmapped copies below
            if (busy && !f) {
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
            __syncwarp();
        }

        // ---- management, behind one vote: long copies by the whole warp, chunk completion, next chunk ----
        {
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
                busy = false; fastable = false;
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
                    fastable = clen != n && n >= 32 && clen >= 24;
                    op_safe = fastable ? n - MFLIMIT : 0;          // match end + MFLIMIT <= n (non-last sequence, lz4/lz4.c:1990,:2139)
                    ip_safe = fastable ? clen - 21 : 0;            // next window (17 bytes at ip + L + 3 or 4) stays inside the chunk
                    rem = 0;
                    V3_PREFETCH(0);
                }
            }
            if (__all_sync(FULL, exhausted)) break;
        }
    }
#undef V3_PREFETCH
}

template <u32 T>
inline cudaError_t v4_launch_t(const u8* in, const u64* in_off, const u64* csz, const u64* out_off, const u64* usz,
                               u8* out, u32* status, u32* bad, u32 n, u32* ticket, int sm_count, cudaStream_t st)
{
    const u32 teams_per_block = V2_BLOCK / T;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v4_kernel<T>, V2_BLOCK, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (n + teams_per_block - 1) / teams_per_block;
    if (grid > need) grid = need;
    e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    decompress_v4_kernel<T><<<grid, V2_BLOCK, 0, st>>>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket);
    return cudaGetLastError();
}

inline cudaError_t v4_launch_decompress(int team, const u8* in, const u64* in_off, const u64* csz, const u64* out_off,
                                        const u64* usz, u8* out, u32* status, u32* bad, u32 n, u32* ticket,
                                        int sm_count, cudaStream_t st)
{
    if (team == 16) return v4_launch_t<16>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    return v4_launch_t<8>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
}

}  // namespace lz4b200
