// kernels_v5.cuh -- generation-5 decode: v4 with the compressed stream read through a shared-memory ring.
//
// The v4 profile (profiles/r01_final_decompress_v4_*) has every iteration wait for a global load whose address
// is known only after the previous token is decoded: the next token and its 16-byte window.  At ~4000 sequences per
// 64 KiB chunk that round trip (L2 or HBM) is the per-chunk chain: ~3 ms per chunk however few chunks are in flight.
// But the compressed stream is read strictly forward, so its ADDRESSES are predictable: v5 keeps, per team, a ring
// of four 16*T-byte segments in shared memory, filled by cp.async (LDGSTS, no register staging) two segments ahead
// of the read position.  Token, offset bytes, length byte and the literals are then shared-memory reads; the only
// global loads left on the common path are the match source bytes, which v4 already issues one sequence early.
// Also: teams that ran out of chunks no longer drag their warp through the management block every iteration, and
// everything the common path does not touch (chunk base pointer, sizes, chunk id) lives in shared memory per team,
// and everything behind the common path's vote is a separate non-inlined function, so that the register allocation
// of the hot loop is not shaped by the slow code (v5's first build spilled the just-loaded match bytes inside the
// loop, which exposes the very load latency the pipelining hides).
#pragma once
#include "common.cuh"
#include "kernels_v2.cuh"
#include "kernels_v3.cuh"
#include "kernels_v4.cuh"

namespace lz4b200 {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, u32 src_bytes)
{
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

struct V5Team { const u8* src; u32 clen, n, chunk, raw, cap; };   // cold per-team state (aligned coordinates, see below)
struct V5Params {
    const u8* in; const u64* in_off; const u64* csz; const u64* out_off; const u64* usz; u8* out;
    u32* status; u32* bad_chunks; u32* ticket; u32 n_chunks;
    // 0: every chunk's end-of-block rules are checked against its own size (lzbench hands the exact size to the decoder);
    // else: against this capacity, as LZ4F_decompress does, which decodes every block into a buffer of the frame's block
    // size (lz4/lz4frame.c:1384-1830) -- a block that ends closer than 12 / 5 bytes behind its last match is then accepted,
    // and so it must be here (found by scripts/fuzz_gpu.py on a damaged frame).  The decoded size is checked either way.
    u32 cap_all;
    // optional work list (kernels_v7.cuh leaves the chunks that are mostly long copies to this kernel): tickets index
    // list[0 .. *list_count) instead of the chunks themselves
    const u32* list; const u32* list_count;
};
// loop-carried state of one lane, handed to the cold function through local memory
struct V5State { u8* dst; u32 ip, op, tok, pm_len, op_safe, ip_safe, rem; bool busy, exhausted, fastable; u32 lrem; };

// request stream segment g (16*T bytes at the 16-aligned base of the chunk + g*SEG) into its ring slot
template <u32 T>
__device__ __forceinline__ void v5_request(u8* ring, const u8* src, u32 clen, u32 g, u32 j)
{
    constexpr u32 SEG = 16 * T;
    const u8* a = src + (size_t)g * SEG + j * 16;
    cp_async16(ring + (g & 3u) * SEG + j * 16, a, a < src + clen ? 16u : 0u);
}

// Everything that is not the common path: one sequence through the general decoder for the teams that need it,
// warp-wide long copies, chunk completion, next chunk, ring restart.  Returns true when the warp has no work left.
template <u32 T>
__device__ __noinline__ bool v5_cold(V5State& S, const bool f, const V5Params& P, u8* const ring, V5Team& ts)
{
    constexpr u32 SEG = 16 * T, RING = 4 * SEG;
    const u32 lane = lane_id();
    const u32 j = lane & (T - 1);
    const u32 tbase = lane & ~(T - 1);
    const unsigned tmask = T == 32 ? FULL : (((1u << (T & 31)) - 1u) << tbase);
    bool busy = S.busy, exhausted = S.exhausted, fastable = S.fastable;
    u8* dst = S.dst;
    u32 ip = S.ip, op = S.op;
    u32 wl_len = 0, wl_ip = 0, wl_op = 0;          // warp-wide literal copy request
    u32 wm_len = 0, wm_op = 0, wm_off = 0;         // warp-wide match copy request
    u32 err = ST_OK;
    bool finished = false;
    bool restart = false;                          // the read position moved outside the common path: ring upkeep, reload tok
    bool fresh = false;                            // new chunk
    const u8* src = ts.src;
    u32 clen = ts.clen, n = ts.n, capv = ts.cap;
    // stream bytes below `landed` are in the ring (segments seg0 and seg0 + 1 are complete, seg0 + 2 may be in flight)
    const u32 seg0 = ip / SEG;
    const u32 landed = fastable ? (seg0 + 2) * SEG : 0u;
#define V5_RD(POS) ((POS) < landed ? (u32)ring[(POS) & (RING - 1)] : (u32)__ldg(src + (POS)))
    // E: slow block -- one sequence through the general path for the teams that need it
    if (__any_sync(FULL, busy && !f)) {
        __syncwarp();                               // mapped stores of the common path -> visible to the unmapped copies below
        if (busy && !f) {
            if (ts.raw) {                              // raw chunk (_lzbench/lzbench.cpp:311-315)
                wl_len = n; wl_ip = ip; finished = true;
            } else if (n == 0 || clen == ip) {         // lz4/lz4.c:1776-1782 (only possible at the first token)
                err = (n == 0 && clen - ip == 1 && __ldg(src + ip) == 0) ? ST_OK : ST_EMPTY; finished = true;
            } else if (ip >= clen) {
                err = ST_TOKEN_OOB; finished = true;
            } else {
                const u32 token = S.tok;
                u64 len = S.lrem;                       // mid literal run (common path, 16 bytes per iteration): ip is at the next literal
                if (len == 0) { ip++; len = token >> 4; }
                if (S.lrem == 0 && len == RUN_MASK) {
                    if (ip + RUN_MASK >= clen) err = ST_LITLEN_OOB;
                    else {
                        u32 s;
                        do {
                            s = V5_RD(ip); ip++; len += s;
                            if (s == 255 && ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                        } while (s == 255);
                    }
                }
                const u64 cpy = op + len;
                if (err) finished = true;
                else if (cpy + MFLIMIT > capv || ip + len + (2 + 1 + LASTLITERALS) > clen) {  // last sequence
                    if (ip + len != clen || cpy > n) err = ST_LAST_LITERALS;
                    else {
                        const u32 LL = (u32)len;
                        if (LL >= V2_LONG) { wl_len = LL; wl_ip = ip; wl_op = op; }
                        else for (u32 i = j; i < LL; i += T) dst[op + i] = (u8)V5_RD(ip + i);
                        op = (u32)cpy;
                        if (op != n) err = ST_SIZE;
                    }
                    finished = true;
                } else {
                    const u32 LL = (u32)len;
                    if (LL >= V2_LONG) { wl_len = LL; wl_ip = ip; wl_op = op; }
                    else for (u32 i = j; i < LL; i += T) dst[op + i] = (u8)V5_RD(ip + i);
                    const u32 off = V5_RD(ip + LL) | (V5_RD(ip + LL + 1) << 8);
                    ip += LL + 2; op = (u32)cpy;
                    u64 ml64 = token & ML_MASK;
                    if (ml64 == ML_MASK) {
                        u32 s;
                        do {
                            s = V5_RD(ip); ip++; ml64 += s;
                            if (ip + LASTLITERALS - 1 >= clen) { err = ST_MATCHLEN_OOB; break; }
                        } while (s == 255);
                    }
                    ml64 += MINMATCH;
                    if (!err && (off == 0 || off > op)) err = ST_OFFSET;
                    if (!err && (op + ml64 + LASTLITERALS > capv || op + ml64 > n)) err = ST_MATCH_END;
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
                        restart = true;
                    }
                }
            }
        }
        __syncwarp();
    }

    // ---- management: long copies by the whole warp, chunk completion, next chunk ----
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
            P.status[ts.chunk] = err;
            if (err && P.bad_chunks) atomicAdd(P.bad_chunks, 1u);
        }
        busy = false; fastable = false;
    }
    if (!busy && !exhausted) {
        u32 c = 0;
        if (j == 0) c = atomicAdd(P.ticket, 1u);
        c = __shfl_sync(tmask, c, tbase);
        if (c >= (P.list ? *P.list_count : P.n_chunks)) exhausted = true;
        else {
            if (P.list) c = P.list[c];
            src = P.in + P.in_off[c]; clen = (u32)P.csz[c];
            dst = P.out + P.out_off[c]; n = (u32)P.usz[c];
            op = 0; busy = true; S.pm_len = 0; S.rem = 0; S.lrem = 0;
            const bool raw = clen == n;
            fastable = !raw && n >= 32 && clen >= 24;
            capv = P.cap_all > n ? P.cap_all : n;
            S.op_safe = fastable ? capv - MFLIMIT : 0;     // match end + MFLIMIT <= capacity (non-last sequence, lz4/lz4.c:1990,:2139)
            ip = (u32)(reinterpret_cast<uintptr_t>(src) & 15u);
            src -= ip; clen += ip;                         // aligned coordinates
            S.ip_safe = fastable ? clen - 21 : 0;          // next token + 17 bytes (at ip + L + 3 or 4) stay inside the chunk
            __syncwarp(tmask);                             // the team's earlier reads of ts are done
            if (j == 0) { ts.src = src; ts.clen = clen; ts.n = n; ts.chunk = c; ts.raw = raw; ts.cap = capv; }
            restart = true; fresh = true;
        }
    }
    S.busy = busy; S.exhausted = exhausted; S.fastable = fastable; S.dst = dst; S.ip = ip; S.op = op;
    if (__all_sync(FULL, exhausted)) return true;
    // Ring upkeep.  A new chunk starts the ring (three segments from the one holding ip, all landed before anyone
    // reads).  After a slow sequence the read position has moved by an arbitrary amount: request what is missing up
    // to two segments ahead of it; only a jump past the landed part has to wait for data.
    const u32 seg1 = ip / SEG;
    u32 adv = 0;                                           // 0 nothing to do, 1 one segment further, 2 wait for everything
    if (restart && fastable) {
        if (fresh || seg1 > seg0 + 2) {
            v5_request<T>(ring, src, clen, seg1, j); v5_request<T>(ring, src, clen, seg1 + 1, j);
            v5_request<T>(ring, src, clen, seg1 + 2, j);
            adv = 2;
        } else if (seg1 != seg0) {
            for (u32 g = seg0 + 3; g <= seg1 + 2; g++) v5_request<T>(ring, src, clen, g, j);
            adv = seg1 == seg0 + 1 ? 1 : 2;
        }
    }
    const bool any2 = __any_sync(FULL, adv == 2);
    if (any2 || __any_sync(FULL, adv == 1) || __any_sync(FULL, fresh)) {
        cp_async_commit();
        if (any2) cp_async_wait<0>(); else cp_async_wait<1>();
        __syncwarp();                                      // ring and ts visible to the whole team
    }
    if (restart) S.lrem = 0;
    if (restart) S.tok = fastable ? (u32)ring[ip & (RING - 1)] : (ip < clen ? (u32)__ldg(src + ip) : 0u);
#undef V5_RD
    return false;
}

template <u32 T>
__global__ void __launch_bounds__(V2_BLOCK, 7)     // <= 72 registers: 28 warps = 112 teams per SM, one wave for 16384 chunks
decompress_v5_kernel(const __grid_constant__ V5Params P)
{
    const u32 lane = lane_id();
    u32 j = lane & (T - 1);
    u32 tbase = lane & ~(T - 1);
    asm volatile("" : "+r"(j), "+r"(tbase));           // keep them in registers: ptxas otherwise re-reads SR_TID.X three times per iteration
    constexpr u32 NX = (16 + T - 1) / T;               // literal bytes per lane (runs of <= 14 on the common path)
    constexpr u32 NM = (18 + T - 1) / T;               // bytes of a short match (<= 18) per lane
    constexpr u32 SEG = 16 * T, RING = 4 * SEG;        // one cp.async of 16 bytes per lane fills a segment
    __shared__ __align__(16) u8 ring_all[V2_BLOCK / T][RING];
    __shared__ V5Team team_all[V2_BLOCK / T];
    __shared__ volatile u32 v5_sink;
    u8* const ring = ring_all[threadIdx.x / T];
    V5Team& ts = team_all[threadIdx.x / T];

    // ts.src is the chunk start rounded DOWN to 16 bytes, and ip / ts.clen / ip_safe count from there (the first
    // token is at ip = chunk start & 15): stream byte p is ring[p % RING], segment g is the 16-aligned ts.src + g*SEG
    bool busy = false, exhausted = false;
    bool first = true;                                 // the first pass goes through the cold function to fetch the first chunks
    u8* dst = nullptr;
    u32 ip = 0, op = 0;
    u32 tok = 0;                                       // token at stream position ip
    u32 pm_len = 0;                                    // pending match part: pvb[] loaded, stored one iteration later at op - pm_len, offset roff
    u32 pvb[NM];
    #pragma unroll
    for (u32 k = 0; k < NM; k++) pvb[k] = 0;
    u32 op_safe = 0, ip_safe = 0;                      // fast path allowed while op + L <= op_safe and ip + L <= ip_safe
    bool fastable = false;                             // busy, not a raw chunk, big enough for the fast path at all
    u32 rem = 0, roff = 0;                             // long match in progress: bytes still owed; offset of the last fast match
    u32 lrem = 0;                                      // long literal run in progress: literals still owed before the token's match; ip is at the next one

    for (;;) {
        // ---- A: decode the current token of every team (warp level, no branches) ----
        // A match longer than 18 bytes (token nibble 15 + one length byte) is executed 18 bytes per iteration:
        // `rem` bytes are still owed at offset `roff`; while rem != 0 the (already read) next token waits.
        // A literal run of 15..269 bytes (token nibble 15 + one length byte, 3 % of the sequences of the bench corpus
        // and almost all of what used to leave the common path) is executed 16 bytes per iteration in the same way:
        // `lrem` literals are still owed and ip has moved to the next of them; the token (and its match) waits.
        const bool cont = rem != 0, lcont = lrem != 0;
        const u32 Lt = tok >> 4, M = tok & ML_MASK;
        const u32 e1 = ring[(ip + 1) & (RING - 1)];      // literal length byte, if the nibble is 15
        const bool lext = Lt == RUN_MASK;
        const u32 LL = lcont ? lrem : (lext ? 15u + e1 : Lt);                    // literals owed before the match
        const u32 lpos = ip + (lcont ? 0u : (lext ? 2u : 1u));                    // stream position of the next literal
        const bool lonly = !cont && LL > 16u;                                      // this iteration: 16 literals, nothing else
        const u32 L = cont ? 0u : (lonly ? 16u : LL);
        const u32 cpy = op + L;
        const u32 mp = lpos + LL;                        // stream position of the match offset
        const u32 lo = ring[mp & (RING - 1)];
        const u32 hi = ring[(mp + 1) & (RING - 1)];
        const u32 e = ring[(mp + 2) & (RING - 1)];       // match length byte, if the nibble is 15
        const bool ext = M == ML_MASK;
        const u32 off = cont ? roff : (lo | (hi << 8));
        const u32 ml_total = cont ? rem : (ext ? 19u + e : M + MINMATCH);
        const u32 ml = lonly ? 0u : (ml_total < 18u ? ml_total : 18u);          // this iteration's part
        const u32 nip = mp + 2 + (ext ? 1u : 0u);
        // fast: literal run of at most 269 bytes, non-overlapping match of <= 273 bytes, valid offset, far enough from
        // both ends of the block that (a) this is not the last sequence (lz4/lz4.c:1990), (b) the match ends 12 bytes
        // before the end and (c) the next token and its 17 bytes are inside the chunk
        // (bitwise on purpose: with && / || ptxas builds a divergent branch tree here and parks the wait for the
        //  in-flight match loads on its first branch, 20 % of all stall samples)
        const u32 c_cont = cont, c_lcont = lcont, c_lonly = lonly;
        const u32 c_hdr = c_cont | c_lcont | ((u32)(!lext | (e1 != 255u)) & (u32)(mp <= ip_safe));
        const u32 c_lit = (u32)(cpy + 16u <= op_safe);
        const u32 c_mat = (c_cont | ((u32)(!ext | (e != 255u)) & (u32)(cpy + ml_total <= op_safe) & (u32)(off <= cpy))) &
                          (u32)(off >= ml);
        const bool f = ((u32)fastable & c_hdr & (c_lonly ? c_lit : c_mat)) != 0;

        // ---- B: store the previous part's match bytes (loaded one iteration ago), every lane its own positions ----
        {
            const u32 pm_dst = op - pm_len;
            const u32 i0 = (j - pm_dst) & (T - 1);
            const u32 sl = tbase + ((j - roff) & (T - 1));
            u8* p = dst + (pm_dst + i0);
            asm volatile("" : "+l"(p));              // one address computation for the predicated stores below
            #pragma unroll
            for (u32 k = 0; k < NM; k++) {
                const u32 val = __shfl_sync(FULL, pvb[k], sl);
                if (i0 + k * T < pm_len) p[k * T] = (u8)val;
            }
            pm_len = 0;
        }
        // ---- C: literal run, straight from the ring ----
        {
            const u32 i0 = (j - op) & (T - 1);
            u8* p = dst + (op + i0);
            asm volatile("" : "+l"(p));
            const u32 lim = f ? L : 0u;
            #pragma unroll
            for (u32 k = 0; k < NX; k++) {
                const u32 val = ring[(lpos + i0 + k * T) & (RING - 1)];
                if (i0 + k * T < lim) p[k * T] = (u8)val;
            }
        }
        // ---- R: keep the ring two segments ahead of where the next iteration reads.  Placed here because no global
        //      load is outstanding at this point (B consumed last iteration's), so the barrier drains nothing.
        {
            const u32 ip2 = lonly ? lpos + 16u : nip;                       // where ip goes if this iteration is fast
            const bool need = f && !cont && (ip2 / SEG != ip / SEG);        // the read position enters a new segment
            if (__any_sync(FULL, need)) {
                if (need) v5_request<T>(ring, ts.src, ts.clen, ip2 / SEG + 2, j);
                cp_async_commit();
                cp_async_wait<1>();                     // everything but the segment just requested has landed
                __syncwarp();
            }
        }
        // ---- D: this part's source (stored next iteration); at the first part also the next token ----
        if (f) {
            const u32 sbase = cpy - off;
            const u32 t0 = (j - sbase) & (T - 1);
            const u8* q = dst + (sbase + t0);
            asm volatile("" : "+l"(q));
            #pragma unroll
            for (u32 k = 0; k < NM; k++) pvb[k] = t0 + k * T < ml ? (u32)q[k * T] : 0u;
            pm_len = ml;
            if (lonly) { ip = lpos + 16u; lrem = LL - 16u; }
            else if (!cont) {
                tok = ring[nip & (RING - 1)];
                ip = nip; roff = off; lrem = 0;
            }
            if (!lonly) rem = ml_total - ml;
            op = cpy + ml;
        }

        // ---- everything else, behind one vote and one call ----
        // a team that is not busy here has run out of chunks (the cold function hands out the next chunk or marks the team
        // exhausted before it returns), so only busy teams can ask for it
        if (!__any_sync(FULL, busy && !f) && !first) continue;
        first = false;
        V5State S{ dst, ip, op, tok, pm_len, op_safe, ip_safe, rem, busy, exhausted, fastable, lrem };
        const bool done = v5_cold<T>(S, f, P, ring, ts);
        dst = S.dst; ip = S.ip; op = S.op; tok = S.tok; pm_len = S.pm_len; op_safe = S.op_safe; ip_safe = S.ip_safe;
        rem = S.rem; busy = S.busy; exhausted = S.exhausted; fastable = S.fastable; lrem = S.lrem;
        if (done) break;
        {   // Consume every restored value HERE.  They come back through local memory, and ptxas gives those loads
            // the same scoreboard as the match-source loads of the loop: without a consumer on this side of the
            // back edge, the first instruction of the loop waits for that scoreboard in EVERY iteration, i.e. for
            // the match loads it is supposed to overlap (seen as 25 % of all stall samples on the loop head).
            const u32 sink = ip ^ op ^ tok ^ pm_len ^ op_safe ^ ip_safe ^ rem ^ lrem ^ (u32)reinterpret_cast<uintptr_t>(dst) ^
                             (u32)(reinterpret_cast<uintptr_t>(dst) >> 32) ^ (u32)busy ^ (u32)exhausted ^ (u32)fastable;
            if (sink == 0x9E3779B9u) v5_sink = sink;            // harmless; a real use of the values that ptxas cannot drop
        }
    }
}

template <u32 T>
inline cudaError_t v5_launch_t(const u8* in, const u64* in_off, const u64* csz, const u64* out_off, const u64* usz,
                               u8* out, u32* status, u32* bad, u32 n, u32* ticket, int sm_count, cudaStream_t st, u32 cap_all,
                               const u32* list, const u32* list_count, bool ticket_cleared = false)
{
    const u32 teams_per_block = V2_BLOCK / T;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v5_kernel<T>, V2_BLOCK, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (n + teams_per_block - 1) / teams_per_block;
    if (grid > need) grid = need;
    if (!ticket_cleared) {
        e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
        if (e != cudaSuccess) return e;
    }
    const V5Params P{ in, in_off, csz, out_off, usz, out, status, bad, ticket, n, cap_all, list, list_count };
    decompress_v5_kernel<T><<<grid, V2_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

inline cudaError_t v5_launch_decompress(int team, const u8* in, const u64* in_off, const u64* csz, const u64* out_off,
                                        const u64* usz, u8* out, u32* status, u32* bad, u32 n, u32* ticket,
                                        int sm_count, cudaStream_t st, u32 cap_all = 0,
                                        const u32* list = nullptr, const u32* list_count = nullptr, bool ticket_cleared = false)
{
    if (team == 16) return v5_launch_t<16>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st, cap_all, list, list_count, ticket_cleared);
    if (team == 4) return v5_launch_t<4>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st, cap_all, list, list_count, ticket_cleared);
    return v5_launch_t<8>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st, cap_all, list, list_count, ticket_cleared);
}

}  // namespace lz4b200
