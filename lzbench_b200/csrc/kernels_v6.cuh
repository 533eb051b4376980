// kernels_v6.cuh -- generation-6 decode: v5 with the match copy pipelined several iterations deep through shared memory.
//
// What is left of a common-path iteration in v5 (gpurun_out/prof_r01_v5t8c) is one L2/HBM round trip: the match source
// bytes are loaded into registers at the end of iteration i and stored at the start of iteration i+1, and nothing of
// iteration i+1 can issue past that store (in-order issue; and all register loads of a warp share one scoreboard, so
// a second register slot does not help -- tried: the first use of the older slot waits for the younger loads too).
// v6 moves the pending match bytes out of the register file: the aligned 16-byte blocks that hold the source are
// fetched with cp.async into a per-team staging slot, cp.async groups count completions in order, and the bytes are
// stored DEPTH iterations later (cp.async.wait_group DEPTH-1 at the top of every iteration).  Lanes no longer read
// only what they wrote themselves (the blocks are fetched by two lanes), so the iteration carries two warp barriers:
// one after the wait (staged bytes -> all lanes), one between the stores and the next fetch (stored bytes -> fetching
// lanes).  With no register load in flight a barrier has nothing to drain, which is what made them expensive in v3.
// A match whose source overlaps the destination of a pending slot, and any team that leaves the common path, first
// flushes the pending slots (wait for everything, store in order) -- always legal, it only waits.
#pragma once
#include "common.cuh"
#include "kernels_v5.cuh"

namespace lz4b200 {

constexpr u32 V6_PART = 16;                          // bytes of a match executed per iteration (17/18-byte matches take two)
constexpr u32 V6_SLOT = 32;                          // staging bytes per pending part: two aligned 16-byte blocks

// pending part metadata in one word: (len << 27) | ((source address & 15) << 23) | (destination mod 2^23);
// the destination is rebuilt from the current output position, which is never 2^23 ahead
__device__ __forceinline__ u32 v6_len(u32 meta) { return meta >> 27; }
__device__ __forceinline__ u32 v6_dst(u32 meta, u32 op) { return op - ((op - meta) & 0x7FFFFFu); }

// A landed slot goes out in two steps so that the shared-memory reads of everything an iteration stores (the slot and
// the literals) are issued together, ahead of the first store that needs them: with 3-7 warps per scheduler an LDS
// directly in front of its ST is a 30-cycle stall each time (short scoreboard: 60 % of the stall samples of the first
// version).  Lane j owns the output positions == j (mod T).
template <u32 T>
struct V6Out { u32 val[(V6_PART + T - 1) / T]; u32 len, i0; u8* p; };

template <u32 T>
__device__ __forceinline__ void v6_load(V6Out<T>& o, u32 meta, const u8* slot, u8* dst, u32 op, u32 j)
{
    constexpr u32 NM = (V6_PART + T - 1) / T;
    const u32 sdst = v6_dst(meta, op), soff = (meta >> 23) & 15u;
    o.len = v6_len(meta);
    o.i0 = (j - sdst) & (T - 1);
    o.p = dst + (sdst + o.i0);
    #pragma unroll
    for (u32 k = 0; k < NM; k++) o.val[k] = slot[soff + o.i0 + k * T];
}

template <u32 T>
__device__ __forceinline__ void v6_put(const V6Out<T>& o)
{
    constexpr u32 NM = (V6_PART + T - 1) / T;
    u8* p = o.p;
    asm volatile("" : "+l"(p));                      // one address computation for the predicated stores below
    #pragma unroll
    for (u32 k = 0; k < NM; k++)
        if (o.i0 + k * T < o.len) p[k * T] = (u8)o.val[k];
}

template <u32 T>
__device__ __forceinline__ void v6_store(u32 meta, const u8* slot, u8* dst, u32 op, u32 j)
{
    V6Out<T> o;
    v6_load<T>(o, meta, slot, dst, op, j);
    v6_put<T>(o);
}

template <u32 T, u32 DEPTH>
__global__ void __launch_bounds__(V2_BLOCK, 7)
decompress_v6_kernel(const __grid_constant__ V5Params P)
{
    const u32 lane = lane_id();
    const u32 j = lane & (T - 1);
    constexpr u32 NX = (16 + T - 1) / T;
    constexpr u32 SEG = 16 * T, RING = 4 * SEG;
    __shared__ __align__(16) u8 ring_all[V2_BLOCK / T][RING];
    __shared__ __align__(16) u8 stage_all[V2_BLOCK / T][DEPTH * V6_SLOT];
    __shared__ V5Team team_all[V2_BLOCK / T];
    __shared__ volatile u32 v6_sink;
    u8* const ring = ring_all[threadIdx.x / T];
    u8* const stage = stage_all[threadIdx.x / T];
    V5Team& ts = team_all[threadIdx.x / T];

    bool busy = false, exhausted = false, fastable = false;
    u8* dst = nullptr;
    u32 ip = 0, op = 0, tok = 0, op_safe = 0, ip_safe = 0, rem = 0, roff = 0, lrem = 0;
    u32 meta[DEPTH];                                 // meta[0]: part fetched DEPTH iterations ago (stored now) ... meta[DEPTH-1]: last iteration's
    #pragma unroll
    for (u32 k = 0; k < DEPTH; k++) meta[k] = 0;
    u32 si = 0;                                      // staging slot of meta[0]; meta[k] is in slot (si + k) % DEPTH

    for (;;) {
        cp_async_wait<DEPTH - 1>();                  // the group committed DEPTH iterations ago has landed
        __syncwarp();                                // ... and is visible to every lane of the team
        // ---- A: decode the current token of every team (warp level, no branches); see kernels_v4.cuh / v5 ----
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
        const u32 ml = lonly ? 0u : (ml_total < V6_PART ? ml_total : V6_PART);
        const u32 nip = mp + 2 + (ext ? 1u : 0u);
        const u32 c_cont = cont, c_lcont = lcont, c_lonly = lonly;           // bitwise on purpose, see kernels_v5.cuh
        const u32 c_hdr = c_cont | c_lcont | ((u32)(!lext | (e1 != 255u)) & (u32)(mp <= ip_safe));
        const u32 c_lit = (u32)(cpy + 16u <= op_safe);
        const u32 c_mat = (c_cont | ((u32)(!ext | (e != 255u)) & (u32)(cpy + ml_total <= op_safe) & (u32)(off <= cpy))) &
                          (u32)(off >= ml);
        const bool f = ((u32)fastable & c_hdr & (c_lonly ? c_lit : c_mat)) != 0;
        const u32 sbase = cpy - off;
        // pending destinations are contiguous and end at op: the match may read nothing at or after the oldest one
        u32 pend_lo = 0xFFFFFFFFu;
        #pragma unroll
        for (u32 k = DEPTH - 1; k >= 1; k--) if (v6_len(meta[k])) pend_lo = v6_dst(meta[k], op);
        const bool flush = busy && (!f || (ml != 0 && sbase + ml > pend_lo));

        // ---- B + C: the part fetched DEPTH iterations ago and this sequence's literals: all reads, then all stores;
        //      the younger parts only if somebody has to ----
        V6Out<T> ob;
        v6_load<T>(ob, meta[0], stage + si * V6_SLOT, dst, op, j);
        u32 lit[NX];
        const u32 li0 = (j - op) & (T - 1);
        #pragma unroll
        for (u32 k = 0; k < NX; k++) lit[k] = ring[(lpos + li0 + k * T) & (RING - 1)];
        v6_put<T>(ob);
        if (__any_sync(FULL, flush)) {
            cp_async_wait<0>();
            __syncwarp();
            #pragma unroll
            for (u32 k = 1; k < DEPTH; k++) {
                const u32 s = si + k < DEPTH ? si + k : si + k - DEPTH;
                v6_store<T>(meta[k], stage + s * V6_SLOT, dst, op, j);
                meta[k] = 0;
            }
        }
        {
            u8* p = dst + (op + li0);
            asm volatile("" : "+l"(p));
            const u32 lim = f ? L : 0u;
            #pragma unroll
            for (u32 k = 0; k < NX; k++)
                if (li0 + k * T < lim) p[k * T] = (u8)lit[k];
        }
        __syncwarp();                                // stores of B and C (all lanes) -> visible to the fetches below
        // ---- D: fetch this part's source blocks into the slot just emptied, and the next ring segment if due ----
        #pragma unroll
        for (u32 k = 0; k + 1 < DEPTH; k++) meta[k] = meta[k + 1];
        meta[DEPTH - 1] = 0;
        if (f) {
            const u32 ip2 = lonly ? lpos + 16u : nip;
            if (!lonly) {
                const u8* q = dst + sbase;
                const u32 soff = (u32)(reinterpret_cast<uintptr_t>(q) & 15u);
                if (j == 0 || (j == 1 && soff + ml > 16u)) cp_async16(stage + si * V6_SLOT + j * 16, q - soff + j * 16, 16u);
                meta[DEPTH - 1] = (ml << 27) | (soff << 23) | (cpy & 0x7FFFFFu);
            }
            if (!cont) {
                if (ip2 / SEG != ip / SEG) v5_request<T>(ring, ts.src, ts.clen, ip2 / SEG + 2, j);
                if (lonly) lrem = LL - 16u;
                else { tok = ring[nip & (RING - 1)]; roff = off; lrem = 0; }
                ip = ip2;
            }
            if (!lonly) rem = ml_total - ml;
            op = cpy + ml;
        }
        cp_async_commit();
        si = si + 1 == DEPTH ? 0u : si + 1;

        // ---- everything else, behind one vote and one call (slow teams have nothing pending: all flushed) ----
        if (!__any_sync(FULL, (busy && !f) || (!busy && !exhausted))) continue;
        V5State S{ dst, ip, op, tok, 0u, op_safe, ip_safe, rem, busy, exhausted, fastable, lrem };
        const bool done = v5_cold<T>(S, f, P, ring, ts);
        dst = S.dst; ip = S.ip; op = S.op; tok = S.tok; op_safe = S.op_safe; ip_safe = S.ip_safe;
        rem = S.rem; busy = S.busy; exhausted = S.exhausted; fastable = S.fastable; lrem = S.lrem;
        if (done) break;
        {   // consume the restored values on this side of the back edge (see kernels_v5.cuh: scoreboard sharing)
            const u32 sink = ip ^ op ^ tok ^ op_safe ^ ip_safe ^ rem ^ lrem ^ (u32)reinterpret_cast<uintptr_t>(dst) ^
                             (u32)(reinterpret_cast<uintptr_t>(dst) >> 32) ^ (u32)busy ^ (u32)exhausted ^ (u32)fastable;
            if (sink == 0x9E3779B9u) v6_sink = sink;
        }
    }
    cp_async_wait<0>();
}

template <u32 T, u32 DEPTH>
inline cudaError_t v6_launch_t(const u8* in, const u64* in_off, const u64* csz, const u64* out_off, const u64* usz,
                               u8* out, u32* status, u32* bad, u32 n, u32* ticket, int sm_count, cudaStream_t st)
{
    const u32 teams_per_block = V2_BLOCK / T;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decompress_v6_kernel<T, DEPTH>, V2_BLOCK, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (n + teams_per_block - 1) / teams_per_block;
    if (grid > need) grid = need;
    e = cudaMemsetAsync(ticket, 0, sizeof(u32), st);
    if (e != cudaSuccess) return e;
    const V5Params P{ in, in_off, csz, out_off, usz, out, status, bad, ticket, n, 0u, nullptr, nullptr };
    decompress_v6_kernel<T, DEPTH><<<grid, V2_BLOCK, 0, st>>>(P);
    return cudaGetLastError();
}

inline cudaError_t v6_launch_decompress(int team, const u8* in, const u64* in_off, const u64* csz, const u64* out_off,
                                        const u64* usz, u8* out, u32* status, u32* bad, u32 n, u32* ticket,
                                        int sm_count, cudaStream_t st, int d = 2)
{
    if (team == 4) {
        // depth <= 3 on 4-lane teams: a ring segment (64 bytes here) is requested two segments ahead and must have landed
        // >= 3 iterations later, which wait_group 3 would not guarantee
        if (d >= 3) return v6_launch_t<4, 3>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
        return v6_launch_t<4, 2>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    }
    if (team == 16) {
        if (d == 3) return v6_launch_t<16, 3>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
        return v6_launch_t<16, 2>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    }
    if (d == 3) return v6_launch_t<8, 3>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    if (d == 4) return v6_launch_t<8, 4>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
    return v6_launch_t<8, 2>(in, in_off, csz, out_off, usz, out, status, bad, n, ticket, sm_count, st);
}

}  // namespace lz4b200
