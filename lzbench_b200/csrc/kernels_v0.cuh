// kernels_v0.cuh -- generation-0 kernels: one warp per chunk, any chunk size, data read and
// written directly in global memory.  They are the always-correct baseline the tiled
// kernels (kernels_v1.cuh) are measured against, and the path for chunk shapes the tiled
// kernels do not take.
//
// Compression follows the shape of the reference's greedy parser
// (lz4/lz4.c:851-1240: single probe, 13-bit hash of 4 bytes, u16 position table,
// backward catch-up :1019, forward count :1092, end rules :883-884,:921) but examines 32
// positions per step and picks the earliest hit by ballot.  Decompression follows the
// sequence grammar and the end-of-block checks of LZ4_decompress_safe
// (lz4/lz4.c:1929-2151) with every read and write bounds-checked.
#pragma once
#include "common.cuh"

namespace lz4b200 {

constexpr int V0_WARPS = 4;                      // warps (= chunks) per CTA
constexpr u32 V0_TABLE = 8192;                   // u16 entries per warp (lz4/lz4.h:579: 16 KB)
constexpr u32 V0_SEG = 65536;                    // match window: positions fit the u16 table

// Write a length tail (lz4 255-run encoding) for value r = len - 15; returns bytes written.
__device__ __forceinline__ u32 v0_put_len_ext(u8* p, u32 r, u32 lane)
{
    u32 nff = r / 255;
    for (u32 i = lane; i < nff; i += 32) p[i] = 255;
    if (lane == 0) p[nff] = (u8)(r - nff * 255);
    return nff + 1;
}

__device__ __forceinline__ u32 v0_emit_sequence(u8* dst, u32 op, const u8* lit_ptr, u32 lit,
                                                u32 off, u32 mlen, u32 lane)
{
    u32 ml = mlen - MINMATCH;
    if (lane == 0) dst[op] = (u8)(((lit < RUN_MASK ? lit : RUN_MASK) << 4) | (ml < ML_MASK ? ml : ML_MASK));
    op++;
    if (lit >= RUN_MASK) op += v0_put_len_ext(dst + op, lit - RUN_MASK, lane);
    for (u32 i = lane; i < lit; i += 32) dst[op + i] = __ldg(lit_ptr + i);
    op += lit;
    if (lane == 0) { dst[op] = (u8)off; dst[op + 1] = (u8)(off >> 8); }
    op += 2;
    if (ml >= ML_MASK) op += v0_put_len_ext(dst + op, ml - ML_MASK, lane);
    return op;
}

__global__ void __launch_bounds__(V0_WARPS * 32)
compress_v0_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off,
                   const u64* __restrict__ usz, u8* __restrict__ stage, u64 slot_stride,
                   u64* __restrict__ csz, u32 n_chunks)
{
    extern __shared__ __align__(16) u16 v0_tables[];
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 chunk = blockIdx.x * V0_WARPS + warp;
    if (chunk >= n_chunks) return;                    // no block-wide barrier below
    u16* tab = v0_tables + warp * V0_TABLE;
    const u8* src = in + in_off[chunk];
    const u32 n = (u32)usz[chunk];
    u8* dst = stage + (u64)chunk * slot_stride;
    u32 op = 0, anchor = 0;

    if (n >= MIN_LENGTH) {                            // lz4/lz4.c:921
        const u32 mfl = n - MFLIMIT;                  // last position a match may start at (:883,:969)
        const u32 mlimit = n - LASTLITERALS;          // matches end at or before this (:884)
        for (u32 seg = 0; seg < n; seg += V0_SEG) {
            const u32 seg_end = seg + V0_SEG < n ? seg + V0_SEG : n;
            if (anchor >= seg_end) continue;
            uint4* t4 = reinterpret_cast<uint4*>(tab);
            for (u32 i = lane; i < V0_TABLE * 2 / 16; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            u32 ip = anchor > seg ? anchor : seg;
            while (ip < seg_end && ip <= mfl) {
                const u32 pos = ip + lane;
                const bool valid = pos <= mfl && pos < seg_end;
                u32 v = 0, h = 0x10000u + lane, cand = 0;
                if (valid) { v = ldg_u32_unaligned(src + pos); h = hash13(v); cand = seg + tab[h]; }
                const unsigned peers = __match_any_sync(FULL, h);   // lanes of this step with my hash
                const unsigned lower = peers & ((1u << lane) - 1u);
                bool ok = false;
                if (valid) {
                    if (lower) cand = ip + (31 - __clz(lower));      // nearest earlier lane
                    ok = cand < pos && ldg_u32_unaligned(src + cand) == v;
                }
                const unsigned hits = __ballot_sync(FULL, ok);
                // Insert only positions up to the match start (all 32 on a miss): a position that
                // will be probed again later must not find itself in the table.
                const int first = hits ? __ffs(hits) - 1 : 31;
                const unsigned upto = peers & (0xFFFFFFFFu >> (31 - first));
                __syncwarp();
                if (valid && (int)lane <= first && (upto >> lane) == 1u) tab[h] = (u16)(pos - seg);   // last writer wins
                __syncwarp();
                if (!hits) { ip += 32; continue; }
                u32 mpos = ip + first;
                u32 mcand = __shfl_sync(FULL, cand, first);
                while (mpos > anchor && mcand > seg && __ldg(src + mpos - 1) == __ldg(src + mcand - 1)) { mpos--; mcand--; }
                u32 mlen = MINMATCH;
                for (;;) {
                    const u32 a = mpos + mlen + lane;
                    const bool eq = a < mlimit && __ldg(src + a) == __ldg(src + mcand + mlen + lane);
                    const unsigned m = __ballot_sync(FULL, eq);
                    if (m == FULL) { mlen += 32; continue; }
                    mlen += __ffs(~m) - 1;
                    break;
                }
                op = v0_emit_sequence(dst, op, src + anchor, mpos - anchor, mpos - mcand, mlen, lane);
                ip = anchor = mpos + mlen;
                if (lane == 0 && ip - 2 >= seg && ip <= mfl)               // lz4/lz4.c:1146 insert ip-2
                    tab[hash13(ldg_u32_unaligned(src + ip - 2))] = (u16)(ip - 2 - seg);
                __syncwarp();
            }
        }
    }
    {   // last literals (lz4/lz4.c:1204-1231)
        const u32 last = n - anchor;
        if (lane == 0) dst[op] = (u8)((last < RUN_MASK ? last : RUN_MASK) << 4);
        op++;
        if (last >= RUN_MASK) op += v0_put_len_ext(dst + op, last - RUN_MASK, lane);
        for (u32 i = lane; i < last; i += 32) dst[op + i] = __ldg(src + anchor + i);
        op += last;
    }
    if (lane == 0) csz[chunk] = op;
}

// ---- decode ---------------------------------------------------------------------------

__global__ void __launch_bounds__(V0_WARPS * 32)
decompress_v0_kernel(const u8* __restrict__ in, const u64* __restrict__ in_off,
                     const u64* __restrict__ csz, const u64* __restrict__ out_off,
                     const u64* __restrict__ usz, u8* out, u32* __restrict__ status,
                     u32* __restrict__ bad_chunks, u32 n_chunks)
{
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 chunk = blockIdx.x * V0_WARPS + warp;
    if (chunk >= n_chunks) return;
    const u8* src = in + in_off[chunk];
    const u32 clen = (u32)csz[chunk];
    u8* dst = out + out_off[chunk];
    const u32 n = (u32)usz[chunk];
    u32 err = ST_OK;

    if (clen == n) {                                  // raw chunk (_lzbench/lzbench.cpp:311-315)
        coop_copy(dst, src, n, lane, 32);
    } else if (n == 0) {                              // lz4/lz4.c:1776
        err = (clen == 1 && __ldg(src) == 0) ? ST_OK : ST_EMPTY;
    } else if (clen == 0) {
        err = ST_EMPTY;
    } else {
        u32 ip = 0, op = 0;
        for (;;) {
            if (ip >= clen) { err = ST_TOKEN_OOB; break; }
            const u32 token = __ldg(src + ip++);
            u64 len = token >> 4;                      // 64-bit: a hostile 255-run cannot wrap it
            if (len == RUN_MASK) {
                if (ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                u32 s;
                do {
                    s = __ldg(src + ip++); len += s;
                    if (s == 255 && ip + RUN_MASK >= clen) { err = ST_LITLEN_OOB; break; }
                } while (s == 255);
                if (err) break;
            }
            const u64 cpy = op + len;
            if (cpy + MFLIMIT > n || ip + len + (2 + 1 + LASTLITERALS) > clen) {     // last sequence
                if (ip + len != clen || cpy > n) { err = ST_LAST_LITERALS; break; }
                for (u32 i = lane; i < (u32)len; i += 32) dst[op + i] = __ldg(src + ip + i);
                op = (u32)cpy;
                if (op != n) err = ST_SIZE;
                break;
            }
            for (u32 i = lane; i < (u32)len; i += 32) dst[op + i] = __ldg(src + ip + i);
            ip += (u32)len; op = (u32)cpy;
            const u32 off = __ldg(src + ip) | (__ldg(src + ip + 1) << 8);
            ip += 2;
            u64 ml64 = token & ML_MASK;
            if (ml64 == ML_MASK) {
                u32 s;
                do {
                    s = __ldg(src + ip++); ml64 += s;
                    if (ip + LASTLITERALS - 1 >= clen) { err = ST_MATCHLEN_OOB; break; }
                } while (s == 255);
                if (err) break;
            }
            ml64 += MINMATCH;
            if (off == 0 || off > op) { err = ST_OFFSET; break; }
            if (op + ml64 + LASTLITERALS > n) { err = ST_MATCH_END; break; }
            const u32 ml = (u32)ml64;
            __syncwarp();                              // literals above are visible to every lane
            const u8* m = dst + op - off;
            if (off >= ml) {                           // no overlap
                for (u32 i = lane; i < ml; i += 32) dst[op + i] = m[i];
            } else if (off >= 32) {                    // overlap, but a 32-byte step never reads its own writes
                for (u32 i = 0; i < ml; i += 32) {
                    if (i + lane < ml) dst[op + i + lane] = m[i + lane];
                    __syncwarp();
                }
            } else {                                   // periodic extension of the last `off` bytes
                u32 r = lane % off;
                const u32 step = 32 % off;
                for (u32 i = lane; i < ml; i += 32) {
                    dst[op + i] = m[r];
                    r += step; if (r >= off) r -= off;
                }
            }
            __syncwarp();
            op += ml;
        }
    }
    if (lane == 0) {
        status[chunk] = err;
        if (err && bad_chunks) atomicAdd(bad_chunks, 1u);
    }
}

}  // namespace lz4b200
