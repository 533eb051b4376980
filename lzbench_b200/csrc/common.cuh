// common.cuh -- device helpers shared by every kernel of the LZ4 chunked path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lz4b200 {

typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

constexpr unsigned FULL = 0xFFFFFFFFu;

// LZ4 block-format constants (lz4/lz4.c:214-235, lz4/lz4.h:556-558)
constexpr u32 MINMATCH = 4;
constexpr u32 LASTLITERALS = 5;
constexpr u32 MFLIMIT = 12;
constexpr u32 MIN_LENGTH = MFLIMIT + 1;
constexpr u32 ML_MASK = 15;
constexpr u32 RUN_MASK = 15;

// Per-chunk decode status written by the decode kernels (0 = ok).  Values 1..8 are the
// rule numbers of oracle_lz4_validate so the parity tests can compare them directly.
enum DecodeStatus : u32 {
    ST_OK = 0, ST_EMPTY = 1, ST_TOKEN_OOB = 2, ST_LITLEN_OOB = 3, ST_LAST_LITERALS = 4,
    ST_MATCHLEN_OOB = 5, ST_OFFSET = 6, ST_MATCH_END = 7, ST_SIZE = 8
};

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

// 4 bytes at an arbitrary global address through the read-only path: two aligned loads and
// a funnel shift.  Touches only aligned words that contain a requested byte.
__device__ __forceinline__ u32 ldg_u32_unaligned(const u8* p)
{
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const u32* w = reinterpret_cast<const u32*>(a & ~uintptr_t(3));
    u32 sh = (u32)(a & 3) * 8;
    u32 lo = __ldg(w);
    u32 hi = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// LZ4 4-byte hash to 13 bits (lz4/lz4.c:698-702, byU16 table)
__device__ __forceinline__ u32 hash13(u32 v) { return (v * 2654435761u) >> 19; }

// Cooperative global->global copy of n bytes by `nthr` threads (tid in [0,nthr)).
// 16-byte stores once dst is aligned; the source is realigned with funnel shifts, so
// src and dst may have any relative alignment.  Aligned source words that contain a
// requested byte are the only ones read.
__device__ __forceinline__ void coop_copy(u8* dst, const u8* src, u64 n, u32 tid, u32 nthr)
{
    u64 head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > n) head = n;
    if (tid < head) dst[tid] = src[tid];
    dst += head; src += head; n -= head;
    u64 nvec = n >> 4;
    uintptr_t sa = reinterpret_cast<uintptr_t>(src);
    u32 sh = (u32)(sa & 3) * 8;
    const u32* s32 = reinterpret_cast<const u32*>(sa & ~uintptr_t(3));
    uint4* d128 = reinterpret_cast<uint4*>(dst);
    if (sh == 0 && (sa & 15) == 0) {
        const uint4* s128 = reinterpret_cast<const uint4*>(src);
        for (u64 j = tid; j < nvec; j += nthr) d128[j] = __ldg(s128 + j);
    } else {
        for (u64 j = tid; j < nvec; j += nthr) {
            const u32* p = s32 + j * 4;
            u32 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3);
            u32 w4 = sh ? __ldg(p + 4) : 0u;
            uint4 o;
            o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
            d128[j] = o;
        }
    }
    u64 done = nvec << 4;
    u64 tail = n - done;
    if (tid < tail) dst[done + tid] = src[done + tid];
}

}  // namespace lz4b200
