// frame.cuh -- LZ4 frame container around the chunked codec.
// What the reference implements in lz4/lz4frame.c: LZ4F_compressFrame (lz4/lz4frame.c:429-470: header, one LE32-prefixed
// block per blockSize bytes via LZ4F_makeBlock lz4/lz4frame.c:740-763, end mark, optional content checksum) and
// LZ4F_decompress (lz4/lz4frame.c:1384-1830).
// Here a frame with independent blocks IS the chunked hot path plus a few bytes of framing, so the block work stays
// on the GPU: sizes -> scan -> header words -> the existing compaction, and for decoding a size pass (a frame does
// not store decoded block sizes) -> scan -> the existing decoders.  xxHash32 (lz4/xxhash.c:352-389) is written from its
// published definition: per block on the device (blocks are independent), whole-content on the host (one serial chain).
#pragma once
#include "common.cuh"

namespace lz4b200 {

constexpr u32 F_MAGIC = 0x184D2204u;
constexpr u32 F_RAW = 0x80000000u;
constexpr u32 XP1 = 2654435761u, XP2 = 2246822519u, XP3 = 3266489917u, XP4 = 668265263u, XP5 = 374761393u;

__host__ __device__ inline u32 x_rotl(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
__host__ __device__ inline u32 x_round(u32 v, u32 x) { return x_rotl(v + x * XP2, 13) * XP1; }
__host__ __device__ inline u32 x_final(u32 h)
{
    h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    return h;
}

// ---- compress side -------------------------------------------------------------------------------------------
// Block i occupies 4 + csz[i] (+4 with block checksums) bytes in the frame.  A block that did not shrink is stored
// uncompressed (lz4/lz4frame.c:747-756: the block compressor is given srcSize-1 bytes and a failure means "store");
// clamping csz to usz makes the compaction kernels take the raw branch they already have for lzbench's rule.
__global__ void frame_sizes_kernel(u64* __restrict__ csz, const u64* __restrict__ usz, u64* __restrict__ fsz, u32 n, u32 extra)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 c = csz[i];
    const u64 u = usz[i];
    if (c >= u) { c = u; csz[i] = c; }
    fsz[i] = c + 4 + extra;
}

// off[i] comes in as the block's position in the frame; write the LE32 block header there and leave off[i] pointing
// at the payload for the compaction kernel.
__global__ void frame_headers_kernel(u8* __restrict__ out, u64* __restrict__ off, const u64* __restrict__ csz,
                                     const u64* __restrict__ usz, u32 n)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 o = off[i];
    const u32 w = (u32)csz[i] | (csz[i] == usz[i] ? F_RAW : 0u);
    out[o] = (u8)w; out[o + 1] = (u8)(w >> 8); out[o + 2] = (u8)(w >> 16); out[o + 3] = (u8)(w >> 24);
    off[i] = o + 4;
}

// xxHash32 (seed 0) of every block payload, one thread per block: the four accumulators are one thread's registers
// and the 16-byte stripes stream through unaligned word loads.  verify = 0: store the digest behind the payload
// (LZ4F_makeBlock, lz4/lz4frame.c:757-761); verify = 1: compare with the stored one and count mismatches (LZ4F_decompress,
// lz4/lz4frame.c:1590-1601).
__global__ void frame_block_xxh_kernel(u8* __restrict__ buf, const u64* __restrict__ off, const u64* __restrict__ len,
                                       u32 n, int verify, u32* __restrict__ bad)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8* p = buf + off[i];
    const u64 L = len[i];
    u64 k = 0;
    u32 h;
    if (L >= 16) {
        u32 v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0u - XP1;
        for (; k + 16 <= L; k += 16) {
            v1 = x_round(v1, ldg_u32_unaligned(p + k));
            v2 = x_round(v2, ldg_u32_unaligned(p + k + 4));
            v3 = x_round(v3, ldg_u32_unaligned(p + k + 8));
            v4 = x_round(v4, ldg_u32_unaligned(p + k + 12));
        }
        h = x_rotl(v1, 1) + x_rotl(v2, 7) + x_rotl(v3, 12) + x_rotl(v4, 18);
    } else h = XP5;
    h += (u32)L;
    for (; k + 4 <= L; k += 4) h = x_rotl(h + ldg_u32_unaligned(p + k) * XP3, 17) * XP4;
    for (; k < L; k++) h = x_rotl(h + p[k] * XP5, 11) * XP1;
    h = x_final(h);
    u8* q = buf + off[i] + L;
    if (verify) {
        const u32 w = (u32)q[0] | ((u32)q[1] << 8) | ((u32)q[2] << 16) | ((u32)q[3] << 24);
        if (w != h) atomicAdd(bad, 1u);
    } else {
        q[0] = (u8)h; q[1] = (u8)(h >> 8); q[2] = (u8)(h >> 16); q[3] = (u8)(h >> 24);
    }
}

// ---- decompress side -----------------------------------------------------------------------------------------
// A frame stores only the compressed size of each block.  The decoders want the decoded size (output placement, and
// their end-of-block checks are against it), so one warp per block walks the sequence headers first: a 1 KiB window
// of the block in shared memory, lane 0 runs the token / length-byte state machine, all lanes refill the window when
// the walk leaves it.  hdr[i] = block header word (bit 31 = stored uncompressed).  usz[i] = decoded size, clamped to
// block_max (the decoder then reports the block as malformed); csz[i] = payload size.  A compressed block that is
// not smaller than its content cannot be told from a stored one by the decoders' size rule and is counted in
// *unsupported (the reference encoder never emits one: lz4/lz4frame.c:747).
constexpr int FS_WARPS = 8;
constexpr u32 FS_WIN = 1024;

__global__ void __launch_bounds__(FS_WARPS * 32)
frame_block_sizes_kernel(const u8* __restrict__ frame, u64 frame_bytes, const u64* __restrict__ off, const u32* __restrict__ hdr,
                         u64* __restrict__ csz, u64* __restrict__ usz, u32 n, u32 block_max, u32* __restrict__ unsupported)
{
    __shared__ __align__(16) u8 win_all[FS_WARPS][FS_WIN];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 i = blockIdx.x * FS_WARPS + warp;
    if (i >= n) return;
    u8* win = win_all[warp];
    const u32 w = hdr[i], clen = w & ~F_RAW;
    if (lane == 0) csz[i] = clen;
    if (w & F_RAW) { if (lane == 0) usz[i] = clen; return; }
    const u64 src = off[i];                         // payload position in the frame
    u32 ip = 0, op = 0, lit = 0, st = 0, tok = 0;   // st: 0 token, 1 literal-length bytes, 2 match-length bytes
    bool done = false;
    while (!done) {
        const u64 g0 = (src + ip) & ~(u64)15;       // frame is in a cudaMalloc'd buffer: 16-byte aligned windows
        const u32 wb = (u32)(g0 - src + 0);          // may wrap below zero on the first window; used modulo 2^32
        #pragma unroll
        for (u32 r = 0; r < FS_WIN / 512; r++) {
            const u64 g = g0 + r * 512 + lane * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (g + 16 <= ((frame_bytes + 15) & ~(u64)15)) v = __ldg((const uint4*)(frame + g));
            *(uint4*)(win + r * 512 + lane * 16) = v;
        }
        __syncwarp();
        if (lane == 0) {
            while (true) {
                if (ip >= clen) { done = true; break; }          // ran off the block: malformed, the decoder says so
                const u32 rel = ip - wb;
                if (rel >= FS_WIN) break;
                const u32 b = win[rel];
                ip++;
                bool after_lit = false;
                if (st == 0) { tok = b; lit = b >> 4; if (lit == 15) st = 1; else after_lit = true; }
                else if (st == 1) { lit += b; if (b != 255) after_lit = true; if (lit > clen) { done = true; break; } }
                else { op = min(op + b, block_max + 1); if (b != 255) st = 0; }
                if (after_lit) {
                    ip += lit; op = min(op + lit, block_max + 1);
                    if (ip >= clen) { done = true; break; }      // last sequence: literals only
                    ip += 2;
                    op = min(op + 4 + (tok & 15), block_max + 1);
                    st = (tok & 15) == 15 ? 2 : 0;
                }
            }
        }
        __syncwarp();
        ip = __shfl_sync(FULL, ip, 0);
        done = __shfl_sync(FULL, (int)done, 0) != 0;
        st = __shfl_sync(FULL, st, 0);                           // (only lane 0 uses the rest of the state)
    }
    if (lane == 0) {
        if (op > block_max) op = block_max;
        if (op == clen) atomicAdd(unsupported, 1u);
        usz[i] = op;
    }
}

// The decoded size of a slice of blocks is known only on the device, and the decoders are launched without a host
// round trip: if the running total would leave the output allocation, empty the slice (sizes 0 = nothing is read or
// written); the host sees the same total one event later and fails the call.
__global__ void frame_guard_kernel(u64* __restrict__ usz, u64* __restrict__ csz, u32 n, const u64* __restrict__ total, u64 limit)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && *total > limit) { usz[i] = 0; csz[i] = 0; }
}

// ---- host side: xxHash32 and the frame header ------------------------------------------------------------------
inline u32 h_rd32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
inline void h_wr32(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }

inline u32 xxh32_host(const void* data, size_t len)
{
    const u8* p = (const u8*)data; const u8* const end = p + len;
    u32 h;
    if (len >= 16) {
        u32 v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0u - XP1;
        for (; p + 16 <= end; p += 16) {
            v1 = x_round(v1, h_rd32(p)); v2 = x_round(v2, h_rd32(p + 4));
            v3 = x_round(v3, h_rd32(p + 8)); v4 = x_round(v4, h_rd32(p + 12));
        }
        h = x_rotl(v1, 1) + x_rotl(v2, 7) + x_rotl(v3, 12) + x_rotl(v4, 18);
    } else h = XP5;
    h += (u32)len;
    for (; p + 4 <= end; p += 4) h = x_rotl(h + h_rd32(p) * XP3, 17) * XP4;
    for (; p < end; p++) h = x_rotl(h + *p * XP5, 11) * XP1;
    return x_final(h);
}

inline size_t frame_block_bytes(int id) { return id >= 4 && id <= 7 ? (size_t)65536 << (2 * (id - 4)) : 0; }

struct FrameInfo {
    size_t header_bytes = 0, block_max = 0;
    bool block_checksum = false, content_checksum = false, has_content_size = false;
    u64 content_size = 0;
};

// lz4/lz4frame.c:598-700 (LZ4F_compressBegin_usingCDict, header bytes :668-700) -- version 01, independent blocks, no dictionary
inline size_t frame_write_header(u8* dst, int block_id, bool bck, bool cck, bool with_size, u64 content_size)
{
    h_wr32(dst, F_MAGIC);
    u8* p = dst + 4;
    *p++ = (u8)((1u << 6) | (1u << 5) | ((bck ? 1u : 0u) << 4) | ((with_size ? 1u : 0u) << 3) | ((cck ? 1u : 0u) << 2));
    *p++ = (u8)(block_id << 4);
    if (with_size) { h_wr32(p, (u32)content_size); h_wr32(p + 4, (u32)(content_size >> 32)); p += 8; }
    *p = (u8)(xxh32_host(dst + 4, (size_t)(p - dst - 4)) >> 8);
    return (size_t)(p + 1 - dst);
}

// lz4/lz4frame.c:1113-1204 (LZ4F_decodeHeader).  Returns 0, or a negative code with *why set.
inline int frame_read_header(const u8* src, size_t n, FrameInfo* fi, const char** why)
{
    if (n < 7) { *why = "frame shorter than its header"; return -1; }
    const u32 magic = h_rd32(src);
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) { *why = "skippable frame: nothing to decode"; return -2; }
    if (magic != F_MAGIC) { *why = "not an LZ4 frame (magic)"; return -1; }
    const u8 flg = src[4], bd = src[5];
    if ((flg >> 6) != 1) { *why = "unknown frame version"; return -1; }
    if (flg & 2) { *why = "reserved header bit set"; return -1; }
    if (!(flg & 0x20)) { *why = "linked blocks: frame is not chunk-independent (outside the chunked path)"; return -2; }
    if (flg & 1) { *why = "frame needs a dictionary (outside the chunked path)"; return -2; }
    fi->block_checksum = (flg >> 4) & 1; fi->has_content_size = (flg >> 3) & 1; fi->content_checksum = (flg >> 2) & 1;
    if (bd & 0x8F) { *why = "reserved header bit set"; return -1; }
    fi->block_max = frame_block_bytes((bd >> 4) & 7);
    if (!fi->block_max) { *why = "invalid block size id"; return -1; }
    const size_t hp = 6 + (fi->has_content_size ? 8 : 0);
    if (n < hp + 1) { *why = "frame shorter than its header"; return -1; }
    if (src[hp] != (u8)(xxh32_host(src + 4, hp - 4) >> 8)) { *why = "header checksum mismatch"; return -1; }
    if (fi->has_content_size) fi->content_size = (u64)h_rd32(src + 6) | ((u64)h_rd32(src + 10) << 32);
    fi->header_bytes = hp + 1;
    return 0;
}

}  // namespace lz4b200
