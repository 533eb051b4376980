/*
 * oracle/lz4f_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the LZ4 frame container as the reference implements it (lz4/lz4frame.c):
 * constants lz4/lz4frame.c:210-211, header checksum lz4/lz4frame.c:294-298, header written by
 * LZ4F_compressBegin_usingCDict lz4/lz4frame.c:598-700, blocks with LE32 size + "uncompressed" bit by LZ4F_makeBlock
 * lz4/lz4frame.c:740-763, end mark and optional content checksum by LZ4F_compressEnd lz4/lz4frame.c:986-1025, decoder
 * LZ4F_decodeHeader lz4/lz4frame.c:1113-1204 and LZ4F_decompress lz4/lz4frame.c:1384-1830, and xxHash32
 * (XXH32_endian_align, lz4/xxhash.c:352-389) for the header byte and
 * the checksums.  Independent blocks only, as in lzbench's chunked mode.  Pinned against oracle/_ref/liblz4fref.so
 * (the reference's lz4frame.c compiled here) by tests/test_frame.py: both directions cross-decode, header bytes equal.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include "lz4_oracle.h"

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rd32le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline void wr32le(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

/* xxHash32, XXH32_endian_align lz4/xxhash.c:352-389 */
uint32_t oracle_xxh32(const void* data, size_t len, uint32_t seed)
{
    const uint32_t P1 = 2654435761U, P2 = 2246822519U, P3 = 3266489917U, P4 = 668265263U, P5 = 374761393U;
    const uint8_t* p = (const uint8_t*)data; const uint8_t* end = p + len;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t* lim = end - 16;
        do {
            v1 = rotl32(v1 + rd32le(p) * P2, 13) * P1; p += 4;
            v2 = rotl32(v2 + rd32le(p) * P2, 13) * P1; p += 4;
            v3 = rotl32(v3 + rd32le(p) * P2, 13) * P1; p += 4;
            v4 = rotl32(v4 + rd32le(p) * P2, 13) * P1; p += 4;
        } while (p <= lim);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else h = seed + P5;
    h += (uint32_t)len;
    while (p + 4 <= end) { h = rotl32(h + rd32le(p) * P3, 17) * P4; p += 4; }
    while (p < end) { h = rotl32(h + (*p) * P5, 11) * P1; p++; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

static size_t block_bytes(int id) { return id >= 4 && id <= 7 ? (size_t)65536 << (2 * (id - 4)) : 0; }

size_t oracle_lz4f_bound(size_t n, int block_id)
{
    size_t b = block_bytes(block_id);
    if (!b) return 0;
    return 19 + (n / b + 1) * 4 + n + 4 + 4;
}

/* Writes a frame with independent blocks.  Returns the frame size or 0. */
size_t oracle_lz4f_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int block_id,
                            int content_checksum, int with_content_size)
{
    const size_t bsz = block_bytes(block_id);
    if (!bsz || cap < oracle_lz4f_bound(n, block_id)) return 0;
    uint8_t* op = dst;
    wr32le(op, 0x184D2204U); op += 4;
    uint8_t* desc = op;
    *op++ = (uint8_t)((1 << 6) | (1 << 5) | ((with_content_size ? 1 : 0) << 3) | ((content_checksum ? 1 : 0) << 2));
    *op++ = (uint8_t)(block_id << 4);
    if (with_content_size) { wr32le(op, (uint32_t)n); wr32le(op + 4, (uint32_t)((uint64_t)n >> 32)); op += 8; }
    *op = (uint8_t)(oracle_xxh32(desc, (size_t)(op - desc), 0) >> 8); op++;
    for (size_t off = 0; off < n; off += bsz) {
        const size_t part = n - off < bsz ? n - off : bsz;
        int c = oracle_lz4_compress(src + off, (int)part, op + 4, (int)part - 1);      /* lz4/lz4frame.c:747 dstCapacity = srcSize-1 */
        if (c <= 0) { wr32le(op, (uint32_t)part | 0x80000000U); memcpy(op + 4, src + off, part); op += 4 + part; }
        else { wr32le(op, (uint32_t)c); op += 4 + (size_t)c; }
    }
    wr32le(op, 0); op += 4;
    if (content_checksum) { wr32le(op, oracle_xxh32(src, n, 0)); op += 4; }
    return (size_t)(op - dst);
}

/* Reads a frame (independent blocks; block / content checksums verified when present).
 * Returns the decoded size or (size_t)-1. */
size_t oracle_lz4f_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    if (n < 7 || rd32le(src) != 0x184D2204U) return (size_t)-1;
    const uint8_t flg = src[4], bd = src[5];
    if ((flg >> 6) != 1 || (flg & 2) || !(flg & 0x20)) return (size_t)-1;     /* version, reserved, linked blocks */
    const int bck = (flg >> 4) & 1, csz = (flg >> 3) & 1, cck = (flg >> 2) & 1, did = flg & 1;
    const size_t bmax = block_bytes((bd >> 4) & 7);
    if (!bmax || (bd & 0x8F)) return (size_t)-1;
    size_t hp = 6 + (csz ? 8 : 0) + (did ? 4 : 0);
    if (n < hp + 1) return (size_t)-1;
    if (src[hp] != (uint8_t)(oracle_xxh32(src + 4, hp - 4, 0) >> 8)) return (size_t)-1;
    uint64_t want = 0;
    if (csz) want = (uint64_t)rd32le(src + 6) | ((uint64_t)rd32le(src + 10) << 32);
    size_t ip = hp + 1, op = 0;
    for (;;) {
        if (ip + 4 > n) return (size_t)-1;
        const uint32_t w = rd32le(src + ip); ip += 4;
        if (w == 0) break;
        const size_t sz = w & 0x7FFFFFFFU;
        if (sz > bmax || ip + sz + (bck ? 4 : 0) > n) return (size_t)-1;
        if (bck && rd32le(src + ip + sz) != oracle_xxh32(src + ip, sz, 0)) return (size_t)-1;
        if (w & 0x80000000U) {
            if (op + sz > cap) return (size_t)-1;
            memcpy(dst + op, src + ip, sz); op += sz;
        } else {
            /* lz4/lz4frame.c:1654-1712: a block is always decoded with the frame's block size as capacity (into the
             * caller's buffer when that much room is left, else into the context's own buffer and copied) */
            if (cap - op >= bmax) {
                int r = oracle_lz4_decompress_safe(src + ip, (int)sz, dst + op, (int)bmax);
                if (r < 0) return (size_t)-1;
                op += (size_t)r;
            } else {
                uint8_t* tmp = (uint8_t*)malloc(bmax);
                if (!tmp) return (size_t)-1;
                int r = oracle_lz4_decompress_safe(src + ip, (int)sz, tmp, (int)bmax);
                if (r < 0 || (size_t)r > cap - op) { free(tmp); return (size_t)-1; }
                memcpy(dst + op, tmp, (size_t)r);
                free(tmp);
                op += (size_t)r;
            }
        }
        ip += sz + (bck ? 4 : 0);
    }
    if (cck) { if (ip + 4 > n || rd32le(src + ip) != oracle_xxh32(dst, op, 0)) return (size_t)-1; }
    if (csz && want != op) return (size_t)-1;
    return op;
}
