/*
 * oracle/ref_frame.c -- TEST INFRASTRUCTURE.  Thin shim around the reference's own lz4frame.c (compiled from the
 * reference tree into oracle/_ref/liblz4fref.so by oracle/Makefile) so the tests can make and read LZ4 frames with
 * the reference implementation: LZ4F_compressFrame (lz4/lz4frame.c:429-470) and LZ4F_decompress (lz4/lz4frame.c:1384-1830).
 */
#include <stddef.h>
#include <string.h>
#include "lz4frame.h"

size_t ref_lz4f_bound(size_t n, int block_id, int content_checksum, int block_checksum)
{
    LZ4F_preferences_t p; memset(&p, 0, sizeof p);
    p.frameInfo.blockSizeID = (LZ4F_blockSizeID_t)block_id;
    p.frameInfo.contentChecksumFlag = (LZ4F_contentChecksum_t)content_checksum;
    p.frameInfo.blockChecksumFlag = (LZ4F_blockChecksum_t)block_checksum;
    return LZ4F_compressFrameBound(n, &p);
}

/* returns frame size, or 0 on error */
size_t ref_lz4f_compress(const void* src, size_t n, void* dst, size_t cap, int block_id, int independent,
                         int content_checksum, int block_checksum, int with_content_size)
{
    LZ4F_preferences_t p; memset(&p, 0, sizeof p);
    p.frameInfo.blockSizeID = (LZ4F_blockSizeID_t)block_id;
    p.frameInfo.blockMode = independent ? LZ4F_blockIndependent : LZ4F_blockLinked;
    p.frameInfo.contentChecksumFlag = (LZ4F_contentChecksum_t)content_checksum;
    p.frameInfo.blockChecksumFlag = (LZ4F_blockChecksum_t)block_checksum;
    p.frameInfo.contentSize = with_content_size ? n : 0;
    size_t r = LZ4F_compressFrame(dst, cap, src, n, &p);
    return LZ4F_isError(r) ? 0 : r;
}

/* returns decoded size, or (size_t)-1 on error / truncated frame */
size_t ref_lz4f_decompress(const void* src, size_t n, void* dst, size_t cap)
{
    LZ4F_dctx* d = NULL;
    if (LZ4F_isError(LZ4F_createDecompressionContext(&d, LZ4F_VERSION))) return (size_t)-1;
    const char* ip = (const char*)src; char* op = (char*)dst;
    size_t in_left = n, out_left = cap, hint = 1;
    while (hint != 0 && in_left > 0) {
        size_t isz = in_left, osz = out_left;
        hint = LZ4F_decompress(d, op, &osz, ip, &isz, NULL);
        if (LZ4F_isError(hint)) { LZ4F_freeDecompressionContext(d); return (size_t)-1; }
        ip += isz; in_left -= isz; op += osz; out_left -= osz;
        if (isz == 0 && osz == 0) break;
    }
    LZ4F_freeDecompressionContext(d);
    return hint == 0 ? (size_t)(op - (char*)dst) : (size_t)-1;
}
