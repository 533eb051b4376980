/*
 * oracle/ref_plugin.c -- TEST/BASELINE INFRASTRUCTURE.  The reference's two lz4 plugin wrappers
 * (_lzbench/compressors.cpp:343-346 lzbench_lz4_compress, :358-362 lzbench_lz4_decompress)
 * restated as C trampolines around the reference's OWN compiled functions, so the CPU baseline
 * (bench.py cpu_baseline / --impl reference) times exactly what `lzbench -elz4 -b64` times per
 * chunk.  Linked with the reference's lz4.c into oracle/_ref/libref_plugin.so by oracle/Makefile.
 */
#include <stddef.h>
#include <stdint.h>

int LZ4_compress_default(const char* src, char* dst, int srcSize, int dstCapacity);   /* lz4/lz4.h */
int LZ4_decompress_fast(const char* src, char* dst, int originalSize);                /* lz4/lz4.h */

int64_t ref_plugin_lz4_compress(char* in, size_t insize, char* out, size_t outsize, size_t level, size_t p2, char* wm)
{
    (void)level; (void)p2; (void)wm;
    return LZ4_compress_default(in, out, (int)insize, outsize > 0x7FFFFFFF ? 0x7FFFFFFF : (int)outsize);
}

int64_t ref_plugin_lz4_decompress(char* in, size_t insize, char* out, size_t outsize, size_t level, size_t p2, char* wm)
{
    (void)insize; (void)level; (void)p2; (void)wm;
    LZ4_decompress_fast(in, out, (int)outsize);      /* return value ignored, as in the reference */
    return (int64_t)outsize;
}
