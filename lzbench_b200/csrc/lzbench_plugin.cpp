// lzbench_plugin.cpp -- the lzbench-facing wrappers for the comp_desc[] row "lz4_b200".
//
// Same shape as the reference's wrappers for its own GPU codec
// (_lzbench/compressors.cpp:1856-2010 lzbench_nvcomp_init/_deinit/_compress/_decompress):
// init creates the device context and returns it as `workmem`; compress/decompress take
// the plugin signature (_lzbench/lzbench.h:113) and return bytes produced, <= 0 on failure.
// Plain C++11 on top of the C ABI only, because the reference compiles compressors.cpp with
// -std=c++11 (reference Makefile:327-329).
#include "../../include/lz4b200.h"

#include <cstdio>
#include <cstdlib>

namespace {
struct Workmem { lz4b200_ctx* ctx; };
}

extern "C" {

char* lzbench_lz4_b200_init(size_t insize, size_t level, size_t param2)
{
    (void)param2;
    Workmem* w = new Workmem();
    // insize is min(-b chunk, file size) (_lzbench/lzbench.cpp:344,350): a chunk-size hint only.
    if (lz4b200_create(0, 0, insize, &w->ctx) != 0) {
        fprintf(stderr, "lz4_b200: %s\n", lz4b200_last_error());
        delete w;
        return NULL;
    }
    // the harness reuses its malloc'ed buffers for every iteration: page-lock them on first sight
    // (opt out with LZ4B200_HOST_REGISTER=0)
    const char* hr = getenv("LZ4B200_HOST_REGISTER");
    lz4b200_set_option(w->ctx, "host_register", (hr && hr[0] == '0') ? "0" : "1");
    // a row with levels hands LZ4_compress_fast's acceleration down as `level` (the lz4fast rows,
    // _lzbench/compressors.cpp:348-356); the plain lz4_b200 row has level 0 = acceleration 1
    if (level > 1) { char buf[24]; snprintf(buf, sizeof buf, "%zu", level); lz4b200_set_option(w->ctx, "accel", buf); }
    return reinterpret_cast<char*>(w);
}

void lzbench_lz4_b200_deinit(char* workmem)
{
    Workmem* w = reinterpret_cast<Workmem*>(workmem);
    if (!w) return;
    lz4b200_destroy(w->ctx);
    delete w;
}

int64_t lzbench_lz4_b200_compress(char* in, size_t insize, char* out, size_t outsize,
                                  size_t level, size_t param2, char* workmem)
{
    (void)level; (void)param2;
    Workmem* w = reinterpret_cast<Workmem*>(workmem);
    if (!w) return -1;
    size_t clen = 0;
    int64_t r = lz4b200_compress_batch(w->ctx, in, &insize, 1, out, outsize, &clen);
    return r;
}

int64_t lzbench_lz4_b200_decompress(char* in, size_t insize, char* out, size_t outsize,
                                    size_t level, size_t param2, char* workmem)
{
    (void)level; (void)param2;
    Workmem* w = reinterpret_cast<Workmem*>(workmem);
    if (!w) return -1;
    return lz4b200_decompress_batch(w->ctx, in, &insize, &outsize, 1, out);
}

int64_t lzbench_lz4_b200_compress_batch(const size_t* chunk_sizes, size_t n_chunks, size_t* compr_sizes,
                                        char* in, char* out, size_t outsize, char* workmem)
{
    Workmem* w = reinterpret_cast<Workmem*>(workmem);
    if (!w) return -1;
    return lz4b200_compress_batch(w->ctx, in, chunk_sizes, n_chunks, out, outsize, compr_sizes);
}

int64_t lzbench_lz4_b200_decompress_batch(const size_t* chunk_sizes, const size_t* compr_sizes, size_t n_chunks,
                                          char* in, char* out, char* workmem)
{
    Workmem* w = reinterpret_cast<Workmem*>(workmem);
    if (!w) return -1;
    return lz4b200_decompress_batch(w->ctx, in, compr_sizes, chunk_sizes, n_chunks, out);
}

}  // extern "C"
