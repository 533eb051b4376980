#!/usr/bin/env python
"""Apply the lz4_b200 integration to a STAGED COPY of the reference harness (never to /root/reference,
never committed): the same edits INTEGRATION.md lists for a maintainer.  Four files, ~30 changed lines:

  lzbench.h        two trailing, optional members in compressor_desc_t (batch entry points), one new
                   comp_desc[] row "lz4_b200", LZBENCH_COMPRESSOR_COUNT + 1, "cuda" alias extended
  compressors.h    prototypes under BENCH_HAS_LZ4_B200 (NULL fallbacks otherwise, as for BENCH_HAS_NVCOMP)
  lzbench.cpp      lzbench_test hands the whole chunk list to the batch entry points when the row has them

usage: patch_lzbench.py <staged _lzbench dir>
"""
import re
import sys

d = sys.argv[1]


def edit(name, fn):
    p = f"{d}/{name}"
    s = open(p).read()
    t = fn(s)
    assert t != s, f"patch did not apply to {name}"
    open(p, "w").write(t)


def lzbench_h(s):
    s = s.replace("    deinit_func deinit;\n} compressor_desc_t;",
                  "    deinit_func deinit;\n"
                  "    compress_batch_func compress_batch;     /* optional: whole chunk list in one call */\n"
                  "    decompress_batch_func decompress_batch;\n} compressor_desc_t;")
    s = s.replace("typedef void (*deinit_func)(char* workmem);",
                  "typedef void (*deinit_func)(char* workmem);\n"
                  "typedef int64_t (*compress_batch_func)(const size_t* chunk_sizes, size_t n_chunks, size_t* compr_sizes, char* in, char* out, size_t outsize, char* workmem);\n"
                  "typedef int64_t (*decompress_batch_func)(const size_t* chunk_sizes, const size_t* compr_sizes, size_t n_chunks, char* in, char* out, char* workmem);")
    s = s.replace("#define LZBENCH_COMPRESSOR_COUNT 75", "#define LZBENCH_COMPRESSOR_COUNT 76")
    s = s.replace('    { "nvcomp_lz4", "1.2.2",       0,   5,    0,       0, lzbench_nvcomp_compress,     lzbench_nvcomp_decompress,     lzbench_nvcomp_init,     lzbench_nvcomp_deinit },\n',
                  '    { "nvcomp_lz4", "1.2.2",       0,   5,    0,       0, lzbench_nvcomp_compress,     lzbench_nvcomp_decompress,     lzbench_nvcomp_init,     lzbench_nvcomp_deinit },\n'
                  '    { "lz4_b200",   "0.1",         0,   0,    0,       0, lzbench_lz4_b200_compress,   lzbench_lz4_b200_decompress,   lzbench_lz4_b200_init,   lzbench_lz4_b200_deinit, lzbench_lz4_b200_compress_batch, lzbench_lz4_b200_decompress_batch },\n')
    s = s.replace('{ "cuda",  "cudaMemcpy/nvcomp_lz4,0,1,3,5" }', '{ "cuda",  "cudaMemcpy/nvcomp_lz4,0,1,3,5/lz4_b200" }')
    return s


def compressors_h(s):
    return s.replace("#endif // LZBENCH_COMPRESSORS_H",
                     "#ifdef BENCH_HAS_LZ4_B200\n"
                     '        extern "C" {\n'
                     "        char* lzbench_lz4_b200_init(size_t insize, size_t level, size_t);\n"
                     "        void lzbench_lz4_b200_deinit(char* workmem);\n"
                     "        int64_t lzbench_lz4_b200_compress(char *inbuf, size_t insize, char *outbuf, size_t outsize, size_t, size_t, char* workmem);\n"
                     "        int64_t lzbench_lz4_b200_decompress(char *inbuf, size_t insize, char *outbuf, size_t outsize, size_t, size_t, char* workmem);\n"
                     "        int64_t lzbench_lz4_b200_compress_batch(const size_t* chunk_sizes, size_t n_chunks, size_t* compr_sizes, char* in, char* out, size_t outsize, char* workmem);\n"
                     "        int64_t lzbench_lz4_b200_decompress_batch(const size_t* chunk_sizes, const size_t* compr_sizes, size_t n_chunks, char* in, char* out, char* workmem);\n"
                     "        }\n"
                     "#else\n"
                     "        #define lzbench_lz4_b200_init NULL\n"
                     "        #define lzbench_lz4_b200_deinit NULL\n"
                     "        #define lzbench_lz4_b200_compress NULL\n"
                     "        #define lzbench_lz4_b200_decompress NULL\n"
                     "        #define lzbench_lz4_b200_compress_batch NULL\n"
                     "        #define lzbench_lz4_b200_decompress_batch NULL\n"
                     "#endif\n\n#endif // LZBENCH_COMPRESSORS_H")


def lzbench_cpp(s):
    s = s.replace("            complen = lzbench_compress(params, chunk_sizes, desc->compress, compr_sizes, inbuf, compbuf, comprsize, param1, param2, workmem);",
                  "            if (desc->compress_batch) { compr_sizes.resize(chunk_sizes.size()); complen = desc->compress_batch(&chunk_sizes[0], chunk_sizes.size(), &compr_sizes[0], (char*)inbuf, (char*)compbuf, comprsize, workmem); }\n"
                  "            else complen = lzbench_compress(params, chunk_sizes, desc->compress, compr_sizes, inbuf, compbuf, comprsize, param1, param2, workmem);")
    s = s.replace("            decomplen = lzbench_decompress(params, chunk_sizes, desc->decompress, compr_sizes, compbuf, decomp, param1, param2, workmem);",
                  "            if (desc->decompress_batch) decomplen = desc->decompress_batch(&chunk_sizes[0], &compr_sizes[0], compr_sizes.size(), (char*)compbuf, (char*)decomp, workmem);\n"
                  "            else decomplen = lzbench_decompress(params, chunk_sizes, desc->decompress, compr_sizes, compbuf, decomp, param1, param2, workmem);")
    return s


edit("lzbench.h", lzbench_h)
edit("compressors.h", compressors_h)
edit("lzbench.cpp", lzbench_cpp)
print("patched", d)
