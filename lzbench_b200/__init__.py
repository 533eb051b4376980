"""lzbench_b200 -- B200-native LZ4 block codec for lzbench's chunked mode (-b#).

The product is the C-ABI shared library ``lzbench_b200/_lib/liblz4b200.so`` (CUDA kernels for
sm_100a + C/C++ host code, ``include/lz4b200.h``).  This package is only the thin ctypes
binding used by the tests and ``bench.py``; it holds no codec logic and no CPU fallback: if
the library is missing, import fails loudly.
"""
from .codec import (  # noqa: F401
    Lz4B200,
    LibraryMissing,
    chunk_list,
    compress_bound,
    corpus,
    get_compress_bound,
    last_error,
    lib,
    lib_path,
    lzbench_compress,
    lzbench_decompress,
)

__all__ = [
    "Lz4B200", "LibraryMissing", "chunk_list", "compress_bound", "corpus", "get_compress_bound", "last_error",
    "lib", "lib_path", "lzbench_compress", "lzbench_decompress",
]
