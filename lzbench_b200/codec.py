"""ctypes binding of include/lz4b200.h plus the host-side mirror of lzbench's chunk loops.

Names follow the reference: ``lzbench_compress`` / ``lzbench_decompress`` take the same
arguments as _lzbench/lzbench.cpp:266-298 / :301-329 (chunk_sizes, compr_sizes, inbuf,
outbuf, workmem) and return the same integers, so the parity tests read like the reference's
own round-trip check (:433-458).  All compute happens in the shared library.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "_lib", "liblz4b200.so")

CORPUS_KINDS = {"mixed": 0, "random": 1, "repeat": 2, "zero": 3}


class LibraryMissing(ImportError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(lib_path):
        raise LibraryMissing(
            f"{lib_path} not found: build it with `make` (or __graft_entry__.build()); "
            "there is no CPU fallback for the lz4_b200 path")
    L = C.CDLL(lib_path)
    sz, vp, i64, u64 = C.c_size_t, C.c_void_p, C.c_int64, C.c_uint64
    L.lz4b200_version.restype = C.c_char_p
    L.lz4b200_last_error.restype = C.c_char_p
    L.lz4b200_compress_bound.restype = sz
    L.lz4b200_compress_bound.argtypes = [sz]
    L.lz4b200_device_count.restype = C.c_int
    L.lz4b200_create.restype = C.c_int
    L.lz4b200_create.argtypes = [C.c_int, sz, sz, C.POINTER(vp)]
    L.lz4b200_destroy.argtypes = [vp]
    L.lz4b200_destroy.restype = None
    L.lz4b200_compress_batch.restype = i64
    L.lz4b200_compress_batch.argtypes = [vp, vp, vp, sz, vp, sz, vp]
    L.lz4b200_decompress_batch.restype = i64
    L.lz4b200_decompress_batch.argtypes = [vp, vp, vp, vp, sz, vp]
    L.lz4b200_compress_device.restype = C.c_int
    L.lz4b200_compress_device.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp]
    L.lz4b200_decompress_device.restype = C.c_int
    L.lz4b200_decompress_device.argtypes = [vp, vp, vp, vp, sz, sz, vp, vp, vp]
    L.lz4b200_launch_count.restype = u64
    L.lz4b200_launch_count.argtypes = [vp]
    L.lz4b200_kernel_times_ms.restype = C.c_int
    L.lz4b200_kernel_times_ms.argtypes = [vp, vp, C.c_int]
    L.lz4b200_debug_prof.restype = C.c_int
    L.lz4b200_debug_prof.argtypes = [vp, vp]
    L.lz4b200_set_option.restype = C.c_int
    L.lz4b200_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.lz4b200_unregister_host.restype = C.c_int
    L.lz4b200_unregister_host.argtypes = [vp, vp]
    L.lz4b200_compress_device_bound.restype = sz
    L.lz4b200_compress_device_bound.argtypes = [sz, sz]
    L.lzbench_lz4_b200_init.restype = vp
    L.lzbench_lz4_b200_init.argtypes = [sz, sz, sz]
    L.lzbench_lz4_b200_deinit.restype = None
    L.lzbench_lz4_b200_deinit.argtypes = [vp]
    for name in ("lzbench_lz4_b200_compress", "lzbench_lz4_b200_decompress"):
        f = getattr(L, name)
        f.restype = i64
        f.argtypes = [vp, sz, vp, sz, sz, sz, vp]
    L.lzbench_lz4_b200_compress_batch.restype = i64
    L.lzbench_lz4_b200_compress_batch.argtypes = [vp, sz, vp, vp, vp, sz, vp]
    L.lzbench_lz4_b200_decompress_batch.restype = i64
    L.lzbench_lz4_b200_decompress_batch.argtypes = [vp, vp, sz, vp, vp, vp]
    L.lz4b200_memcpy_batch.restype = i64
    L.lz4b200_memcpy_batch.argtypes = [vp, vp, sz, vp]
    L.lz4b200_frame_bound.restype = sz
    L.lz4b200_frame_bound.argtypes = [sz, C.c_int]
    L.lz4b200_frame_compress.restype = i64
    L.lz4b200_frame_compress.argtypes = [vp, vp, sz, vp, sz, C.c_int, C.c_int]
    L.lz4b200_frame_decompress.restype = i64
    L.lz4b200_frame_decompress.argtypes = [vp, vp, sz, vp, sz, C.c_int]
    L.lz4b200_frame_content_size.restype = i64
    L.lz4b200_frame_content_size.argtypes = [vp, sz]
    L.lz4b200_corpus_fill.restype = C.c_int
    L.lz4b200_corpus_fill.argtypes = [C.c_int, u64, u64, u64, vp, C.c_int]
    return L


lib = _load()


FRAME_CONTENT_CHECKSUM, FRAME_BLOCK_CHECKSUM, FRAME_CONTENT_SIZE = 1, 2, 4     # include/lz4b200.h
FRAME_SKIP_CHECKSUMS = 1


def last_error() -> str:
    return lib.lz4b200_last_error().decode()


def compress_bound(n: int) -> int:
    """lz4/lz4.h:171 LZ4_COMPRESSBOUND."""
    return int(lib.lz4b200_compress_bound(n))


def get_compress_bound(insize: int) -> int:
    """_lzbench/lzbench.h:17 GET_COMPRESS_BOUND (capacity lzbench gives the compressed image)."""
    return insize + insize // 6 + 16 * 1024


def chunk_list(file_sizes, chunk_size: int) -> np.ndarray:
    """The chunk_sizes vector lzbench_test builds (_lzbench/lzbench.cpp:366-373)."""
    if np.isscalar(file_sizes):
        file_sizes = [int(file_sizes)]
    out = []
    for fs in file_sizes:
        fs = int(fs)
        full, rem = divmod(fs, chunk_size)
        out.extend([chunk_size] * full)
        if rem:
            out.append(rem)
    return np.asarray(out, dtype=np.uint64)


def corpus(kind: str, nbytes: int, seed: int = 1, offset: int = 0, out: Optional[np.ndarray] = None,
           threads: int = 0) -> np.ndarray:
    """Deterministic synthetic corpus; see csrc/corpus.c."""
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size >= nbytes
    if threads <= 0:
        threads = min(os.cpu_count() or 1, 64)
    rc = lib.lz4b200_corpus_fill(CORPUS_KINDS[kind], seed, offset, nbytes, out.ctypes.data, threads)
    if rc != 0:
        raise ValueError("lz4b200_corpus_fill: bad arguments (offset must be a multiple of 64 KiB)")
    return out[:nbytes]


def _ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):      # torch tensor (host or device)
        return a.data_ptr()
    raise TypeError(type(a))


def _sizes(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


class Lz4B200:
    """One codec context on one GPU (lz4b200_create / lz4b200_destroy)."""

    def __init__(self, device: int = 0, max_total_bytes: int = 0, max_chunk_bytes: int = 0):
        h = C.c_void_p()
        rc = lib.lz4b200_create(device, max_total_bytes, max_chunk_bytes, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"lz4b200_create failed ({rc}): {last_error()}")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib.lz4b200_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launches(self) -> int:
        return int(lib.lz4b200_launch_count(self._h))

    def set_option(self, key: str, value: str):
        if lib.lz4b200_set_option(self._h, key.encode(), value.encode()) != 0:
            raise ValueError(last_error())

    def kernel_times_ms(self):
        buf = (C.c_float * 256)()
        n = lib.lz4b200_kernel_times_ms(self._h, buf, 256)
        if n < 0:
            raise RuntimeError(last_error())
        return [float(buf[i]) for i in range(n)]

    def debug_prof(self):
        buf = (C.c_uint64 * 16)()
        if lib.lz4b200_debug_prof(self._h, buf) != 0:
            raise RuntimeError(last_error())
        return [int(x) for x in buf]

    # -- host-buffer batched calls (the reference-facing hot path) -----------------------
    def compress_batch(self, inbuf, chunk_sizes, outbuf=None) -> Tuple[np.ndarray, np.ndarray, int]:
        cs = _sizes(chunk_sizes)
        total = int(cs.sum())
        if outbuf is None:
            outbuf = np.empty(get_compress_bound(total), dtype=np.uint8)
        compr = np.zeros(len(cs), dtype=np.uint64)
        cap = outbuf.nbytes if isinstance(outbuf, np.ndarray) else outbuf.numel()
        r = lib.lz4b200_compress_batch(self._h, _ptr(inbuf), cs.ctypes.data, len(cs), _ptr(outbuf), cap,
                                       compr.ctypes.data)
        if r < 0:
            raise RuntimeError(f"lz4b200_compress_batch failed ({r}): {last_error()}")
        return outbuf, compr, int(r)

    def decompress_batch(self, inbuf, compr_sizes, chunk_sizes, outbuf=None) -> Tuple[np.ndarray, int]:
        cs = _sizes(chunk_sizes)
        zs = _sizes(compr_sizes)
        assert len(cs) == len(zs)
        if outbuf is None:
            outbuf = np.zeros(int(cs.sum()), dtype=np.uint8)
        r = lib.lz4b200_decompress_batch(self._h, _ptr(inbuf), zs.ctypes.data, cs.ctypes.data, len(cs), _ptr(outbuf))
        return outbuf, int(r)

    def memcpy_batch(self, inbuf, n: int, outbuf) -> int:
        """PCIe yardstick: n bytes host -> device -> host, no kernel (the reference's cudaMemcpy row)."""
        return int(lib.lz4b200_memcpy_batch(self._h, _ptr(inbuf), n, _ptr(outbuf)))

    # -- LZ4 frame container (lz4/lz4frame.h LZ4F_compressFrame / LZ4F_decompress; independent blocks) ------
    def frame_compress(self, data, block_id: int = 4, content_checksum: bool = False, block_checksum: bool = False,
                       content_size: bool = False, outbuf=None) -> np.ndarray:
        n = data.nbytes if isinstance(data, np.ndarray) else data.numel()
        if outbuf is None:
            outbuf = np.empty(lib.lz4b200_frame_bound(n, block_id), dtype=np.uint8)
        flags = (FRAME_CONTENT_CHECKSUM if content_checksum else 0) | (FRAME_BLOCK_CHECKSUM if block_checksum else 0) \
            | (FRAME_CONTENT_SIZE if content_size else 0)
        r = lib.lz4b200_frame_compress(self._h, _ptr(data) if n else None, n, _ptr(outbuf), outbuf.nbytes, block_id, flags)
        if r < 0:
            raise RuntimeError(f"lz4b200_frame_compress failed ({r}): {last_error()}")
        return outbuf[:r]

    def frame_decompress(self, frame, capacity: int, skip_checksums: bool = False):
        """Returns (return code, output[:decoded size]); the code is negative for a rejected frame."""
        out = np.zeros(max(capacity, 1), dtype=np.uint8)
        r = lib.lz4b200_frame_decompress(self._h, _ptr(frame), frame.nbytes, _ptr(out), capacity,
                                         FRAME_SKIP_CHECKSUMS if skip_checksums else 0)
        return int(r), out[:max(int(r), 0)]

    # -- device-resident calls (torch CUDA tensors; asynchronous on `stream`) ---------------
    def compress_device(self, d_in, d_chunk_sizes, n_chunks: int, max_chunk: int, d_out, d_compr_sizes,
                        d_total, stream: int = 0):
        rc = lib.lz4b200_compress_device(self._h, _ptr(d_in), _ptr(d_chunk_sizes), n_chunks, max_chunk,
                                         _ptr(d_out), _ptr(d_compr_sizes), _ptr(d_total), stream)
        if rc != 0:
            raise RuntimeError(f"lz4b200_compress_device failed ({rc}): {last_error()}")

    def decompress_device(self, d_in, d_compr_sizes, d_chunk_sizes, n_chunks: int, max_chunk: int, d_out,
                          d_bad=None, stream: int = 0):
        rc = lib.lz4b200_decompress_device(self._h, _ptr(d_in), _ptr(d_compr_sizes), _ptr(d_chunk_sizes),
                                           n_chunks, max_chunk, _ptr(d_out), _ptr(d_bad), stream)
        if rc != 0:
            raise RuntimeError(f"lz4b200_decompress_device failed ({rc}): {last_error()}")


# ---- mirrors of the reference's chunk loops (same argument meaning, same return values) ------

def lzbench_compress(chunk_sizes, inbuf: np.ndarray, outbuf: np.ndarray, workmem: Lz4B200):
    """_lzbench/lzbench.cpp:266-298 with the codec fixed to lz4_b200: returns (sum, compr_sizes)."""
    _, compr, total = workmem.compress_batch(inbuf, chunk_sizes, outbuf)
    return total, compr


def lzbench_decompress(chunk_sizes, compr_sizes, inbuf: np.ndarray, outbuf: np.ndarray, workmem: Lz4B200) -> int:
    """_lzbench/lzbench.cpp:301-329 with the codec fixed to lz4_b200: returns sum or <= 0."""
    _, total = workmem.decompress_batch(inbuf, compr_sizes, chunk_sizes, outbuf)
    return total
