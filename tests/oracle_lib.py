"""ctypes access to the CHECKER: oracle/liboracle.so (C restatement) and, when present,
oracle/_ref/liblz4ref.so (the reference's own lz4 1.9.3 compiled by oracle/Makefile).
Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblz4ref.so")

vp = C.c_void_p


class SeqStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_seq", "lit_bytes", "match_bytes", "zero_lit_seq", "overlap_seq", "short_off_seq", "max_seq_in_block")]


orc = C.CDLL(ORACLE_SO)
orc.oracle_lz4_compress_bound.argtypes = [C.c_int]
orc.oracle_lz4_compress.argtypes = [vp, C.c_int, vp, C.c_int]
orc.oracle_lz4_decompress_safe.argtypes = [vp, C.c_int, vp, C.c_int]
orc.oracle_lz4_compress_fast.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
orc.oracle_lz4_validate.argtypes = [vp, C.c_int, C.c_int, C.POINTER(SeqStats)]
orc.oracle_lzbench_compress.restype = C.c_int64
orc.oracle_lzbench_compress.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, C.c_size_t]
orc.oracle_lzbench_decompress.restype = C.c_int64
orc.oracle_lzbench_decompress.argtypes = [vp, vp, vp, C.c_size_t, vp, vp]
orc.oracle_lzbench_compress_bound.restype = C.c_size_t
orc.oracle_lzbench_compress_bound.argtypes = [C.c_size_t]
orc.oracle_mt_run.restype = C.c_int64
orc.oracle_mt_run.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_size_t, vp, vp, C.c_size_t]

ref = None
if os.path.exists(REF_SO):
    ref = C.CDLL(REF_SO)
    ref.LZ4_compress_default.argtypes = [vp, vp, C.c_int, C.c_int]
    ref.LZ4_decompress_safe.argtypes = [vp, vp, C.c_int, C.c_int]
    ref.LZ4_compress_fast.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    ref.LZ4_decompress_fast.argtypes = [vp, vp, C.c_int]
    ref.LZ4_compressBound.argtypes = [C.c_int]
    ref.LZ4_versionString.restype = C.c_char_p

PLUGIN_COMPRESS = C.cast(orc.oracle_plugin_lz4_compress, vp)
PLUGIN_DECOMPRESS = C.cast(orc.oracle_plugin_lz4_decompress, vp)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def oracle_compress(data: np.ndarray) -> np.ndarray:
    data = _u8(data)
    out = np.empty(orc.oracle_lz4_compress_bound(len(data)), dtype=np.uint8)
    n = orc.oracle_lz4_compress(data.ctypes.data, len(data), out.ctypes.data, len(out))
    assert n > 0
    return out[:n].copy()


def oracle_decompress(block: np.ndarray, n: int):
    """Returns (return code, output array)."""
    block = _u8(block)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    src = block if len(block) else np.zeros(1, dtype=np.uint8)
    r = orc.oracle_lz4_decompress_safe(src.ctypes.data, len(block), out.ctypes.data, n)
    return r, out[:n]


def oracle_validate(block: np.ndarray, n: int, stats: SeqStats = None) -> int:
    block = _u8(block)
    src = block if len(block) else np.zeros(1, dtype=np.uint8)
    return orc.oracle_lz4_validate(src.ctypes.data, len(block), n, C.byref(stats) if stats is not None else None)


def ref_compress(data: np.ndarray) -> np.ndarray:
    data = _u8(data)
    out = np.empty(ref.LZ4_compressBound(len(data)), dtype=np.uint8)
    n = ref.LZ4_compress_default(data.ctypes.data, out.ctypes.data, len(data), len(out))
    assert n > 0
    return out[:n].copy()


def ref_decompress_safe(block: np.ndarray, n: int):
    block = _u8(block)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    src = block if len(block) else np.zeros(1, dtype=np.uint8)
    r = ref.LZ4_decompress_safe(src.ctypes.data, out.ctypes.data, len(block), n)
    return r, out[:n]


def oracle_chunked_compress(data: np.ndarray, chunk_sizes: np.ndarray):
    """lzbench_compress through the restated lz4 plugin: (packed bytes, compr_sizes)."""
    data = _u8(data)
    cs = np.ascontiguousarray(chunk_sizes, dtype=np.uint64)
    cap = int(orc.oracle_lzbench_compress_bound(len(data)))
    out = np.empty(cap, dtype=np.uint8)
    compr = np.zeros(len(cs), dtype=np.uint64)
    total = orc.oracle_lzbench_compress(PLUGIN_COMPRESS, cs.ctypes.data, len(cs), compr.ctypes.data,
                                        data.ctypes.data, out.ctypes.data, cap)
    assert total > 0
    return out[:total].copy(), compr


def oracle_chunked_decompress(packed: np.ndarray, compr_sizes, chunk_sizes):
    packed = _u8(packed)
    cs = np.ascontiguousarray(chunk_sizes, dtype=np.uint64)
    zs = np.ascontiguousarray(compr_sizes, dtype=np.uint64)
    out = np.zeros(int(cs.sum()), dtype=np.uint8)
    r = orc.oracle_lzbench_decompress(PLUGIN_DECOMPRESS, cs.ctypes.data, zs.ctypes.data, len(cs),
                                      packed.ctypes.data, out.ctypes.data)
    return int(r), out


def split_chunks(packed: np.ndarray, compr_sizes):
    off = 0
    for z in compr_sizes:
        z = int(z)
        yield packed[off:off + z]
        off += z


def block(*parts) -> np.ndarray:
    """Hand-build a block from ints / bytes pieces."""
    b = bytearray()
    for p in parts:
        if isinstance(p, int):
            b.append(p)
        else:
            b.extend(p)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


# The SURVEY.md section-4 accept/reject table (+ a few more), as (name, block, n, expected output or None).
KATS = [
    ("all-literal-13", block(0xD0, b"abcdefghijklm"), 13, b"abcdefghijklm"),
    ("empty", block(0x00), 0, b""),
    ("lit4-match12-last5", block(0x48, b"abcd", 4, 0, 0x50, b"vwxyz"), 21, b"abcdabcdabcdabcdvwxyz"),
    ("last4-violates-lastliterals", block(0x48, b"abcd", 4, 0, 0x40, b"wxyz"), 20, None),
    ("match-in-last-12", block(0x10, b"a", 1, 0, 0x70, b"bcdefgh"), 13, None),
    ("rle-offset1", block(0x1F, b"a", 1, 0, 5, 0x50, b"vwxyz"), 30, b"a" * 25 + b"vwxyz"),
    ("period3", block(0x38, b"abc", 3, 0, 0x50, b"vwxyz"), 20, b"abcabcabcabcabcvwxyz"),
    ("offset-beyond-start", block(0x48, b"abcd", 5, 0, 0x50, b"vwxyz"), 21, None),
    ("offset-zero", block(0x48, b"abcd", 0, 0, 0x50, b"vwxyz"), 21, None),
    ("truncated", block(0x48, b"abcd", 4), 21, None),
    ("literal-run-270", block(0xF0, 255, 0, bytes(range(256)) + bytes(range(14))), 270, bytes(range(256)) + bytes(range(14))),
    ("wrong-size", block(0xD0, b"abcdefghijklm"), 15, None),
]


# ---- LZ4 frame container: oracle/lz4f_oracle.c and oracle/_ref/liblz4fref.so -------------------
REF_FRAME_SO = os.path.join(ROOT, "oracle", "_ref", "liblz4fref.so")
orc.oracle_xxh32.restype = C.c_uint32
orc.oracle_xxh32.argtypes = [vp, C.c_size_t, C.c_uint32]
orc.oracle_lz4f_bound.restype = C.c_size_t
orc.oracle_lz4f_bound.argtypes = [C.c_size_t, C.c_int]
orc.oracle_lz4f_compress.restype = C.c_size_t
orc.oracle_lz4f_compress.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.c_int, C.c_int, C.c_int]
orc.oracle_lz4f_decompress.restype = C.c_size_t
orc.oracle_lz4f_decompress.argtypes = [vp, C.c_size_t, vp, C.c_size_t]

reff = None
if os.path.exists(REF_FRAME_SO):
    reff = C.CDLL(REF_FRAME_SO)
    reff.ref_lz4f_bound.restype = C.c_size_t
    reff.ref_lz4f_bound.argtypes = [C.c_size_t, C.c_int, C.c_int, C.c_int]
    reff.ref_lz4f_compress.restype = C.c_size_t
    reff.ref_lz4f_compress.argtypes = [vp, C.c_size_t, vp, C.c_size_t] + [C.c_int] * 5
    reff.ref_lz4f_decompress.restype = C.c_size_t
    reff.ref_lz4f_decompress.argtypes = [vp, C.c_size_t, vp, C.c_size_t]
    reff.XXH32.restype = C.c_uint32
    reff.XXH32.argtypes = [vp, C.c_size_t, C.c_uint32]

BAD = C.c_size_t(-1).value


def _ptr(a):
    return a.ctypes.data if len(a) else np.zeros(1, dtype=np.uint8).ctypes.data


def oracle_frame_compress(data, block_id=4, content_checksum=0, with_content_size=0) -> np.ndarray:
    data = _u8(data)
    out = np.empty(orc.oracle_lz4f_bound(len(data), block_id), dtype=np.uint8)
    n = orc.oracle_lz4f_compress(_ptr(data), len(data), out.ctypes.data, len(out), block_id, content_checksum,
                                 with_content_size)
    assert n > 0
    return out[:n].copy()


def oracle_frame_decompress(frame, cap):
    frame = _u8(frame)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    r = orc.oracle_lz4f_decompress(_ptr(frame), len(frame), out.ctypes.data, cap)
    return (r, out[:r]) if r != BAD else (BAD, out[:0])


def ref_frame_compress(data, block_id=4, independent=1, content_checksum=0, block_checksum=0,
                       with_content_size=0) -> np.ndarray:
    data = _u8(data)
    out = np.empty(reff.ref_lz4f_bound(len(data), block_id, content_checksum, block_checksum) + 64, dtype=np.uint8)
    n = reff.ref_lz4f_compress(_ptr(data), len(data), out.ctypes.data, len(out), block_id, independent,
                               content_checksum, block_checksum, with_content_size)
    assert n > 0
    return out[:n].copy()


def ref_frame_decompress(frame, cap):
    frame = _u8(frame)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    r = reff.ref_lz4f_decompress(_ptr(frame), len(frame), out.ctypes.data, cap)
    return (r, out[:r]) if r != BAD else (BAD, out[:0])
