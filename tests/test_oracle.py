"""CPU tests: pin the oracle (C restatement) against the reference compiled here (oracle/_ref),
against the SURVEY section-4 accept/reject table and against the committed golden vectors."""
import json
import os

import numpy as np
import pytest

import lzbench_b200 as L
from tests import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(O.ref is None, reason="oracle/_ref not built (reference tree absent)")


def corpora(n):
    rng = np.random.default_rng(7)
    yield "mixed", L.corpus("mixed", n, seed=1)
    yield "random", L.corpus("random", n, seed=4)
    yield "repeat", L.corpus("repeat", n, seed=5)
    yield "zero", L.corpus("zero", n, seed=5)
    yield "lowent", rng.integers(0, 4, n, dtype=np.uint8)


@needs_ref
def test_ref_is_lz4_193():
    assert O.ref.LZ4_versionString() == b"1.9.3"


@needs_ref
@pytest.mark.parametrize("chunk", [1, 12, 13, 64, 4096, 65536, 65546, 65547, 200000, 1 << 20])
def test_oracle_compress_is_byte_identical_to_reference(chunk):
    for name, data in corpora(1 << 20):
        for off in range(0, min(len(data), 4 * chunk), chunk):
            part = data[off:off + chunk]
            a, b = O.oracle_compress(part), O.ref_compress(part)
            assert np.array_equal(a, b), (name, chunk, off)
            assert O.oracle_validate(a, len(part)) == 0
            r, out = O.oracle_decompress(a, len(part))
            assert r == len(part) and np.array_equal(out, part)
            r2, out2 = O.ref_decompress_safe(a, len(part))
            assert r2 == len(part) and np.array_equal(out2, part)


@needs_ref
def test_oracle_limited_output_matches_reference():
    data = L.corpus("mixed", 65536, seed=9)
    full = O.ref_compress(data)
    import ctypes as C
    for cap in [len(full) - 1, len(full), len(full) + 1, 100, 40000]:
        a = np.zeros(cap + 8, dtype=np.uint8); b = np.zeros(cap + 8, dtype=np.uint8)
        ra = O.orc.oracle_lz4_compress(data.ctypes.data, len(data), a.ctypes.data, cap)
        rb = O.ref.LZ4_compress_default(data.ctypes.data, b.ctypes.data, len(data), cap)
        assert ra == rb
        assert np.array_equal(a[:ra], b[:rb])


@pytest.mark.parametrize("kat", O.KATS, ids=[k[0] for k in O.KATS])
def test_kat_table(kat):
    name, blk, n, expect = kat
    r, out = O.oracle_decompress(blk, n)
    v = O.oracle_validate(blk, n)
    if expect is None:
        assert r < 0 or r != n, name
        assert v != 0
    else:
        assert r == n and out.tobytes() == expect
        assert v == 0
    if O.ref is not None and name != "offset-zero":      # documented deviation: reference accepts offset 0
        r2, out2 = O.ref_decompress_safe(blk, n)
        if expect is None:
            assert r2 < 0 or r2 != n
        else:
            assert r2 == n and out2.tobytes() == expect


@needs_ref
def test_corrupted_blocks_agree_with_reference():
    rng = np.random.default_rng(11)
    data = L.corpus("mixed", 8192, seed=3)
    good = O.oracle_compress(data)
    agree = 0
    for _ in range(2000):
        bad = good.copy()
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
        r, out = O.oracle_decompress(bad, len(data))
        r2, out2 = O.ref_decompress_safe(bad, len(data))
        if r == len(data):            # oracle accepts => reference accepts with the same bytes
            assert r2 == len(data) and np.array_equal(out, out2)
        if (r == len(data)) == (r2 == len(data)):
            agree += 1
    assert agree >= 1990              # the reference's shortcut path accepts a few streams the careful loop rejects


def test_chunk_loops_roundtrip():
    data = L.corpus("mixed", 300000, seed=2)
    for chunk in (4096, 65536, 100000):
        cs = L.chunk_list(len(data), chunk)
        packed, compr = O.oracle_chunked_compress(data, cs)
        assert int(compr.sum()) == len(packed)
        r, out = O.oracle_chunked_decompress(packed, compr, cs)
        assert r == len(data) and np.array_equal(out, data)


def test_raw_store_rule():
    # a chunk whose block would be exactly chunk-sized is stored raw (_lzbench/lzbench.cpp:284-289):
    # construct by brute force is impractical; check the decoder side: compr == chunk => raw copy
    data = np.arange(100, dtype=np.uint8)
    r, out = O.oracle_chunked_decompress(data, [100], [100])
    assert r == 100 and np.array_equal(out, data)


def test_golden_vectors():
    """Vectors generated from oracle/_ref by tests/golden/make_golden.py (committed)."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        g = json.load(f)
    for v in g["vectors"]:
        data = L.corpus(v["kind"], v["n"], seed=v["seed"])
        import hashlib
        assert hashlib.sha256(data.tobytes()).hexdigest() == v["plain_sha256"], "corpus generator drifted"
        blk = O.oracle_compress(data)
        assert len(blk) == v["clen"]
        assert hashlib.sha256(blk.tobytes()).hexdigest() == v["block_sha256"]
    for v in g["blocks"]:
        blk = np.frombuffer(bytes.fromhex(v["block_hex"]), dtype=np.uint8)
        r, out = O.oracle_decompress(blk, v["n"])
        assert r == v["n"]
        assert out.tobytes().hex() == v["plain_hex"]


def test_mt_driver_matches_serial():
    import ctypes as C
    data = L.corpus("mixed", 1 << 20, seed=6)
    cs = L.chunk_list(len(data), 65536)
    stride = 65536 + 65536 // 255 + 16
    packed = np.zeros(len(cs) * stride, dtype=np.uint8)
    compr = np.zeros(len(cs), dtype=np.uint64)
    ns = O.orc.oracle_mt_run(O.PLUGIN_COMPRESS, 0, 3, cs.ctypes.data, compr.ctypes.data, len(cs),
                             data.ctypes.data, packed.ctypes.data, stride)
    assert ns > 0
    _, compr_serial = O.oracle_chunked_compress(data, cs)
    assert np.array_equal(compr, compr_serial)
    out = np.zeros_like(data)
    ns = O.orc.oracle_mt_run(O.PLUGIN_DECOMPRESS, 1, 3, cs.ctypes.data, compr.ctypes.data, len(cs),
                             out.ctypes.data, packed.ctypes.data, stride)
    assert ns > 0 and np.array_equal(out, data)


@needs_ref
@pytest.mark.parametrize("accel", [0, 1, 2, 4, 8, 17, 64, 99])
def test_oracle_compress_fast_is_byte_identical_to_reference(accel):
    """the `lz4fast` levels (LZ4_compress_fast, _lzbench/compressors.cpp:348-356)."""
    for name, data in corpora(1 << 18):
        for chunk in (4096, 65536, 1 << 18):
            part = data[:chunk]
            cap = O.orc.oracle_lz4_compress_bound(len(part))
            a = np.zeros(cap, dtype=np.uint8); b = np.zeros(cap, dtype=np.uint8)
            ra = O.orc.oracle_lz4_compress_fast(part.ctypes.data, len(part), a.ctypes.data, cap, accel)
            rb = O.ref.LZ4_compress_fast(part.ctypes.data, b.ctypes.data, len(part), cap, accel)
            assert ra == rb and np.array_equal(a[:ra], b[:rb]), (name, chunk, accel)
