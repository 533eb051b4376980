"""GPU parity tests (run on the B200 box): every call goes through the C ABI of
lzbench_b200/_lib/liblz4b200.so; the oracle (and oracle/_ref when it travelled) is the checker.

Bars: decompression bit-exact; every GPU-compressed chunk decodes bit-exactly through the oracle
decoder AND the reference's LZ4_decompress_safe / LZ4_decompress_fast; every oracle/reference
compressed chunk decodes bit-exactly on the GPU; compressed size <= 1.03 x reference on the same
chunking (BASELINE.json north_star)."""
import ctypes as C
import zlib

import numpy as np
import pytest

import lzbench_b200 as L
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu

RATIO_TOLERANCE = 1.03

CORPORA = [("mixed", 1), ("random", 4), ("repeat", 5), ("zero", 5)]


@pytest.fixture(autouse=True)
def _default_kernels(gpu_ctx):
    """The context is shared by the session: every test starts from the default kernel selection."""
    gpu_ctx.set_option("kernels", "default")
    gpu_ctx.set_option("team", "0")
    yield


def check_gpu_blocks_with_reference(data, packed, compr, cs):
    off_p = off_u = 0
    for z, n in zip(compr.tolist(), cs.tolist()):
        z, n = int(z), int(n)
        blk, plain = packed[off_p:off_p + z], data[off_u:off_u + n]
        if z == n:                      # raw-stored chunk
            assert np.array_equal(blk, plain)
        else:
            assert O.oracle_validate(blk, n) == 0, (off_u, z, n)
            r, out = O.oracle_decompress(blk, n)
            assert r == n and np.array_equal(out, plain), ("oracle decode", off_u)
            if O.ref is not None:
                r, out = O.ref_decompress_safe(blk, n)
                assert r == n and np.array_equal(out, plain), ("LZ4_decompress_safe", off_u)
                out = np.zeros(n + 64, dtype=np.uint8)
                r = O.ref.LZ4_decompress_fast(blk.ctypes.data, out.ctypes.data, n)
                assert r == z and np.array_equal(out[:n], plain), ("LZ4_decompress_fast", off_u)
        off_p += z
        off_u += n


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
@pytest.mark.parametrize("chunk", [4096, 65536, 100000, 1 << 20])
@pytest.mark.parametrize("kind,seed", CORPORA)
def test_compress_cross_decodes_and_ratio(gpu_ctx, kernels, chunk, kind, seed):
    gpu_ctx.set_option("kernels", kernels)
    n = 2 * (1 << 20) + 777                      # ragged last chunk
    data = L.corpus(kind, n, seed=seed)
    cs = L.chunk_list(n, chunk)
    packed, compr, total = gpu_ctx.compress_batch(data, cs)
    assert total > 0 and int(compr.sum()) == total
    check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
    _, ref_compr = O.oracle_chunked_compress(data, cs)
    assert total <= RATIO_TOLERANCE * int(ref_compr.sum()) + 64, (total, int(ref_compr.sum()))


@pytest.mark.parametrize("chunk", [4096, 16384, 65536, 100000])
@pytest.mark.parametrize("kind,seed", CORPORA)
def test_compress_c3_cross_decodes_and_ratio(gpu_ctx, chunk, kind, seed):
    """The generation-3 compressor (kernels_c3.cuh, option compress=v3; chunks above 64 KiB stay on c2's segments): same bars."""
    gpu_ctx.set_option("compress", "v3")
    n = 2 * (1 << 20) + 777
    data = L.corpus(kind, n, seed=seed)
    cs = L.chunk_list(n, chunk)
    packed, compr, total = gpu_ctx.compress_batch(data, cs)
    assert total > 0 and int(compr.sum()) == total
    check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
    _, ref_compr = O.oracle_chunked_compress(data, cs)
    assert total <= RATIO_TOLERANCE * int(ref_compr.sum()) + 64, (total, int(ref_compr.sum()))
    out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
    assert r == n and np.array_equal(out, data)


def test_compress_c3_small_ragged_and_misaligned_chunks(gpu_ctx):
    """c3 on chunk lists whose chunks start at any byte alignment (the realigning load) and have any size from 1 byte up."""
    gpu_ctx.set_option("compress", "v3")
    rng = np.random.default_rng(7)
    sizes = [1, 2, 5, 11, 12, 13, 14, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 255, 256, 270, 271, 1000, 2047, 2048, 2049,
             2059, 2060, 2061, 4095, 4097, 6143, 6156, 65535, 65536, 65533, 40000]
    sizes += [int(x) for x in rng.integers(1, 9000, 300)]
    cs = np.asarray(sizes, dtype=np.uint64)
    n = int(cs.sum())
    for kind, seed in CORPORA:
        data = L.corpus(kind, n, seed=seed)
        packed, compr, total = gpu_ctx.compress_batch(data, cs)
        check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
        out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
        assert r == n and np.array_equal(out, data)
        _, ref_compr = O.oracle_chunked_compress(data, cs)
        assert total <= RATIO_TOLERANCE * int(ref_compr.sum()) + 8 * len(cs), (kind, total, int(ref_compr.sum()))


def test_compress_c3_structured_data(gpu_ctx):
    """Data that stresses the run logic of c3: short periods (every position its own run), long matches that start in one
    tile and end many tiles later, matches at the maximum offset, a chunk that is one match."""
    gpu_ctx.set_option("compress", "v3")
    rng = np.random.default_rng(11)
    parts = []
    for period in (1, 2, 3, 5, 7, 12, 31, 33, 63, 64, 65, 127, 255, 256, 257, 1000, 2047, 2048, 2049, 5000):
        pat = rng.integers(0, 256, period, dtype=np.uint8)
        parts.append(np.tile(pat, 65536 // period + 1)[:65536])
    blk = rng.integers(0, 256, 30000, dtype=np.uint8)                    # a long block seen twice, 65 535 bytes apart at most
    far = np.concatenate([blk, rng.integers(0, 256, 65536 - 2 * 30000, dtype=np.uint8), blk])
    parts.append(far[:65536])
    words = [bytes(rng.integers(97, 123, int(k), dtype=np.uint8)) for k in rng.integers(2, 9, 50)]
    text = b" ".join(words[int(i)] for i in rng.integers(0, 50, 20000))
    parts.append(np.frombuffer(text[:65536], dtype=np.uint8))
    data = np.concatenate(parts)
    cs = L.chunk_list(len(data), 65536)
    packed, compr, total = gpu_ctx.compress_batch(data, cs)
    check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
    _, ref_compr = O.oracle_chunked_compress(data, cs)
    assert total <= RATIO_TOLERANCE * int(ref_compr.sum()) + 64 * len(cs), (total, int(ref_compr.sum()))
    out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
    assert r == len(data) and np.array_equal(out, data)


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
@pytest.mark.parametrize("chunk", [4096, 65536, 100000, 1 << 20])
@pytest.mark.parametrize("kind,seed", CORPORA)
def test_decompress_reference_blocks_bit_exact(gpu_ctx, kernels, chunk, kind, seed):
    gpu_ctx.set_option("kernels", kernels)
    n = 2 * (1 << 20) + 777
    data = L.corpus(kind, n, seed=seed)
    cs = L.chunk_list(n, chunk)
    packed, compr = O.oracle_chunked_compress(data, cs)
    out, r = gpu_ctx.decompress_batch(packed, compr, cs)
    assert r == n
    assert np.array_equal(out, data)


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
def test_small_and_ragged_chunks(gpu_ctx, kernels):
    gpu_ctx.set_option("kernels", kernels)
    rng = np.random.default_rng(5)
    sizes = [1, 2, 5, 11, 12, 13, 14, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 255, 256, 270, 271, 1000,
             4095, 4097, 65535, 65536, 65537, 65546, 65547, 70000]
    sizes += [int(x) for x in rng.integers(1, 3000, 200)]
    cs = np.asarray(sizes, dtype=np.uint64)
    n = int(cs.sum())
    for kind, seed in CORPORA:
        data = L.corpus(kind, n, seed=seed)
        packed, compr, total = gpu_ctx.compress_batch(data, cs)
        check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
        out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
        assert r == n and np.array_equal(out, data)
        p2, c2 = O.oracle_chunked_compress(data, cs)
        out, r = gpu_ctx.decompress_batch(p2, c2, cs)
        assert r == n and np.array_equal(out, data)


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
def test_raw_stored_chunks(gpu_ctx, kernels):
    """compr_sizes[i] == chunk_sizes[i] means raw bytes (_lzbench/lzbench.cpp:311-315)."""
    gpu_ctx.set_option("kernels", kernels)
    data = L.corpus("mixed", 3 * 65536 + 100, seed=8)
    cs = L.chunk_list(len(data), 65536)
    packed, compr = O.oracle_chunked_compress(data, cs)
    # replace chunk 1 by its raw bytes
    parts = list(O.split_chunks(packed, compr))
    parts[1] = data[65536:131072]
    compr = compr.copy(); compr[1] = 65536
    packed = np.concatenate(parts)
    out, r = gpu_ctx.decompress_batch(packed, compr, cs)
    assert r == len(data) and np.array_equal(out, data)


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
def test_kat_table_on_gpu(gpu_ctx, kernels):
    gpu_ctx.set_option("kernels", kernels)
    for name, blk, n, expect in O.KATS:
        if n == 0:
            continue                      # lzbench never forms empty chunks; covered by the oracle test
        out, r = gpu_ctx.decompress_batch(blk, [len(blk)], [n])
        if expect is None:
            assert r <= 0, name
        else:
            assert r == n, name
            assert out.tobytes() == expect, name


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
def test_corrupted_blocks_never_crash_and_match_oracle(gpu_ctx, kernels):
    gpu_ctx.set_option("kernels", kernels)
    rng = np.random.default_rng(13)
    data = L.corpus("mixed", 8192, seed=3)
    good = O.oracle_compress(data)
    blocks, expect = [], []
    for _ in range(512):
        bad = good.copy()
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
        blocks.append(bad)
        r, out = O.oracle_decompress(bad, len(data))
        expect.append((r == len(data), out))
    # one batch with all of them: count of failures must match; then each alone for exactness
    packed = np.concatenate(blocks)
    compr = np.full(len(blocks), len(good), dtype=np.uint64)
    cs = np.full(len(blocks), len(data), dtype=np.uint64)
    out, r = gpu_ctx.decompress_batch(packed, compr, cs)
    n_bad = sum(1 for ok, _ in expect if not ok)
    assert (r == -n_bad) if n_bad else (r == len(data) * len(blocks))
    out = out.reshape(len(blocks), len(data))
    for i, (ok, ref_out) in enumerate(expect):
        if ok:
            assert np.array_equal(out[i], ref_out), i


def _coarse_corpus(rng):
    """Large chunks that are a handful of long copies -- what kernels_coarse.cuh takes -- beside chunks it must leave alone."""
    n = 1 << 18
    rnd = rng.integers(0, 256, n, dtype=np.uint8)
    zero = np.zeros(n, dtype=np.uint8)
    per = np.tile(rng.integers(0, 256, 300, dtype=np.uint8), n // 300 + 1)[:n]
    half = np.concatenate([rnd[:n // 2], zero[:n // 2]])                 # one long literal run, then one long match
    back = np.concatenate([rnd[:100000], rnd[40000:100000], np.zeros(n - 160000, dtype=np.uint8)])   # a far match of 60 000 bytes
    text = L.corpus("mixed", n, seed=21)                                 # thousands of sequences: not for this path
    return [rnd, zero, per, half, back, text]


@pytest.mark.parametrize("coarse", ["1", "0"])
def test_coarse_chunks_decode_bit_exact(gpu_ctx, coarse):
    """Chunks above 64 KiB made of few long copies (incompressible, run-length, short-period, raw-stored), compressed by the
    reference (one match per run) and by this library (one match per 48 KiB segment): bit-exact with and without the coarse path."""
    gpu_ctx.set_option("coarse", coarse)
    try:
        rng = np.random.default_rng(31)
        parts = _coarse_corpus(rng)
        data = np.concatenate(parts)
        cs = np.asarray([len(p) for p in parts], dtype=np.uint64)
        for packed, compr in (O.oracle_chunked_compress(data, cs), gpu_ctx.compress_batch(data, cs)[:2]):
            total = int(np.asarray(compr).sum())
            out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
            assert r == len(data) and np.array_equal(out, data)
        # one of them stored raw (compr == size), and a ragged list around the 64 KiB limit of the path
        packed, compr = O.oracle_chunked_compress(data, cs)
        blocks = list(O.split_chunks(packed, compr))
        blocks[1] = parts[1]; compr = compr.copy(); compr[1] = len(parts[1])
        out, r = gpu_ctx.decompress_batch(np.concatenate(blocks), compr, cs)
        assert r == len(data) and np.array_equal(out, data)
        sizes = np.asarray([65535, 65536, 65537, 70000, 200000, 13, 1 << 20, 4 << 20], dtype=np.uint64)
        for kind, seed in (("random", 4), ("zero", 5), ("repeat", 5)):
            d2 = L.corpus(kind, int(sizes.sum()), seed=seed)
            for packed, compr in (O.oracle_chunked_compress(d2, sizes), gpu_ctx.compress_batch(d2, sizes)[:2]):
                total = int(np.asarray(compr).sum())
                out, r = gpu_ctx.decompress_batch(packed[:total], compr, sizes)
                assert r == len(d2) and np.array_equal(out, d2), kind
    finally:
        gpu_ctx.set_option("coarse", "1")


def _build_block(seqs, last):
    """An LZ4 block from (literals, offset, match length) triples and the last literals."""
    def ext(v):
        b = bytearray()
        while v >= 255:
            b.append(255); v -= 255
        b.append(v)
        return bytes(b)
    out = bytearray()
    for lit, off, ml in seqs:
        out.append((min(len(lit), 15) << 4) | min(ml - 4, 15))
        if len(lit) >= 15: out += ext(len(lit) - 15)
        out += lit
        out += bytes([off & 255, off >> 8])
        if ml - 4 >= 15: out += ext(ml - 4 - 15)
    out.append(min(len(last), 15) << 4)
    if len(last) >= 15: out += ext(len(last) - 15)
    out += last
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()


def test_coarse_chunks_chained_matches(gpu_ctx):
    """Hand-built blocks of long matches that read each other's output in every way the coarse path distinguishes: offsets that
    the earlier match's period divides (forwarded: same round), offsets it does not divide (next round), sources that straddle
    two matches, sources in literals, overlapping and non-overlapping matches; several chunks share the grid."""
    rng = np.random.default_rng(41)
    def lit(n): return bytes(rng.integers(0, 256, n, dtype=np.uint8))
    blocks = [
        [(lit(7), 7, 100000), (b"", 14, 80000), (b"", 10, 70000), (b"", 65000, 50000), (lit(3), 3, 90000), (b"", 60001, 66000)],
        [(lit(1), 1, 300000), (b"", 255, 200000), (lit(300), 300, 65536), (b"", 600, 70000), (b"", 450, 80000)],
        [(lit(70000), 65535, 65535), (b"", 65535, 131070), (lit(5), 5, 4), (b"", 9, 100000), (b"", 18, 100000), (b"", 27, 100000)],
        [(lit(64), 64, 49088), (b"", 256, 49152), (b"", 192, 49152), (b"", 224, 49152), (b"", 160, 49152), (b"", 32, 49152)],
    ]
    packed, compr, cs, plain = [], [], [], []
    for seqs in blocks:
        blk = _build_block(seqs, lit(12))
        n = sum(len(l) + m for l, _, m in seqs) + 12
        r, out = O.oracle_decompress(blk, n)
        assert r == n, (r, n)
        if O.ref is not None:
            r2, out2 = O.ref_decompress_safe(blk, n)
            assert r2 == n and np.array_equal(out, out2)
        packed.append(blk); compr.append(len(blk)); cs.append(n); plain.append(out)
    for reps in (1, 3):                                  # few chunks (many CTAs per chunk) and some more
        pk = np.concatenate(packed * reps)
        out, r = gpu_ctx.decompress_batch(pk, np.asarray(compr * reps, dtype=np.uint64), np.asarray(cs * reps, dtype=np.uint64))
        assert r == sum(cs) * reps
        assert np.array_equal(out, np.concatenate(plain * reps))


def test_coarse_chunks_corrupted_same_verdict_as_oracle(gpu_ctx):
    """Damaged blocks of the kind the coarse path takes: the same accept / reject decision as the oracle, and the oracle's
    bytes where it accepts (the coarse scan only ever accepts what passes every rule; the rest goes to the team decoder)."""
    rng = np.random.default_rng(37)
    n = 1 << 17
    sources = [rng.integers(0, 256, n, dtype=np.uint8), np.zeros(n, dtype=np.uint8),
               np.concatenate([rng.integers(0, 256, n // 2, dtype=np.uint8), np.zeros(n // 2, dtype=np.uint8)])]
    for data in sources:
        good = O.oracle_compress(data)
        blocks, expect = [], []
        for i in range(96):
            bad = good.copy()
            for _ in range(int(rng.integers(1, 3))):
                # the first bytes (token, length bytes, offset) and the last ones decide most; the rest of a literal run does not
                where = int(rng.integers(0, min(600, len(bad)))) if i % 3 else int(rng.integers(max(0, len(bad) - 600), len(bad)))
                bad[where] = int(rng.integers(0, 256))
            blocks.append(bad)
            r, out = O.oracle_decompress(bad, n)
            expect.append((r == n, out))
        packed = np.concatenate(blocks)
        compr = np.full(len(blocks), len(good), dtype=np.uint64)
        cs = np.full(len(blocks), n, dtype=np.uint64)
        out, r = gpu_ctx.decompress_batch(packed, compr, cs)
        n_bad = sum(1 for ok, _ in expect if not ok)
        assert (r == -n_bad) if n_bad else (r == n * len(blocks))
        out = out.reshape(len(blocks), n)
        for i, (ok, ref_out) in enumerate(expect):
            if ok:
                assert np.array_equal(out[i], ref_out), i


@pytest.mark.parametrize("v7big", ["1", "0"])
def test_large_compressible_chunks_and_their_corruptions(gpu_ctx, v7big):
    """Chunks above 64 KiB of compressible data go through the sequence table too (kernels_v7.cuh, BIG: 16-bit position fields
    whose wrap-arounds pass 2 counts, long matches in pieces): bit-exact on reference blocks of several sizes, and the oracle's
    verdict (and bytes) on damaged ones."""
    gpu_ctx.set_option("v7big", v7big)
    try:
        rng = np.random.default_rng(43)
        sizes = np.asarray([65537, 70000, 131072, 200001, 1 << 20, 3 << 20], dtype=np.uint64)
        data = L.corpus("mixed", int(sizes.sum()), seed=9)
        long_run = np.zeros(400000, dtype=np.uint8); long_run[::1000] = 7         # matches far above 32 768 bytes between text
        data[300000:700000] = long_run
        packed, compr = O.oracle_chunked_compress(data, sizes)
        out, r = gpu_ctx.decompress_batch(packed, compr, sizes)
        assert r == len(data) and np.array_equal(out, data)
        p2, c2, t2 = gpu_ctx.compress_batch(data, sizes)
        out, r = gpu_ctx.decompress_batch(p2[:t2], c2, sizes)
        assert r == len(data) and np.array_equal(out, data)
        n = 150000
        plain = data[:n]
        good = O.oracle_compress(plain)
        blocks, expect = [], []
        for _ in range(64):
            bad = good.copy()
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
            blocks.append(bad)
            rr, oo = O.oracle_decompress(bad, n)
            expect.append((rr == n, oo))
        pk = np.concatenate(blocks)
        out, r = gpu_ctx.decompress_batch(pk, np.full(len(blocks), len(good), dtype=np.uint64), np.full(len(blocks), n, dtype=np.uint64))
        n_bad = sum(1 for ok, _ in expect if not ok)
        assert (r == -n_bad) if n_bad else (r == n * len(blocks))
        out = out.reshape(len(blocks), n)
        for i, (ok, oo) in enumerate(expect):
            if ok:
                assert np.array_equal(out[i], oo), i
    finally:
        gpu_ctx.set_option("v7big", "1")


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
def test_plugin_signature_wrappers(gpu_ctx, kernels):
    """lzbench_lz4_b200_{init,compress,decompress,deinit}: the unpatched-harness path, one chunk per call."""
    wm = L.lib.lzbench_lz4_b200_init(65536, 0, 0)
    assert wm
    try:
        data = L.corpus("mixed", 200000, seed=4)
        cs = L.chunk_list(len(data), 65536)
        cap = L.get_compress_bound(65536)
        out = np.zeros(cap, dtype=np.uint8)
        back = np.zeros(65536, dtype=np.uint8)
        off = 0
        for n in cs.tolist():
            n = int(n)
            part = np.ascontiguousarray(data[off:off + n])
            clen = L.lib.lzbench_lz4_b200_compress(part.ctypes.data, n, out.ctypes.data, cap, 0, 0, wm)
            assert 0 < clen <= L.compress_bound(n)
            assert O.oracle_validate(out[:clen], n) == 0
            dlen = L.lib.lzbench_lz4_b200_decompress(out.ctypes.data, clen, back.ctypes.data, n, 0, 0, wm)
            assert dlen == n and np.array_equal(back[:n], part)
            off += n
    finally:
        L.lib.lzbench_lz4_b200_deinit(wm)


def test_output_capacity_rule(gpu_ctx):
    """lzbench_compress returns 0 when the image does not fit (_lzbench/lzbench.cpp:286)."""
    data = L.corpus("random", 1 << 20, seed=4)
    cs = L.chunk_list(len(data), 65536)
    small = np.zeros(len(data), dtype=np.uint8)        # random data expands
    _, _, total = gpu_ctx.compress_batch(data, cs, small)
    assert total == 0


@pytest.mark.parametrize("kernels", ["v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"])
def test_device_resident_roundtrip_256m(gpu_ctx, kernels):
    """Full-size property test: encode -> decode on the device, compare CRC of the whole image."""
    import torch
    gpu_ctx.set_option("kernels", kernels)
    n = 256 << 20
    chunk = 65536
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    L.corpus("mixed", n, seed=2, out=host.numpy())
    d_in = host.cuda()
    n_chunks = n // chunk
    d_cs = torch.full((n_chunks,), chunk, dtype=torch.int64, device="cuda")
    d_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8, device="cuda")
    d_compr = torch.empty(n_chunks, dtype=torch.int64, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    gpu_ctx.compress_device(d_in, d_cs, n_chunks, chunk, d_out, d_compr, d_total, s)
    torch.cuda.synchronize()
    total = int(d_total.item())
    assert int(d_compr.sum().item()) == total
    assert 0.4 * n < total < 0.75 * n
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    gpu_ctx.decompress_device(d_out, d_compr, d_cs, n_chunks, chunk, d_back, d_bad, s)
    torch.cuda.synchronize()
    assert int(d_bad.item()) == 0
    assert torch.equal(d_back, d_in)
    # spot-check 64 chunks of the GPU image through the reference decoder
    compr = d_compr.cpu().numpy().astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(compr)[:-1]]).astype(np.int64)
    packed = d_out[:total].cpu().numpy()
    plain = host.numpy()
    for i in np.linspace(0, n_chunks - 1, 64).astype(int):
        blk = packed[offs[i]:offs[i] + int(compr[i])]
        r, out = O.oracle_decompress(blk, chunk)
        assert r == chunk and zlib.crc32(out.tobytes()) == zlib.crc32(plain[i * chunk:(i + 1) * chunk].tobytes())


def test_device_resident_roundtrip_5g_64bit_offsets(gpu_ctx):
    """More than 4 GiB in one batch: every offset array is 64-bit (BASELINE configs[3] puts 8 GiB on the box).
    Property check at full size: encode -> decode on the device, whole image compared on the device."""
    import torch
    gpu_ctx.set_option("kernels", "v4")
    n = 5 << 30
    chunk = 65536
    host = torch.empty(n, dtype=torch.uint8)
    L.corpus("mixed", n, seed=3, out=host.numpy())
    d_in = host.cuda()
    del host
    n_chunks = n // chunk
    d_cs = torch.full((n_chunks,), chunk, dtype=torch.int64, device="cuda")
    d_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8, device="cuda")
    d_compr = torch.empty(n_chunks, dtype=torch.int64, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    gpu_ctx.compress_device(d_in, d_cs, n_chunks, chunk, d_out, d_compr, d_total, s)
    torch.cuda.synchronize()
    total = int(d_total.item())
    assert total > (4 << 30) * 0.55 and int(d_compr.sum().item()) == total
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    gpu_ctx.decompress_device(d_out, d_compr, d_cs, n_chunks, chunk, d_back, d_bad, s)
    torch.cuda.synchronize()
    assert int(d_bad.item()) == 0
    assert torch.equal(d_back, d_in)
    # the last chunks lie beyond the 4 GiB mark of both images: check a few against the oracle too
    compr = d_compr.cpu().numpy().astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(compr)[:-1]]).astype(np.int64)
    for i in (n_chunks - 1, n_chunks - 7, 70000):
        blk = d_out[int(offs[i]):int(offs[i]) + int(compr[i])].cpu().numpy()
        plain = d_in[i * chunk:(i + 1) * chunk].cpu().numpy()
        r, out = O.oracle_decompress(blk, chunk)
        assert r == chunk and np.array_equal(out, plain)


def test_many_small_files_mode(gpu_ctx):
    """lzbench -j joins many small files and cuts every FILE into chunks, so the chunk list is
    ragged (_lzbench/lzbench.cpp:366-373, :537-620): thousands of files from 1 byte up, chunked at 4 KiB and 64 KiB."""
    rng = np.random.default_rng(11)
    files = np.concatenate([rng.integers(1, 40, 2000), rng.integers(40, 5000, 6000), rng.integers(5000, 300000, 150)])
    rng.shuffle(files)
    n = int(files.sum())
    data = L.corpus("mixed", n, seed=6)
    for chunk in (4096, 65536):
        cs = L.chunk_list(files, chunk)
        assert int(cs.sum()) == n and int(cs.max()) <= chunk and len(cs) >= len(files)
        packed, compr, total = gpu_ctx.compress_batch(data, cs)
        out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
        assert r == n and np.array_equal(out, data)
        p2, c2 = O.oracle_chunked_compress(data, cs)
        assert total <= RATIO_TOLERANCE * len(p2)
        out, r = gpu_ctx.decompress_batch(p2, c2, cs)
        assert r == n and np.array_equal(out, data)
        # spot-check blocks with the reference decoder (every 37th chunk)
        off_p = np.concatenate([[0], np.cumsum(compr)]).astype(np.int64)
        off_u = np.concatenate([[0], np.cumsum(cs)]).astype(np.int64)
        for i in range(0, len(cs), 37):
            z, u = int(compr[i]), int(cs[i])
            blk, plain = packed[off_p[i]:off_p[i] + z], data[off_u[i]:off_u[i] + u]
            if z == u:
                assert np.array_equal(blk, plain)
            else:
                rr, o = O.oracle_decompress(blk, u)
                assert rr == u and np.array_equal(o, plain)



@pytest.mark.parametrize("accel", [2, 4, 8, 33])
def test_lz4fast_acceleration_levels(gpu_ctx, accel):
    """`accel` = LZ4_compress_fast's acceleration (the lz4fast rows, _lzbench/compressors.cpp:348-356).
    Same bars as level 1, measured against the reference at the same acceleration."""
    n = 4 << 20
    try:
        for kind, seed in CORPORA:
            data = L.corpus(kind, n, seed=seed)
            for chunk in (4096, 65536, 1 << 20):
                cs = L.chunk_list(n, chunk)
                gpu_ctx.set_option("accel", str(accel))
                packed, compr, total = gpu_ctx.compress_batch(data, cs)
                out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
                assert r == n and np.array_equal(out, data)
                check_gpu_blocks_with_reference(data[:4 * chunk], packed, compr[:4], cs[:4])
                ref_total = 0
                tmp = np.empty(L.compress_bound(chunk), dtype=np.uint8)
                for i, u in enumerate(cs.tolist()):
                    part = data[i * chunk:i * chunk + int(u)]
                    z = O.orc.oracle_lz4_compress_fast(part.ctypes.data, len(part), tmp.ctypes.data, len(tmp), accel)
                    ref_total += min(z, len(part))                   # lzbench stores a chunk raw when it does not shrink
                # + 4 bytes per chunk: on all-zero 4 KiB chunks the first probed grid position costs up to three more
                # literals than the reference's first match at position 1 (27 / 29 vs 26 bytes per chunk)
                assert total <= RATIO_TOLERANCE * ref_total + 4 * len(cs), (kind, chunk, accel, total, ref_total)
    finally:
        gpu_ctx.set_option("accel", "1")


def test_segment_boundary_chunk_sizes(gpu_ctx):
    """Chunks larger than 64 KiB are compressed in 48 KiB segments whose inner ones may match up to their last byte; the
    end-of-block rules of the whole chunk then rest on the last segment, which must not be tiny.  Found by
    scripts/fuzz_gpu.py: a 49 153-byte chunk came out with a single literal behind its last match (rejected by
    LZ4_decompress_safe).  Every size from a segment multiple to 20 bytes past it, on every corpus."""
    sizes = [49152 * m + k for m in (1, 2, 3) for k in range(0, 21)] + [70000]      # 70000 forces the segmented path
    cs = np.asarray(sizes, dtype=np.uint64)
    n = int(cs.sum())
    for kind, seed in CORPORA + [("mixed", 9)]:
        data = L.corpus(kind, n, seed=seed)
        packed, compr, total = gpu_ctx.compress_batch(data, cs)
        check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
        out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
        assert r == n and np.array_equal(out, data)


def _dependency_stress(n, seed):
    """Data that exercises what the lane-per-sequence decoder (v7) does inside a group of 32 sequences: records that
    copy from the previous record (chains for source forwarding), short periodic runs (overlapping matches executed by
    one lane), matches that straddle literal/match boundaries, literal runs and matches around the 64-byte lane limit."""
    rng = np.random.default_rng(seed)
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    words = [rng.integers(97, 123, int(rng.integers(2, 12))).astype(np.uint8) for _ in range(300)]
    while pos < n:
        kind = int(rng.integers(0, 7))
        if kind == 0:                                    # records: counter + constant + noise, 8..40 bytes each
            rl = int(rng.integers(8, 41)); cnt = int(rng.integers(4, 200))
            rec = rng.integers(0, 256, rl).astype(np.uint8)
            buf = np.tile(rec, cnt).reshape(cnt, rl)
            buf[:, 0] = np.arange(cnt) & 255
            if rl > 12: buf[:, rl // 2] = rng.integers(0, 256, cnt)
            b = buf.reshape(-1)
        elif kind == 1:                                  # short periodic runs: period 1..7, length 5..80
            parts = []
            for _ in range(int(rng.integers(4, 60))):
                per = int(rng.integers(1, 8)); ln = int(rng.integers(5, 81))
                parts.append(np.resize(rng.integers(0, 256, per).astype(np.uint8), ln))
                parts.append(rng.integers(0, 256, int(rng.integers(0, 6))).astype(np.uint8))
            b = np.concatenate(parts)
        elif kind == 2:                                  # text from a small vocabulary
            b = np.concatenate([np.append(words[int(i)], 32) for i in rng.zipf(1.3, int(rng.integers(20, 400))) % len(words)]).astype(np.uint8)
        elif kind == 3:                                  # literal runs around the lane limit
            b = rng.integers(0, 256, int(rng.integers(56, 75))).astype(np.uint8)
        elif kind == 4 and pos > 200:                    # a copy of earlier output, 56..74 bytes or long, near or far
            ln = int(rng.integers(56, 75)) if rng.integers(0, 2) else int(rng.integers(100, 3000))
            back = int(rng.integers(ln, min(pos, 65000) + 1)) if pos > ln else pos
            b = out[pos - back:pos - back + ln].copy()
        elif kind == 5:                                  # long runs
            b = np.full(int(rng.integers(64, 5000)), int(rng.integers(0, 256)), dtype=np.uint8)
        else:
            b = rng.integers(0, 256, int(rng.integers(1, 600))).astype(np.uint8)
        b = b[:n - pos]
        out[pos:pos + len(b)] = b
        pos += len(b)
    return out


@pytest.mark.parametrize("kernels", ["v5", "v7"])
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_dependency_stress_blocks(gpu_ctx, kernels, seed):
    gpu_ctx.set_option("kernels", kernels)
    n = 6 * (1 << 20) + 4321
    data = _dependency_stress(n, seed)
    for chunk in (65536, 16384, 3000):
        cs = L.chunk_list(n, chunk)
        packed, compr = O.oracle_chunked_compress(data, cs)
        out, r = gpu_ctx.decompress_batch(packed, compr, cs)
        assert r == n
        bad = np.flatnonzero(out != data)
        assert bad.size == 0, (chunk, int(bad[0]), int(bad[0]) // chunk, int(bad[0]) % chunk)
        packed, compr, total = gpu_ctx.compress_batch(data, cs)
        out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
        assert r == n and np.array_equal(out, data)


def test_v7_unaligned_output_and_input(gpu_ctx):
    """The lane-per-sequence decoder flushes its window with 16-byte stores relative to the chunk start rounded down to
    16 bytes: every misalignment of the output and of the compressed stream."""
    data = L.corpus("mixed", 400000, seed=12)
    for lead in range(0, 17):
        sizes = [lead + 1] + [20000 + 3 * lead, 65536, 31, 65535, 50000]
        cs = np.asarray(sizes, dtype=np.uint64)
        n = int(cs.sum())
        packed, compr = O.oracle_chunked_compress(data[:n], cs)
        out, r = gpu_ctx.decompress_batch(packed, compr, cs)
        assert r == n and np.array_equal(out, data[:n]), lead


def test_contexts_on_two_devices_in_one_process():
    """The shared-memory opt-in of the tiled kernels is per device: a second context on another GPU must work
    (the driver runs one process per GPU; a host application may not)."""
    if L.lib.lz4b200_device_count() < 2:
        pytest.skip("one GPU visible")
    data = L.corpus("mixed", 3 * 65536 + 5, seed=4)
    cs = L.chunk_list(len(data), 65536)
    for dev in (0, 1, 0):
        with L.Lz4B200(dev) as ctx:
            packed, compr, total = ctx.compress_batch(data, cs)
            check_gpu_blocks_with_reference(data, packed[:total], compr, cs)
            out, r = ctx.decompress_batch(packed[:total], compr, cs)
            assert r == len(data) and np.array_equal(out, data)


def test_registered_host_buffers_can_be_released(gpu_ctx):
    """host_register page-locks caller buffers on first sight; lz4b200_unregister_host releases one before it is freed."""
    data = L.corpus("mixed", 4 << 20, seed=5)
    cs = L.chunk_list(len(data), 65536)
    gpu_ctx.set_option("host_register", "1")
    try:
        packed, compr, total = gpu_ctx.compress_batch(data, cs)
        assert total > 0
        assert L.lib.lz4b200_unregister_host(gpu_ctx._h, data.ctypes.data) == 0
        assert L.lib.lz4b200_unregister_host(gpu_ctx._h, data.ctypes.data) == -1
        out, r = gpu_ctx.decompress_batch(packed[:total], compr, cs)
        assert r == len(data) and np.array_equal(out, data)
    finally:
        gpu_ctx.set_option("host_register", "0")
