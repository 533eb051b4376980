"""GPU parity tests for the LZ4 frame container, through the C ABI (lz4b200_frame_*).
Checker: oracle/lz4f_oracle.c and, when it travelled, the reference's own lz4frame.c (oracle/_ref/liblz4fref.so).
Bars: a GPU frame decodes bit-exactly through LZ4F_decompress and the oracle; reference and oracle frames decode
bit-exactly on the GPU; frame size <= 1.03 x the reference's; damaged frames are rejected, never mis-decoded."""
import numpy as np
import pytest

import lzbench_b200 as L
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 12, 13, 65535, 65536, 65537, 300000, (4 << 20) + 5, (9 << 20) + 123]


def decoders():
    yield "oracle", O.oracle_frame_decompress
    if O.reff is not None:
        yield "LZ4F_decompress", O.ref_frame_decompress


@pytest.mark.parametrize("block_id", [4, 5, 6, 7])
@pytest.mark.parametrize("kind,seed", [("mixed", 1), ("random", 4), ("repeat", 5), ("zero", 5)])
def test_gpu_frames_decode_with_reference(gpu_ctx, block_id, kind, seed):
    for n in SIZES:
        d = L.corpus(kind, max(n, 1), seed=seed)[:n]
        for flags in ({}, {"content_checksum": True, "content_size": True}, {"block_checksum": True},
                      {"content_checksum": True, "block_checksum": True, "content_size": True}):
            f = gpu_ctx.frame_compress(d, block_id, **flags)
            for name, dec in decoders():
                r, out = dec(f, n + 16)
                assert r == n and np.array_equal(out, d), (name, n, flags)
            r, out = gpu_ctx.frame_decompress(f, n)
            assert r == n and np.array_equal(out, d), ("gpu", n, flags, L.last_error())
            want = L.lib.lz4b200_frame_content_size(f.ctypes.data, len(f))
            assert want == (n if flags.get("content_size") else -3)
        if O.reff is not None and n >= 65536:
            ref = O.ref_frame_compress(d, block_id, 1)
            assert len(gpu_ctx.frame_compress(d, block_id)) <= 1.03 * len(ref) + 16, (n, kind)


@pytest.mark.parametrize("block_id", [4, 5, 6, 7])
@pytest.mark.parametrize("kind,seed", [("mixed", 2), ("random", 3), ("repeat", 6), ("zero", 1)])
def test_reference_frames_decode_on_gpu(gpu_ctx, block_id, kind, seed):
    for n in SIZES:
        d = L.corpus(kind, max(n, 1), seed=seed)[:n]
        frames = [O.oracle_frame_compress(d, block_id, 1, 1)]
        if O.reff is not None:
            frames += [O.ref_frame_compress(d, block_id, 1, cck, bck, csz)
                       for cck, bck, csz in ((0, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 1))]
        for f in frames:
            r, out = gpu_ctx.frame_decompress(f, n + 7)
            assert r == n and np.array_equal(out, d), (n, L.last_error())


def test_frames_with_short_blocks(gpu_ctx):
    """Blocks may be shorter than the frame's block size (a streaming encoder that flushes): build such a frame by
    hand from oracle blocks of ragged sizes."""
    d = L.corpus("mixed", 1 << 20, seed=11)
    cuts = [0, 1000, 70000, 70013, 200000, 265536, 700001, 1 << 20]
    parts = [np.frombuffer(b"\x04\x22\x4d\x18\x60\x40\x82", dtype=np.uint8)]       # 64 KiB blocks, no flags
    for a, b in zip(cuts, cuts[1:]):
        for o in range(a, b, 65536):
            p = d[o:min(o + 65536, b)]
            blk = O.oracle_compress(p)
            if len(blk) >= len(p):
                parts += [np.frombuffer(np.uint32(len(p) | 0x80000000).tobytes(), dtype=np.uint8), p]
            else:
                parts += [np.frombuffer(np.uint32(len(blk)).tobytes(), dtype=np.uint8), blk]
    parts.append(np.zeros(4, dtype=np.uint8))
    f = np.concatenate(parts)
    for name, dec in decoders():
        r, out = dec(f, len(d))
        assert r == len(d) and np.array_equal(out, d), name
    r, out = gpu_ctx.frame_decompress(f, len(d))
    assert r == len(d) and np.array_equal(out, d), L.last_error()


def test_damaged_and_unsupported_frames_rejected(gpu_ctx):
    d = L.corpus("mixed", 400000, seed=2)
    f = gpu_ctx.frame_compress(d, 4, content_checksum=True, block_checksum=True, content_size=True)
    cases = []
    for pos in (0, 4, 5, 6, 14, 15, 40, 1000, len(f) - 1, len(f) - 6, len(f) - 9):
        g = f.copy(); g[pos] ^= 0x10; cases.append((pos, g))
    cases += [("cut3", f[:len(f) - 3].copy()), ("cut20", f[:20].copy()), ("cut5", f[:5].copy())]
    for tag, g in cases:
        assert O.oracle_frame_decompress(g, len(d) + 16)[0] == O.BAD, tag
        r, _ = gpu_ctx.frame_decompress(g, len(d) + 16)
        assert r < 0, tag
    # undamaged frame, but the caller's buffer is too small
    assert gpu_ctx.frame_decompress(f, len(d) - 1)[0] < 0
    # with verification off a damaged content checksum is not looked at, block damage still is
    g = f.copy(); g[len(f) - 1] ^= 1
    r, out = gpu_ctx.frame_decompress(g, len(d), skip_checksums=True)
    assert r == len(d) and np.array_equal(out, d)
    if O.reff is not None:
        linked = O.ref_frame_compress(d, 4, 0)
        assert gpu_ctx.frame_decompress(linked, len(d))[0] == -2
        assert "linked" in L.last_error()


def test_frame_256mib_round_trip(gpu_ctx):
    n = 256 << 20
    d = L.corpus("mixed", n, seed=5)
    f = gpu_ctx.frame_compress(d, 4, block_checksum=True)
    r, out = gpu_ctx.frame_decompress(f, n)
    assert r == n and np.array_equal(out, d)
    for name, dec in decoders():
        r, out = dec(f, n)
        assert r == n and np.array_equal(out, d), name


def test_block_rules_are_checked_against_the_frame_block_size(gpu_ctx):
    """LZ4F_decompress decodes every block into a buffer of the frame's block size, so the end-of-block rules are
    relative to that capacity (lz4/lz4frame.c:1654-1712): this block ends 4 bytes behind its last match, which
    LZ4_decompress_safe rejects when told the exact size (27) and accepts with room to spare.  Frames must follow the
    frame decoder, chunks the block decoder."""
    blk = np.frombuffer(bytes([0x4F]) + b"abcd" + bytes([4, 0, 0, 0x40]) + b"wxyz", dtype=np.uint8)
    want = np.frombuffer(b"abcd" + (b"abcd" * 5)[:19] + b"wxyz", dtype=np.uint8)
    frame = np.concatenate([np.frombuffer(b"\x04\x22\x4d\x18\x60\x40\x82", dtype=np.uint8),
                            np.frombuffer(np.uint32(len(blk)).tobytes(), dtype=np.uint8), blk, np.zeros(4, np.uint8)])
    for name, dec in decoders():
        r, out = dec(frame, 100)
        assert r == len(want) and np.array_equal(out, want), name
    r, out = gpu_ctx.frame_decompress(frame, 100)
    assert r == len(want) and np.array_equal(out, want), L.last_error()
    # the same block as an lzbench chunk of exactly 27 bytes is malformed for every block decoder
    assert O.oracle_decompress(blk, 27)[0] < 0
    if O.ref is not None:
        assert O.ref_decompress_safe(blk, 27)[0] < 0
    _, r = gpu_ctx.decompress_batch(blk, np.array([len(blk)], dtype=np.uint64), np.array([27], dtype=np.uint64))
    assert r < 0
