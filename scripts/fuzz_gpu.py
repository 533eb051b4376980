"""Extended randomized check of the default kernels on the GPU box (not part of the pytest suite: different seeds every
run unless FUZZ_SEED is set).  Round trips on random mixtures of the corpora at random chunk sizes, reference-compressed
blocks through the GPU decoder, and corrupted blocks whose accept / reject decision and output must match the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lzbench_b200 as L
from tests import oracle_lib as O

seed = int(os.environ.get("FUZZ_SEED", time.time_ns() % (1 << 31)))
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(seed)
print("fuzz seed", seed, flush=True)
ctx = L.Lz4B200(0)
if os.environ.get("FUZZ_KERNELS"): ctx.set_option("kernels", os.environ["FUZZ_KERNELS"])
if os.environ.get("FUZZ_COMPRESS"): ctx.set_option("compress", os.environ["FUZZ_COMPRESS"])
ACCEL = int(os.environ.get("FUZZ_ACCEL", "1"))
ctx.set_option("accel", str(ACCEL))
KINDS = ["mixed", "random", "repeat", "zero"]
t_end = time.time() + secs
rounds = blocks = corrupt = ratio_fail = frames_bad = 0
while time.time() < t_end:
    # ---- a random mixture, random chunking ----
    parts = []; desc = []
    for _ in range(int(rng.integers(1, 6))):
        k = KINDS[int(rng.integers(0, 4))]
        ln = int(rng.integers(1, 1 << int(rng.integers(4, 22))))
        sd = int(rng.integers(1, 1000))
        parts.append(L.corpus(k, ln, seed=sd)); desc.append((k, ln, sd))
    data = np.concatenate(parts)
    n = len(data)
    chunk = int(rng.choice([13, 64, 1000, 4096, 16384, 65535, 65536, 65537, 65547, 100000, 262144, 1 << 20,
                            49152 + int(rng.integers(0, 24)), 98304 + int(rng.integers(0, 24)), 65536 + int(rng.integers(0, 40))]))
    if os.environ.get("FUZZ_ACCEL") is None:
        ACCEL = int(rng.choice([1, 1, 1, 2, 4, 9]))
        ctx.set_option("accel", str(ACCEL))
    if rng.random() < 0.3:                               # ragged list, lzbench -j style
        files = rng.integers(1, max(2, n // 3), int(rng.integers(1, 40)))
        files = files[np.cumsum(files) <= n]
        files = np.append(files, n - int(files.sum())) if int(files.sum()) < n else files
        cs = L.chunk_list(files[files > 0], chunk)
    else:
        cs = L.chunk_list(n, chunk)
    packed, compr, total = ctx.compress_batch(data, cs)
    out, r = ctx.decompress_batch(packed[:total], compr, cs)
    if not (r == n and np.array_equal(out, data)):
        bad = np.nonzero(out[:n] != data)[0] if r == n else np.array([-1])
        print("ROUND TRIP FAILURE", seed, rounds, "chunk", chunk, "n", n, "r", r, "first mismatch", int(bad[0]), "count", len(bad), "parts", desc, flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez_compressed("gpurun_out/fuzz_fail.npz", data=data, cs=cs, packed=packed[:total], compr=compr, out=out[:n])
        raise SystemExit(1)
    p2, c2 = O.oracle_chunked_compress(data, cs)          # reference at acceleration 1: a lower bound for the accelerated levels too
    out, r = ctx.decompress_batch(p2, c2, cs)
    assert r == n and np.array_equal(out, data), ("reference blocks", seed, rounds)
    if ACCEL == 1 and total > 1.03 * len(p2) + 64 * len(cs):
        print("RATIO", seed, rounds, "chunk", chunk, "n", n, "chunks", len(cs), "gpu", total, "ref", len(p2), "parts", desc, flush=True)
        worst = sorted(range(len(cs)), key=lambda i: int(compr[i]) - int(c2[i]))[-3:]
        print("   worst chunks (size, gpu, ref):", [(int(cs[i]), int(compr[i]), int(c2[i])) for i in worst], flush=True)
        ratio_fail += 1
    # every GPU block must be accepted by the oracle validator
    offp = 0; offu = 0
    step = 1 if len(cs) <= 400 else max(1, len(cs) // 100)
    offs_p = np.concatenate([[0], np.cumsum(compr)]).astype(np.int64)
    offs_u = np.concatenate([[0], np.cumsum(cs)]).astype(np.int64)
    for i in range(0, len(cs), step):
        z, u = int(compr[i]), int(cs[i])
        if z != u:
            assert O.oracle_validate(packed[offs_p[i]:offs_p[i] + z], u) == 0, ("validator", seed, rounds, i)
        blocks += 1
    # ---- corrupted blocks: same accept / reject and same output as the oracle ----
    u = int(rng.choice([300, 4096, 8192, 65536]))
    src = L.corpus(KINDS[int(rng.integers(0, 3))], u, seed=int(rng.integers(1, 1000)))
    good = O.oracle_compress(src)
    if len(good) < u:
        bl, exp = [], []
        for _ in range(256):
            bad = good.copy()
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
            bl.append(bad)
            rr, oo = O.oracle_decompress(bad, u)
            exp.append((rr == u, oo))
        pk = np.concatenate(bl)
        out, r = ctx.decompress_batch(pk, np.full(len(bl), len(good), dtype=np.uint64), np.full(len(bl), u, dtype=np.uint64))
        nb = sum(1 for ok, _ in exp if not ok)
        assert (r == -nb) if nb else (r == u * len(bl)), ("corrupt count", seed, rounds, r, nb)
        out = out.reshape(len(bl), u)
        for i, (ok, oo) in enumerate(exp):
            if ok:
                assert np.array_equal(out[i], oo), ("corrupt output", seed, rounds, i)
        corrupt += len(bl)
    # ---- frames ----
    if rounds % 4 == 0:
        bid = int(rng.integers(4, 8))
        f = ctx.frame_compress(data, bid, content_checksum=bool(rng.integers(0, 2)), block_checksum=bool(rng.integers(0, 2)),
                               content_size=bool(rng.integers(0, 2)))
        rr, oo = O.oracle_frame_decompress(f, n + 8)
        assert rr == n and np.array_equal(oo, data), ("frame -> oracle", seed, rounds)
        rr, oo = ctx.frame_decompress(O.oracle_frame_compress(data, bid, 1, 1), n)
        assert rr == n and np.array_equal(oo, data), ("oracle frame -> gpu", seed, rounds, L.last_error())
        # damaged frames: same verdict as the oracle, and when both accept, the same bytes
        for _ in range(8):
            g = f.copy()
            for _ in range(int(rng.integers(1, 4))):
                g[int(rng.integers(0, len(g)))] ^= 1 << int(rng.integers(0, 8))
            if rng.random() < 0.2: g = g[:int(rng.integers(1, len(g)))].copy()
            ro, oo = O.oracle_frame_decompress(g, n + 8)
            rg, og = ctx.frame_decompress(g, n + 8)
            if ro == O.BAD:
                assert rg < 0, ("damaged frame accepted", seed, rounds, rg)
            else:
                assert rg == ro and np.array_equal(og, oo), ("damaged frame verdict", seed, rounds, rg, ro, L.last_error())
            frames_bad += 1
    rounds += 1
print(f"fuzz {'ok' if not ratio_fail else 'RATIO FAILURES ' + str(ratio_fail)}: {rounds} rounds, {blocks} blocks validated, {corrupt} corrupted blocks, {frames_bad} damaged frames, seed {seed}", flush=True)
