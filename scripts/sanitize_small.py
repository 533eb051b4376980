"""Small all-generations round trip for compute-sanitizer memcheck (keeps the run short)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lzbench_b200 as L
from tests import oracle_lib as O

ctx = L.Lz4B200(0)
rng = np.random.default_rng(3)
sizes = [1, 12, 13, 100, 4096, 65536, 65536, 70000, 65535, 33333] + [int(x) for x in rng.integers(1, 5000, 40)]
cs = np.asarray(sizes, dtype=np.uint64)
n = int(cs.sum())
for kind in ("mixed", "random", "zero"):
    data = L.corpus(kind, n, seed=2)
    p_ref, c_ref = O.oracle_chunked_compress(data, cs)
    for gen in ("v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"):
        ctx.set_option("kernels", gen)
        packed, compr, total = ctx.compress_batch(data, cs)
        out, r = ctx.decompress_batch(packed[:total], compr, cs)
        assert r == n and np.array_equal(out, data), (kind, gen)
        out, r = ctx.decompress_batch(p_ref, c_ref, cs)
        assert r == n and np.array_equal(out, data), (kind, gen, "ref blocks")
    # corrupted blocks must be rejected without touching memory out of bounds
    bad = p_ref.copy()
    for i in rng.integers(0, len(bad), 200): bad[int(i)] = int(rng.integers(0, 256))
    for gen in ("v0", "v1", "v2", "v3", "v4", "v5", "v6", "v7"):
        ctx.set_option("kernels", gen)
        ctx.decompress_batch(bad, c_ref, cs)
print("sanitize_small ok")
