"""Debug aid: what do the compressed blocks of large incompressible / run-length chunks look like (sequence list)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, lzbench_b200 as L
def seqs(blk, limit=12):
    ip = 0; out = []; n = len(blk)
    while len(out) < limit:
        t = blk[ip]; ip += 1; l = t >> 4
        if l == 15:
            while True:
                b = blk[ip]; ip += 1; l += b
                if b != 255: break
        ip += l
        if ip >= n: out.append((l, 0, 0)); return out
        off = blk[ip] | (blk[ip + 1] << 8); ip += 2; m = t & 15
        if m == 15:
            while True:
                b = blk[ip]; ip += 1; m += b
                if b != 255: break
        out.append((l, off, m + 4))
    return out
ctx = L.Lz4B200(0)
for kind in ("random", "zero", "repeat"):
    n = 1 << 20
    d = L.corpus(kind, n, seed=5)
    cs = L.chunk_list(n, n)
    packed, compr, total = ctx.compress_batch(d, cs)
    print(kind, "compressed", total, "sequences:", seqs(bytes(packed[:total])))
