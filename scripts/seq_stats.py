"""Sequence statistics of the GPU compressor's output on the bench corpus (what the decoders see)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lzbench_b200 as L
from tests import oracle_lib as O
n = 16 << 20
data = L.corpus("mixed", n, seed=2)
cs = L.chunk_list(n, 65536)
with L.Lz4B200(0) as ctx:
    packed, compr, total = ctx.compress_batch(data, cs)
for name, (p, c) in {"gpu": (packed[:total], compr), "oracle": O.oracle_chunked_compress(data, cs)}.items():
    st = O.SeqStats(); off = 0
    for z in c.tolist():
        z = int(z); assert O.oracle_validate(p[off:off + z], 65536, st) == 0; off += z
    print(name, "ratio %.2f%%" % (100 * off / n), "seq/chunk %.0f" % (st.n_seq / len(cs)), "lit %.1f%%" % (100 * st.lit_bytes / n),
          "zero-lit %.1f%%" % (100 * st.zero_lit_seq / st.n_seq), "overlap %.2f%%" % (100 * st.overlap_seq / st.n_seq), "max seq/chunk", st.max_seq_in_block)
