"""Generate tests/golden/golden.json from the REFERENCE compiled here (oracle/_ref/liblz4ref.so,
built by oracle/Makefile from /root/reference/lz4/lz4.c).  Run in the build container only:

    python tests/golden/make_golden.py

Each vector pins (corpus bytes, reference LZ4_compress_default output) by SHA-256; each small
block pins exact compressed bytes + plaintext in hex."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lzbench_b200 as L          # noqa: E402
from tests import oracle_lib as O  # noqa: E402

assert O.ref is not None, "build oracle/_ref first (make -C oracle)"
vectors = []
for kind, seed in (("mixed", 1), ("mixed", 2), ("random", 4), ("repeat", 5), ("zero", 5)):
    for n in (13, 1000, 4096, 65536, 65547, 262144):
        data = L.corpus(kind, n, seed=seed)
        blk = O.ref_compress(data)
        vectors.append({"kind": kind, "seed": seed, "n": n, "clen": int(len(blk)),
                        "plain_sha256": hashlib.sha256(data.tobytes()).hexdigest(),
                        "block_sha256": hashlib.sha256(blk.tobytes()).hexdigest()})
blocks = []
for kind, seed, n in (("mixed", 1, 200), ("mixed", 3, 333), ("repeat", 5, 150), ("zero", 5, 100), ("random", 4, 40)):
    data = L.corpus(kind, n, seed=seed)
    blk = O.ref_compress(data)
    r, out = O.ref_decompress_safe(blk, n)
    assert r == n and np.array_equal(out, data)
    blocks.append({"kind": kind, "seed": seed, "n": n, "block_hex": blk.tobytes().hex(), "plain_hex": data.tobytes().hex()})
with open(os.path.join(os.path.dirname(__file__), "golden.json"), "w") as f:
    json.dump({"generator": "oracle/_ref/liblz4ref.so = reference lz4 " + O.ref.LZ4_versionString().decode(),
               "vectors": vectors, "blocks": blocks}, f, indent=1)
print("wrote", len(vectors), "vectors,", len(blocks), "blocks")
