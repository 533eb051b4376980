"""CPU test of the N>1 host logic with world_size 2 over gloo: contiguous chunk ranges per rank, all-gather of
shard totals, global offsets.  The per-rank codec here is the ORACLE (no GPU in this container); the exchange
and the layout arithmetic are exactly what bench.py runs over NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, chunk, tmpdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import lzbench_b200 as L
    from lzbench_b200 import sharding as S
    from tests import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cs = L.chunk_list(n, chunk)
    lo, hi, b0, b1 = S.shard_byte_range(cs, rank, world)
    # every rank generates only its own bytes of the corpus (block-independent generator)
    data = L.corpus("mixed", b1 - b0, seed=3, offset=b0)
    packed, compr = O.oracle_chunked_compress(data, cs[lo:hi])
    totals = S.exchange_totals(len(packed))
    base = S.shard_base_offset(totals, rank)
    np.save(os.path.join(tmpdir, f"shard{rank}.npy"), packed)
    np.save(os.path.join(tmpdir, f"compr{rank}.npy"), compr)
    with open(os.path.join(tmpdir, f"meta{rank}.txt"), "w") as f:
        f.write(f"{lo} {hi} {b0} {b1} {base} {sum(totals)}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    import lzbench_b200 as L
    from tests import oracle_lib as O
    n, chunk, world = 9 * 65536 + 12345, 65536, 2     # odd chunk count, ragged tail
    mp.spawn(_worker, args=(world, 29517, n, chunk, str(tmp_path)), nprocs=world, join=True)
    cs = L.chunk_list(n, chunk)
    data = L.corpus("mixed", n, seed=3)
    ref_packed, ref_compr = O.oracle_chunked_compress(data, cs)
    image = np.zeros(len(ref_packed), dtype=np.uint8)
    compr = np.zeros(len(cs), dtype=np.uint64)
    for r in range(world):
        lo, hi, b0, b1, base, total = map(int, open(tmp_path / f"meta{r}.txt").read().split())
        assert total == len(ref_packed)
        shard = np.load(tmp_path / f"shard{r}.npy")
        image[base:base + len(shard)] = shard
        compr[lo:hi] = np.load(tmp_path / f"compr{r}.npy")
    assert np.array_equal(compr, ref_compr)
    assert np.array_equal(image, ref_packed)
    rc, out = O.oracle_chunked_decompress(image, compr, cs)
    assert rc == n and np.array_equal(out, data)


def test_shard_ranges_partition():
    from lzbench_b200 import sharding as S
    for n in (1, 7, 16384, 131072):
        for world in (1, 2, 4, 8):
            ranges = [S.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
