"""Multi-GPU partitioning of the chunked path.

Chunks share no state (fresh table per chunk, offsets never leave the chunk: lz4/lz4.c:1361-1364,
_lzbench/lzbench.cpp:275-296), so the path shards by contiguous chunk ranges with no data-path
collective.  The only exchange is the all-gather of the per-rank compressed totals (8 bytes per rank)
that turns rank-local offsets into offsets in the global concatenation lzbench expects.
One process per GPU; the process group is NCCL on the GPU box and gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank g of G owns chunks [g*n/G, (g+1)*n/G)."""
    return n_chunks * rank // world, n_chunks * (rank + 1) // world


def shard_byte_range(chunk_sizes: np.ndarray, rank: int, world: int) -> Tuple[int, int, int, int]:
    """(first chunk, last chunk + 1, first byte, last byte + 1) of the rank's slice of the uncompressed image."""
    lo, hi = shard_range(len(chunk_sizes), rank, world)
    cs = np.asarray(chunk_sizes, dtype=np.uint64)
    b0 = int(cs[:lo].sum())
    return lo, hi, b0, b0 + int(cs[lo:hi].sum())


def exchange_totals(local_total: int, group=None, device=None) -> List[int]:
    """All-gather of the shard totals.  Returns the list of every rank's compressed size."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.tensor([int(local_total)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [int(x.item()) for x in out]


def shard_base_offset(totals: List[int], rank: int) -> int:
    """Exclusive scan of the totals: where this rank's compressed shard starts in the global image."""
    return int(sum(totals[:rank]))
