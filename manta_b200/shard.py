"""Multi-GPU plumbing of the candidate path.

Candidate loci (graph edges) are independent, so the path shards with NO data-path collective; the reference already
defines the partition -- ``EdgeRetrieverBin(binCount, binIndex)`` gives bin i the edges whose cumulative observation
count lies in [floor(T*i/B), floor(T*(i+1)/B))  (GSC/EdgeRetrieverBin.cpp:38-61).  Rank r of G processes the bins
{b : b mod G == r}.  The single collective is the gather of the per-rank serialized result buffers to rank 0
(variable length: lengths are all-gathered first, then a padded gather), over NCCL/NVLink on GPUs and gloo in the
CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def bin_bounds(total_observations: int, bin_count: int, bin_index: int) -> Tuple[int, int]:
    """[begin, end) of cumulative observation count owned by one bin (EdgeRetrieverBin.cpp:38-61)"""
    assert 0 <= bin_index < bin_count
    begin = (total_observations * bin_index) // bin_count
    end = (total_observations * (bin_index + 1)) // bin_count
    return begin, end


def bins_for_rank(bin_count: int, rank: int, world: int) -> List[int]:
    return [b for b in range(bin_count) if b % world == rank]


def assign_edges(observation_counts: Sequence[int], bin_count: int, rank: int, world: int) -> np.ndarray:
    """indices of the edges this rank processes: edge e belongs to the bin containing the cumulative count at which
    the edge starts (edges are never split, like EdgeRetrieverBin::jumpToFirstEdge / next)"""
    counts = np.asarray(observation_counts, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]) if len(counts) else np.zeros(0, dtype=np.int64)
    total = int(counts.sum())
    mine = np.zeros(len(counts), dtype=bool)
    for b in bins_for_rank(bin_count, rank, world):
        lo, hi = bin_bounds(total, bin_count, b)
        mine |= (starts >= lo) & (starts < hi) & (counts > 0)
    return np.nonzero(mine)[0]


def gather_buffers(buf: np.ndarray, dst: int = 0, device: Optional[str] = None) -> Optional[List[np.ndarray]]:
    """gather one uint8 buffer per rank to `dst` (None elsewhere) with torch.distributed (nccl or gloo)"""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device else torch.device("cpu")
    flat = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    n = torch.tensor([flat.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    padded = torch.zeros(max(mx, 1), dtype=torch.uint8, device=dev)
    if flat.size:
        padded[:flat.size] = torch.from_numpy(flat.copy()).to(dev)
    out = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, out, dst=dst)
    if rank != dst:
        return None
    return [o[:sizes[i]].cpu().numpy() for i, o in enumerate(out)]
