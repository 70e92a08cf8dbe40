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


def gather_buffers(buf, dst: int = 0, device: Optional[str] = None, state: Optional[dict] = None):
    """Gather one byte buffer per rank to `dst` with torch.distributed (nccl or gloo).

    `buf` is one array or a sequence of arrays (concatenated on the fly, no host-side copy of the whole buffer).
    Returns on `dst` a list with one uint8 array per rank (views into a receive buffer that the next call with the
    same `state` overwrites), None elsewhere.  `state` (a dict the caller keeps) caches the staging tensors between
    calls: device staging on every rank, pinned receive buffer on `dst`."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device else torch.device("cpu")
    on_gpu = dev.type == "cuda"
    parts = [buf] if isinstance(buf, np.ndarray) else list(buf)
    flats = [np.ascontiguousarray(p).view(np.uint8).reshape(-1) for p in parts]
    total = int(sum(f.size for f in flats))
    st = state if state is not None else {}
    # lengths first (they may differ per rank and per call), one small collective and one host sync
    mine = torch.tensor([total], dtype=torch.int64, device=dev)
    allsz = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allsz, mine)
    sizes = [int(x) for x in allsz.cpu().tolist()]
    mx = max(max(sizes), 1)
    if st.get("cap", 0) < mx:
        cap = mx + mx // 8
        st["cap"] = cap
        st["stage"] = torch.zeros(cap, dtype=torch.uint8, device=dev)
        if rank == dst:
            st["recv"] = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
            st["host"] = torch.empty((world, cap), dtype=torch.uint8, pin_memory=on_gpu)
    stage = st["stage"]
    o = 0
    for f in flats:
        if f.size:
            stage[o:o + f.size].copy_(torch.from_numpy(f), non_blocking=True)
            o += f.size
    send = stage[:mx]
    out = [r[:mx] for r in st["recv"]] if rank == dst else None
    dist.gather(send, out, dst=dst)
    if rank != dst:
        # wait for the send: the next step's persistent kernels would otherwise occupy every SM and keep the
        # collective's own kernel (and with it the receiving rank) waiting until they finish
        if on_gpu:
            torch.cuda.current_stream(dev).synchronize()
        return None
    host = st["host"]
    for i in range(world):
        if sizes[i]:
            host[i, :sizes[i]].copy_(out[i][:sizes[i]], non_blocking=True)
    if on_gpu:
        torch.cuda.current_stream(dev).synchronize()
    return [host[i, :sizes[i]].numpy() for i in range(world)]
