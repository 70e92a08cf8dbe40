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
    """(begin, end] of the running observation total owned by one bin: floor(T*i/B) computed in double like
    getBoundaryCount (GSC/EdgeRetrieverBin.cpp:38-41)"""
    assert 0 <= bin_index < bin_count
    import math
    begin = int(math.floor((float(total_observations) * float(bin_index)) / float(bin_count)))
    end = int(math.floor((float(total_observations) * float(bin_index + 1)) / float(bin_count)))
    return begin, end


def bins_for_rank(bin_count: int, rank: int, world: int) -> List[int]:
    return [b for b in range(bin_count) if b % world == rank]


class FlatLocusGraph:
    """CSR copy of an SVLocusSet (include/manta_b200_graph.hpp mb200::FlatLocusGraph)"""

    def __init__(self, locusNodeOff, nodeEdgeOff, edgeTo, edgeCount):
        self.locusNodeOff = np.ascontiguousarray(locusNodeOff, dtype=np.uint32)
        self.nodeEdgeOff = np.ascontiguousarray(nodeEdgeOff, dtype=np.uint64)
        self.edgeTo = np.ascontiguousarray(edgeTo, dtype=np.uint32)
        self.edgeCount = np.ascontiguousarray(edgeCount, dtype=np.uint32)

    @classmethod
    def from_pair_counts(cls, observation_counts: Sequence[int]) -> "FlatLocusGraph":
        """one two-node locus per candidate, all observations on the node0 -> node1 edge (what a graph of isolated
        breakend pairs looks like)"""
        c = np.asarray(observation_counts, dtype=np.uint32)
        n = len(c)
        edge_to = np.tile(np.array([1, 0], dtype=np.uint32), n)
        edge_count = np.zeros(2 * n, dtype=np.uint32)
        edge_count[0::2] = c
        return cls(np.arange(n + 1, dtype=np.uint32) * 2, np.arange(2 * n + 1, dtype=np.uint64), edge_to, edge_count)

    @property
    def n_loci(self) -> int:
        return len(self.locusNodeOff) - 1


EDGE_INFO_DT = np.dtype([("locusIndex", "<u4"), ("nodeIndex1", "<u4"), ("nodeIndex2", "<u4")])


def enumerate_bin(graph: FlatLocusGraph, bin_count: int, bin_index: int, graph_node_max_edge_count: int = 0) -> np.ndarray:
    """the edges of one bin, through the library's mb200_edge_bin_enumerate (== the reference's EdgeRetrieverBin,
    tests/test_edge_bins.py); host arithmetic, no GPU needed"""
    import ctypes as C
    from . import library_path
    lib = C.CDLL(library_path())
    cap = max(16, len(graph.edgeTo))
    out = np.zeros(cap, dtype=EDGE_INFO_DT)
    n = C.c_uint64(0)
    st = lib.mb200_edge_bin_enumerate(C.c_uint32(graph.n_loci), graph.locusNodeOff.ctypes, graph.nodeEdgeOff.ctypes,
                                      graph.edgeTo.ctypes, graph.edgeCount.ctypes, C.c_uint32(graph_node_max_edge_count),
                                      C.c_uint32(bin_count), C.c_uint32(bin_index), out.ctypes, C.c_uint64(cap), C.byref(n))
    if st != 0:
        raise RuntimeError("mb200_edge_bin_enumerate failed with status %d" % st)
    return out[:n.value]


def assign_edges(observation_counts: Sequence[int], bin_count: int, rank: int, world: int) -> np.ndarray:
    """indices of the candidates (one two-node locus each) this rank processes: the bins {b : b mod world == rank} of a
    `bin_count`-bin split, enumerated by the library's EdgeRetrieverBin"""
    g = FlatLocusGraph.from_pair_counts(observation_counts)
    mine = [enumerate_bin(g, bin_count, b)["locusIndex"] for b in bins_for_rank(bin_count, rank, world)]
    return np.sort(np.concatenate(mine)) if mine else np.zeros(0, dtype=np.uint32)


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


class NcclGather:
    """The path's one collective through the library's C ABI (mb200_gather_results: lengths all-gathered, then ncclSend /
    ncclRecv straight into the merged buffer on the root).  The communicator is the library's own: rank 0 draws the NCCL
    unique id and `bootstrap_broadcast` (any transport: here torch.distributed, gloo or nccl) hands it to the others.
    torch only provides the device / pinned staging memory."""

    def __init__(self, ctx, rank: int, world: int, device: str, bootstrap_broadcast):
        import torch
        self.ctx, self.rank, self.world = ctx, rank, world
        self.torch, self.dev = torch, torch.device(device)
        uid = bootstrap_broadcast(ctx.nccl_unique_id() if rank == 0 else None)
        self.comm = ctx.nccl_comm_init(world, uid, rank)
        self.cap = 0
        self.stage = self.recv = self.host = None

    def close(self):
        if self.comm is not None:
            self.ctx.nccl_comm_destroy(self.comm)
            self.comm = None

    def gather(self, buf, dst: int = 0):
        """`buf`: one array or a sequence of arrays (this rank's result buffer).  Returns on `dst` one uint8 array per rank
        (views into a pinned buffer that the next call overwrites), None elsewhere."""
        torch = self.torch
        parts = [buf] if isinstance(buf, np.ndarray) else list(buf)
        flats = [np.ascontiguousarray(p).view(np.uint8).reshape(-1) for p in parts]
        total = int(sum(f.size for f in flats))
        if self.stage is None or self.stage.numel() < total:
            self.stage = torch.empty(total + total // 8 + 64, dtype=torch.uint8, device=self.dev)
        o = 0
        for f in flats:
            if f.size:
                self.stage[o:o + f.size].copy_(torch.from_numpy(f), non_blocking=True)
                o += f.size
        torch.cuda.current_stream(self.dev).synchronize()  # the library gathers on its own stream
        for _ in range(2):
            if self.rank == dst and self.recv is None:
                self.cap = max(self.cap, 1 << 16)
                self.recv = torch.empty(self.cap, dtype=torch.uint8, device=self.dev)
                self.host = torch.empty(self.cap, dtype=torch.uint8, pin_memory=True)
            st, sizes = self.ctx.gather_results(self.comm, dst, self.stage.data_ptr(), total,
                                                self.recv.data_ptr() if self.rank == dst else 0,
                                                self.cap if self.rank == dst else 0, self.world)
            if st == 0:
                break
            from ._cabi import MB200_ERR_POOL_OVERFLOW
            if st != MB200_ERR_POOL_OVERFLOW:
                self.ctx._raise(st)
            # every rank got the same sizes and the same status: the root grows its buffer, everybody calls again
            self.cap = int(sum(sizes) * 1.25) + 64
            self.recv = None
        else:
            raise RuntimeError("gather_results: merged buffer still too small")
        if self.rank != dst:
            return None
        n = int(sum(sizes))
        self.host[:n].copy_(self.recv[:n], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        out, o = [], 0
        h = self.host.numpy()
        for s_ in sizes:
            out.append(h[o:o + s_])
            o += s_
        return out
