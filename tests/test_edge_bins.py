"""Row a1: the per-bin edge partition (mb200_edge_bin_enumerate / mb200::EdgeRetrieverBin, include/manta_b200_graph.hpp)
against the reference's own EdgeRetrieverBin class compiled from its sources (oracle/_ref, ref_graph_shim.cpp), on the graphs
of the reference's unit tests (GSC/test/EdgeRetrieverBinTest.cpp, expectations transcribed) and on random locus sets, for
every bin of several bin counts, with and without the noisy-node filter.  Host arithmetic only: runs without a GPU."""
import ctypes as C
import os
import random

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EDGE_DT = np.dtype([("locusIndex", "<u4"), ("nodeIndex1", "<u4"), ("nodeIndex2", "<u4")])


@pytest.fixture(scope="module")
def lib():
    import manta_b200
    return C.CDLL(manta_b200.library_path())


@pytest.fixture(scope="module")
def reflib(ref):
    return ref.lib


class RefGraph:
    """a locus set built by the reference (SVLocus::addNode/linkNodes, SVLocusSet::merge) and its flattened copy"""

    def __init__(self, reflib, specs, minMergeEdgeObservations=1):
        # specs: [(nodes [(tid, begin, end)], links [(from, to, count, remoteCount)], selfMerge)]
        nodeOff, linkOff, nodes, links, selfm = [0], [0], [], [], []
        for ns, ls, sm in specs:
            nodes += [x for n in ns for x in n]
            links += [x for l in ls for x in l]
            nodeOff.append(nodeOff[-1] + len(ns))
            linkOff.append(linkOff[-1] + len(ls))
            selfm.append(1 if sm else 0)
        A = lambda v, dt: np.asarray(v if len(v) else [0], dtype=dt)
        self.lib = reflib
        self.h = C.c_void_p()
        st = reflib.ref_graph_build(C.c_uint32(len(specs)), A(nodeOff, np.uint32).ctypes, A(nodes, np.int32).ctypes,
                                    A(linkOff, np.uint32).ctypes, A(links, np.uint32).ctypes, A(selfm, np.uint8).ctypes,
                                    C.c_uint32(minMergeEdgeObservations), C.byref(self.h))
        assert st == 0
        nLoci, nNodes, nOut, tot = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint32()
        reflib.ref_graph_sizes(self.h, C.byref(nLoci), C.byref(nNodes), C.byref(nOut), C.byref(tot))
        self.nLoci, self.total = nLoci.value, tot.value
        self.locusNodeOff = np.zeros(nLoci.value + 1, dtype=np.uint32)
        self.nodeEdgeOff = np.zeros(nNodes.value + 1, dtype=np.uint64)
        self.edgeTo = np.zeros(max(1, nOut.value), dtype=np.uint32)
        self.edgeCount = np.zeros(max(1, nOut.value), dtype=np.uint32)
        reflib.ref_graph_flat(self.h, self.locusNodeOff.ctypes, self.nodeEdgeOff.ctypes, self.edgeTo.ctypes, self.edgeCount.ctypes)

    def close(self):
        self.lib.ref_graph_free(self.h)

    def ref_edges(self, maxEdges, binCount, binIndex):
        self.lib.ref_graph_bin_edges.restype = C.c_int64
        cap = 4096
        out = np.zeros(3 * cap, dtype=np.uint32)
        n = self.lib.ref_graph_bin_edges(self.h, C.c_uint32(maxEdges), C.c_uint32(binCount), C.c_uint32(binIndex), out.ctypes,
                                         C.c_uint64(cap))
        assert 0 <= n <= cap
        return [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(n)]

    def our_edges(self, lib, maxEdges, binCount, binIndex, cap=4096):
        out = np.zeros(cap, dtype=EDGE_DT)
        n = C.c_uint64(0)
        st = lib.mb200_edge_bin_enumerate(C.c_uint32(self.nLoci), self.locusNodeOff.ctypes, self.nodeEdgeOff.ctypes,
                                          self.edgeTo.ctypes, self.edgeCount.ctypes, C.c_uint32(maxEdges), C.c_uint32(binCount),
                                          C.c_uint32(binIndex), out.ctypes, C.c_uint64(cap), C.byref(n))
        return st, [tuple(int(x) for x in e) for e in out[:min(cap, n.value)]], n.value


def pair(t1, b1, e1, t2, b2, e2, both=False, count=1):
    """test/testSVLocusUtil.cpp locusAddPair"""
    return ([(t1, b1, e1), (t2, b2, e2)], [(0, 1, count, count if both else 0)], False)


def test_reference_unit_cases(lib, reflib):
    # test_EdgeRetrieverOneBin (EdgeRetrieverBinTest.cpp:32-85)
    g = RefGraph(reflib, [pair(1, 10, 20, 2, 30, 40), pair(3, 10, 20, 4, 30, 40)])
    assert g.ref_edges(0, 1, 0) == [(0, 0, 1), (1, 0, 1)]
    assert g.our_edges(lib, 0, 1, 0)[1] == [(0, 0, 1), (1, 0, 1)]
    g.close()
    # every node has two edges, cutoff 1: nothing is returned
    g = RefGraph(reflib, [pair(1, 10, 20, 2, 30, 40), pair(1, 10, 20, 6, 30, 40), pair(7, 10, 20, 2, 30, 40), pair(7, 10, 20, 6, 30, 40)])
    assert g.ref_edges(1, 1, 0) == [] and g.our_edges(lib, 1, 1, 0)[1] == []
    assert g.our_edges(lib, 0, 1, 0)[1] == g.ref_edges(0, 1, 0) != []
    g.close()
    # test_EdgeRetrieverManyBin (:87-136): a star locus + a pair, two bins
    star = ([(1, 10, 20), (2, 30, 40), (3, 30, 40), (4, 30, 40)], [(0, 1, 1, 0), (0, 2, 1, 0), (0, 3, 1, 0)], False)
    g = RefGraph(reflib, [star, pair(7, 10, 20, 8, 30, 40)])
    want = {0: [(0, 0, 1), (0, 0, 2)], 1: [(0, 0, 3), (1, 0, 1)]}
    for b in range(2):
        assert g.ref_edges(0, 2, b) == want[b] and g.our_edges(lib, 0, 2, b)[1] == want[b]
    g.close()
    # test_EdgeRetrieverManyBin2 (:138-180): six pairs, three bins of two
    g = RefGraph(reflib, [pair(2 * i + 1, 10, 20, 2 * i + 2, 30, 40) for i in range(6)])
    for b in range(3):
        want = [(2 * b, 0, 1), (2 * b + 1, 0, 1)]
        assert g.ref_edges(0, 3, b) == want and g.our_edges(lib, 0, 3, b)[1] == want
    g.close()
    # test_EdgeRetrieverOddBin (:182-215): seven edges over three bins, each exactly once
    g = RefGraph(reflib, [pair(2 * i + 1, 10, 20, 2 * i + 2, 30, 40) for i in range(7)])
    ours = [e for b in range(3) for e in g.our_edges(lib, 0, 3, b)[1]]
    assert ours == [e for b in range(3) for e in g.ref_edges(0, 3, b)] and len(ours) == 7
    g.close()
    # test_EdgeRetrieverOddBinSelfEdge (:217-): self edges after mergeSelfOverlap
    def selfpair(t):
        return ([(t, 10, 20), (t, 10, 20)], [(0, 1, 1, 1)], True)
    g = RefGraph(reflib, [selfpair(1), selfpair(3), selfpair(5), pair(7, 10, 20, 8, 30, 40), selfpair(9), pair(11, 10, 20, 12, 30, 40),
                          pair(13, 10, 20, 14, 30, 40)])
    ours = [e for b in range(3) for e in g.our_edges(lib, 0, 3, b)[1]]
    assert ours == [e for b in range(3) for e in g.ref_edges(0, 3, b)] and len(ours) == 7
    g.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_locus_sets(lib, reflib, seed):
    """random pair evidence over a small genome: overlapping intervals merge into multi-node loci, repeated pairs raise the
    counts, a few hub regions collect many partners (the noisy-node filter), some pairs are local (self edges)"""
    rng = random.Random(seed)
    specs = []
    hubs = [(rng.randrange(3), rng.randrange(0, 4000, 100)) for _ in range(3)]
    for _ in range(rng.randrange(20, 160)):
        if rng.random() < 0.25:
            t1, b1 = rng.choice(hubs)
        else:
            t1, b1 = rng.randrange(4), rng.randrange(0, 5000, 10)
        e1 = b1 + rng.randrange(5, 60)
        if rng.random() < 0.15:  # local pair: both ends in the same region
            specs.append(([(t1, b1, e1), (t1, b1, e1)], [(0, 1, rng.randrange(1, 4), rng.randrange(1, 4))], True))
            continue
        t2, b2 = rng.randrange(4), rng.randrange(0, 5000, 10)
        e2 = b2 + rng.randrange(5, 60)
        nodes, links, _ = pair(t1, b1, e1, t2, b2, e2, both=rng.random() < 0.3, count=rng.randrange(1, 5))
        # a locus whose own two regions overlap is merged with itself before it enters the set (as the graph builder does)
        specs.append((nodes, links, t1 == t2 and b1 < e2 and b2 < e1))
    g = RefGraph(reflib, specs)
    assert g.total == int(g.edgeCount[:int(g.nodeEdgeOff[-1])].sum())
    n_edges = 0
    for maxEdges in (0, 1, 2, 3, 10):
        for binCount in (1, 2, 3, 7, 16, 64, 257):
            got, want = [], []
            for b in range(binCount):
                st, e, n = g.our_edges(lib, maxEdges, binCount, b)
                assert st == 0 and n == len(e)
                r = g.ref_edges(maxEdges, binCount, b)
                assert e == r, (maxEdges, binCount, b)
                got += e
            # without the filter the bins partition the edges: every edge once, in graph order
            if maxEdges == 0:
                assert got == g.ref_edges(0, 1, 0)
            n_edges = max(n_edges, len(got))
    assert n_edges >= 10
    g.close()


def test_bad_arguments_and_small_buffer(lib, reflib):
    g = RefGraph(reflib, [pair(2 * i + 1, 10, 20, 2 * i + 2, 30, 40) for i in range(5)])
    st, e, n = g.our_edges(lib, 0, 1, 0, cap=2)
    assert st != 0 and n == 5 and e == [(0, 0, 1), (1, 0, 1)]   # MB200_ERR_POOL_OVERFLOW: the count is still reported
    assert g.our_edges(lib, 0, 0, 0)[0] != 0 and g.our_edges(lib, 0, 3, 3)[0] != 0
    g.close()
