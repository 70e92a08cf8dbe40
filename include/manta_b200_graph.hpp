// manta_b200_graph.hpp -- the SV locus graph as the B200 host side sees it, and the per-bin edge partition.
//
// Row a1 / a14 of SURVEY.md section 8: GenerateSVCandidates walks `EdgeRetrieverBin(set, graphNodeMaxEdgeCount, binCount,
// binIndex)` (GSC/EdgeRetrieverBin.cpp:43-235) over an `SVLocusSet` (lib/svgraph/SVLocusSet.hpp:90) and hands every edge to
// the per-candidate work.  The multi-GPU driver needs the same partition on every rank (rank r takes the bins
// b = r mod G of a 16*G-bin split, DESIGN.md "multi-GPU"), so the graph is flattened ONCE into CSR arrays (plain
// integers: they can be broadcast, checksummed and shipped through the C ABI) and the bins are enumerated on those.
//
//   * FlatLocusGraph        -- CSR copy of the counts the retriever reads: per locus its nodes, per node its out-edge map
//                              (target node index ascending == std::map order, lib/svgraph/SVLocusNode.hpp) with the
//                              directed observation counts.
//   * flattenLocusSet(set)  -- the reference-side binding: a template over the reference's own SVLocusSet accessors
//                              (size(), getLocus(i).size(), getNode(j).getEdgeManager().getMap()), so this header does
//                              not include reference headers.  INTEGRATION.md section 2 shows the call site.
//   * EdgeRetrieverBin      -- same constructor arguments, next()/getEdge() protocol and results as the reference
//                              class, on a FlatLocusGraph.  Pinned to the reference class (compiled from its sources,
//                              oracle/_ref) on the reference's own unit-test graphs and on random graphs, every bin
//                              count, with and without the noisy-node filter (tests/test_edge_bins.py).
//
// Host C++ only (the reference's language); no device code here.
#pragma once

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace mb200 {

/// lib/svgraph/EdgeInfo.hpp:30-39
struct EdgeInfo {
  uint32_t locusIndex = 0, nodeIndex1 = 0, nodeIndex2 = 0;
};

struct FlatLocusGraph {
  std::vector<uint32_t> locusNodeOff;  ///< [nLoci + 1] first node of each locus
  std::vector<uint64_t> nodeEdgeOff;   ///< [nNodes + 1] first out-edge of each node
  std::vector<uint32_t> edgeTo;        ///< target node index INSIDE the locus, ascending per node
  std::vector<uint32_t> edgeCount;     ///< observations on the directed edge node -> edgeTo

  uint32_t lociCount() const { return locusNodeOff.empty() ? 0u : (uint32_t)(locusNodeOff.size() - 1); }
  uint32_t locusSize(uint32_t locus) const { return locusNodeOff[locus + 1] - locusNodeOff[locus]; }
  /// SVLocusNode::size(): entries of the node's edge map
  uint32_t nodeEdges(uint32_t locus, uint32_t node) const
  {
    const uint32_t g = locusNodeOff[locus] + node;
    return (uint32_t)(nodeEdgeOff[g + 1] - nodeEdgeOff[g]);
  }
  /// SVLocus::getEdge(from, to).getCount(); the reference throws when the edge is missing (SVLocusNode.hpp getEdge)
  uint32_t directedCount(uint32_t locus, uint32_t from, uint32_t to) const
  {
    const uint32_t g  = locusNodeOff[locus] + from;
    uint64_t       lo = nodeEdgeOff[g], hi = nodeEdgeOff[g + 1];
    while (lo < hi) {
      const uint64_t mid = (lo + hi) / 2;
      if (edgeTo[mid] < to) lo = mid + 1;
      else hi = mid;
    }
    if (lo == nodeEdgeOff[g + 1] || edgeTo[lo] != to) throw std::runtime_error("SVLocusNode::getEdge: no edge between the nodes");
    return edgeCount[lo];
  }
  /// SVLocus::totalObservationCount() (lib/svgraph/SVLocus.hpp:148): all directed counts of the locus, `unsigned` sum
  uint32_t locusObservations(uint32_t locus) const
  {
    uint32_t sum = 0;
    for (uint64_t e = nodeEdgeOff[locusNodeOff[locus]]; e < nodeEdgeOff[locusNodeOff[locus + 1]]; ++e) sum += edgeCount[e];
    return sum;
  }
  /// SVLocusSet::totalObservationCount() (lib/svgraph/SVLocusSet.hpp:187)
  uint32_t totalObservations() const
  {
    uint32_t sum = 0;
    for (uint32_t c : edgeCount) sum += c;
    return sum;
  }
};

/// CSR copy of an SVLocusSet (any type with the reference's accessors).
template <class LocusSet> FlatLocusGraph flattenLocusSet(const LocusSet& set)
{
  FlatLocusGraph g;
  const unsigned nLoci = set.size();
  g.locusNodeOff.reserve(nLoci + 1);
  g.locusNodeOff.push_back(0);
  g.nodeEdgeOff.push_back(0);
  for (unsigned li = 0; li < nLoci; ++li) {
    const auto&    locus  = set.getLocus(li);
    const unsigned nNodes = locus.size();
    for (unsigned ni = 0; ni < nNodes; ++ni) {
      const auto manager = locus.getNode(ni).getEdgeManager();
      for (const auto& kv : manager.getMap()) {
        g.edgeTo.push_back((uint32_t)kv.first);
        g.edgeCount.push_back((uint32_t)kv.second.getCount());
      }
      g.nodeEdgeOff.push_back(g.edgeTo.size());
    }
    g.locusNodeOff.push_back(g.locusNodeOff.back() + nNodes);
  }
  return g;
}

/// GSC/EdgeRetrieverBin.cpp.  Edges are visited in (locus, node1, node2 >= node1) order; an edge weighs the observations
/// of both directions; bin `binIndex` owns the edges whose running total ends in (begin, end], begin/end =
/// floor(total * index / binCount) computed in double and truncated to `unsigned` (:38-41).  With
/// graphNodeMaxEdgeCount > 0 an edge between two nodes that both have more edges than that is skipped (:148-210), but
/// still counted.
class EdgeRetrieverBin {
public:
  EdgeRetrieverBin(const FlatLocusGraph& graph, unsigned graphNodeMaxEdgeCount, unsigned binCount, unsigned binIndex)
    : _g(graph), _maxEdges(graphNodeMaxEdgeCount)
  {
    if (binCount == 0 || binIndex >= binCount) throw std::invalid_argument("EdgeRetrieverBin: binIndex must be < binCount");
    const double total = (double)_g.totalObservations();
    _begin = (unsigned long)static_cast<unsigned>(std::floor((total * (double)binIndex) / (double)binCount));
    _end   = (unsigned long)static_cast<unsigned>(std::floor((total * ((double)binIndex + 1.0)) / (double)binCount));
  }

  const EdgeInfo& getEdge() const { return _edge; }

  /// advance to the bin's next edge; false when the bin is exhausted (:212-235)
  bool next()
  {
    if (_head >= _end) return false;
    seek(/*catchUp=*/_head < _begin);
    return _head <= _end;
  }

private:
  bool noisy(uint32_t locus, uint32_t node) const { return _g.nodeEdges(locus, node) > _maxEdges; }

  // One routine for both of the reference's loops (jumpToFirstEdge :63-146, advanceEdge :148-210).  They differ in three
  // places: catching up may skip a whole locus by its observation total, it restarts each node's edge scan at the node's
  // own index instead of the last edge's target, and it only stops on an edge once the running total has passed `begin`.
  void seek(const bool catchUp)
  {
    const uint32_t nLoci  = _g.lociCount();
    const bool     filter = (_maxEdges > 0);
    bool           lastFiltered = false;
    if (!catchUp && _head != 0) _edge.nodeIndex2++;
    for (;;) {
      if (lastFiltered && _edge.locusIndex == nLoci) {  // ran off the graph over filtered edges: close the bin
        _head = _end + 1;
        _edge = EdgeInfo();
        return;
      }
      if (_edge.locusIndex >= nLoci) throw std::logic_error("EdgeRetrieverBin: walked past the last locus");
      const uint32_t L = _edge.locusIndex;
      bool           scan = true;
      if (catchUp) {
        const uint32_t locusTotal = _g.locusObservations(L);
        if (_head + locusTotal <= _begin) {
          _head += locusTotal;
          scan = false;
        }
      }
      if (scan) {
        const uint32_t nNodes = _g.locusSize(L);
        while (_edge.nodeIndex1 < nNodes) {
          const uint32_t n1     = _edge.nodeIndex1;
          const bool     noisy1 = filter && noisy(L, n1);
          const uint32_t gnode  = _g.locusNodeOff[L] + n1;
          const uint32_t from   = catchUp ? n1 : _edge.nodeIndex2;  // std::map::lower_bound
          uint64_t       e      = _g.nodeEdgeOff[gnode];
          const uint64_t eEnd   = _g.nodeEdgeOff[gnode + 1];
          while (e < eEnd && _g.edgeTo[e] < from) ++e;
          for (; e < eEnd; ++e) {
            const uint32_t n2 = _g.edgeTo[e];
            uint32_t       w  = _g.edgeCount[e];
            if (n2 != n1) w += _g.directedCount(L, n2, n1);
            _head += w;
            if (catchUp) {
              lastFiltered = false;
              if (!(_head > _begin)) continue;
            }
            _edge.nodeIndex2 = n2;
            if (noisy1 && noisy(L, n2)) {
              lastFiltered = true;
              continue;
            }
            return;
          }
          _edge.nodeIndex1++;
          if (!catchUp) _edge.nodeIndex2 = _edge.nodeIndex1;
        }
        if (catchUp && !(lastFiltered && _head >= _begin)) throw std::logic_error("EdgeRetrieverBin: first edge of the bin not found");
      }
      _edge.locusIndex++;
      _edge.nodeIndex1 = 0;
      _edge.nodeIndex2 = 0;
    }
  }

  const FlatLocusGraph& _g;
  unsigned              _maxEdges;
  unsigned long         _begin = 0, _end = 0, _head = 0;
  EdgeInfo              _edge;
};

/// every edge of one bin, in the retriever's order
inline std::vector<EdgeInfo> enumerateBin(const FlatLocusGraph& g, unsigned graphNodeMaxEdgeCount, unsigned binCount, unsigned binIndex)
{
  std::vector<EdgeInfo> out;
  EdgeRetrieverBin      r(g, graphNodeMaxEdgeCount, binCount, binIndex);
  while (r.next()) out.push_back(r.getEdge());
  return out;
}

}  // namespace mb200
