// edge_bin.cu -- C ABI of the per-bin edge partition (row a1).  Host code only: mb200::EdgeRetrieverBin
// (include/manta_b200_graph.hpp) on the caller's flattened SVLocusSet.
#include <cstdint>
#include <exception>

#include "manta_b200.h"
#include "manta_b200_graph.hpp"

extern "C" int mb200_edge_bin_enumerate(uint32_t nLoci, const uint32_t* locusNodeOff, const uint64_t* nodeEdgeOff,
                                        const uint32_t* edgeTo, const uint32_t* edgeCount, uint32_t graphNodeMaxEdgeCount,
                                        uint32_t binCount, uint32_t binIndex, mb200_edge_info* edges, uint64_t edgeCap,
                                        uint64_t* nEdges)
{
  if (!locusNodeOff || !nodeEdgeOff || !nEdges || (edgeCap && !edges) || binCount == 0 || binIndex >= binCount) return MB200_ERR_BAD_ARG;
  *nEdges = 0;
  const uint32_t nNodes = locusNodeOff[nLoci];
  const uint64_t nOut   = nodeEdgeOff[nNodes];
  if (nOut && (!edgeTo || !edgeCount)) return MB200_ERR_BAD_ARG;
  for (uint32_t l = 0; l < nLoci; ++l)
    if (locusNodeOff[l + 1] < locusNodeOff[l]) return MB200_ERR_BAD_ARG;
  for (uint32_t n = 0; n < nNodes; ++n)
    if (nodeEdgeOff[n + 1] < nodeEdgeOff[n]) return MB200_ERR_BAD_ARG;
  mb200::FlatLocusGraph g;
  g.locusNodeOff.assign(locusNodeOff, locusNodeOff + nLoci + 1);
  g.nodeEdgeOff.assign(nodeEdgeOff, nodeEdgeOff + nNodes + 1);
  g.edgeTo.assign(edgeTo, edgeTo + nOut);
  g.edgeCount.assign(edgeCount, edgeCount + nOut);
  try {
    mb200::EdgeRetrieverBin r(g, graphNodeMaxEdgeCount, binCount, binIndex);
    uint64_t                n = 0;
    while (r.next()) {
      if (n < edgeCap) {
        const mb200::EdgeInfo& e = r.getEdge();
        edges[n].locusIndex      = e.locusIndex;
        edges[n].nodeIndex1      = e.nodeIndex1;
        edges[n].nodeIndex2      = e.nodeIndex2;
      }
      ++n;
    }
    *nEdges = n;
    return n > edgeCap ? MB200_ERR_POOL_OVERFLOW : MB200_OK;
  } catch (const std::exception&) {
    return MB200_ERR_BAD_ARG;  // asymmetric edge maps / counts that do not add up: not a graph the reference can hold
  }
}
