// ref_graph_shim.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libmanta_ref.so, built by oracle/build_ref.sh).
//
// The reference's own SVLocusSet / EdgeRetrieverBin (lib/svgraph/SVLocusSet.cpp, GSC/EdgeRetrieverBin.cpp, compiled where
// they lie) behind a C interface: build a locus set the way the reference's unit tests do (SVLocus::addNode / linkNodes,
// optional mergeSelfOverlap, SVLocusSet::merge), export it through the product's reference-side binding
// mb200::flattenLocusSet, and list the edges the reference retriever returns for one bin.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "applications/GenerateSVCandidates/EdgeRetrieverBin.hpp"
#include "svgraph/SVLocusSet.hpp"

#include "manta_b200_graph.hpp"

#define REF_API extern "C" __attribute__((visibility("default")))

struct RefGraph {
  SVLocusSet            set;
  mb200::FlatLocusGraph flat;
  explicit RefGraph(const SVLocusSetOptions& o) : set(o) {}
};

// One locus per spec.  nodes: (tid, begin, end) triples; links: (from, to, count, remoteCount) with node indices local to
// the spec (SVLocus::linkNodes); selfMerge[s] != 0 -> mergeSelfOverlap() before the set merge (EdgeRetrieverBinTest.cpp:239).
REF_API int ref_graph_build(uint32_t nSpecs, const uint32_t* nodeOff, const int32_t* nodes, const uint32_t* linkOff,
                            const uint32_t* links, const uint8_t* selfMerge, uint32_t minMergeEdgeObservations, void** handle)
{
  try {
    SVLocusSetOptions sopt;
    sopt.minMergeEdgeObservations = minMergeEdgeObservations;
    RefGraph* G = new RefGraph(sopt);
    for (uint32_t s = 0; s < nSpecs; ++s) {
      SVLocus                    locus;
      std::vector<NodeIndexType> ids;
      for (uint32_t n = nodeOff[s]; n < nodeOff[s + 1]; ++n)
        ids.push_back(locus.addNode(GenomeInterval(nodes[3 * n], nodes[3 * n + 1], nodes[3 * n + 2])));
      for (uint32_t l = linkOff[s]; l < linkOff[s + 1]; ++l)
        locus.linkNodes(ids[links[4 * l]], ids[links[4 * l + 1]], links[4 * l + 2], links[4 * l + 3]);
      if (selfMerge && selfMerge[s]) locus.mergeSelfOverlap();
      G->set.merge(locus);
    }
    G->set.checkState(true, true);
    G->flat = mb200::flattenLocusSet(G->set);
    *handle = G;
    return 0;
  } catch (const std::exception& e) {
    if (getenv("REF_GRAPH_VERBOSE")) fprintf(stderr, "ref_graph_build: %s\n", e.what());
    return 1;
  } catch (...) {
    return 1;
  }
}

REF_API void ref_graph_free(void* handle) { delete static_cast<RefGraph*>(handle); }

REF_API void ref_graph_sizes(void* handle, uint32_t* nLoci, uint32_t* nNodes, uint64_t* nOut, uint32_t* totalObservations)
{
  const RefGraph* G  = static_cast<const RefGraph*>(handle);
  *nLoci             = G->flat.lociCount();
  *nNodes            = G->flat.locusNodeOff.back();
  *nOut              = G->flat.edgeTo.size();
  *totalObservations = G->set.totalObservationCount();
}

REF_API void ref_graph_flat(void* handle, uint32_t* locusNodeOff, uint64_t* nodeEdgeOff, uint32_t* edgeTo, uint32_t* edgeCount)
{
  const mb200::FlatLocusGraph& f = static_cast<const RefGraph*>(handle)->flat;
  memcpy(locusNodeOff, f.locusNodeOff.data(), sizeof(uint32_t) * f.locusNodeOff.size());
  memcpy(nodeEdgeOff, f.nodeEdgeOff.data(), sizeof(uint64_t) * f.nodeEdgeOff.size());
  if (!f.edgeTo.empty()) {
    memcpy(edgeTo, f.edgeTo.data(), sizeof(uint32_t) * f.edgeTo.size());
    memcpy(edgeCount, f.edgeCount.data(), sizeof(uint32_t) * f.edgeCount.size());
  }
}

// the reference retriever's edges of one bin: out[3*i..] = (locusIndex, nodeIndex1, nodeIndex2); returns the count
REF_API int64_t ref_graph_bin_edges(void* handle, uint32_t graphNodeMaxEdgeCount, uint32_t binCount, uint32_t binIndex, uint32_t* out,
                                    uint64_t cap)
{
  const RefGraph*  G = static_cast<const RefGraph*>(handle);
  EdgeRetrieverBin edger(G->set, graphNodeMaxEdgeCount, binCount, binIndex);
  uint64_t         n = 0;
  while (edger.next()) {
    if (n < cap) {
      const EdgeInfo& e = edger.getEdge();
      out[3 * n] = e.locusIndex; out[3 * n + 1] = e.nodeIndex1; out[3 * n + 2] = e.nodeIndex2;
    }
    ++n;
  }
  return (int64_t)n;
}
