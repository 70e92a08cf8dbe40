// mb200::SVAlignmentInfo (include/manta_b200_shim.hpp, row a10) against the reference's own class
// (lib/manta/SVScoreInfo.cpp:31-167) through oracle/_ref/libmanta_ref.so.  Host arithmetic only: no GPU, no ctx.
// usage: align_info_test <path to libmanta_ref.so>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "manta_b200_shim.hpp"

typedef int (*small_fn)(const char*, const char*, int32_t, const uint32_t*, uint32_t, uint32_t, uint32_t, int32_t, int32_t, int32_t, int32_t,
                        uint32_t, int32_t*, char*, char*, uint32_t);
typedef int (*jump_fn)(const char*, const char*, const char*, int32_t, uint32_t, uint32_t, int32_t, const uint32_t*, uint32_t, int32_t,
                       const uint32_t*, uint32_t, const mb200_jump_orient*, int32_t, int32_t, int32_t, int32_t, uint32_t, int32_t*, char*,
                       char*, uint32_t);

static uint32_t rngState = 2463534242u;
static uint32_t rnd(uint32_t n)
{
  rngState ^= rngState << 13; rngState ^= rngState >> 17; rngState ^= rngState << 5;
  return rngState % n;
}
static std::string randSeq(unsigned n)
{
  std::string s;
  for (unsigned i = 0; i < n; ++i) s.push_back("ACGTN"[rnd(40) == 0 ? 4 : rnd(4)]);
  return s;
}
static mb200::path_t randPath(unsigned& readLen, unsigned& refLen, unsigned& segFirst, unsigned& segLast)
{
  mb200::path_t p;
  p.push_back({MB200_SEG_SEQ_MATCH, 20 + rnd(80)});
  segFirst = static_cast<unsigned>(p.size());
  for (unsigned k = 0, n = 1 + rnd(2); k < n; ++k) p.push_back({rnd(2) ? MB200_SEG_INSERT : MB200_SEG_DELETE, 10 + rnd(60)});
  // adjacent segments of one type never occur in real paths; merge
  if (p.size() == segFirst + 2 && p[segFirst].type == p[segFirst + 1].type) { p[segFirst].length += p.back().length; p.pop_back(); }
  segLast = static_cast<unsigned>(p.size()) - 1;
  p.push_back({MB200_SEG_SEQ_MATCH, 20 + rnd(80)});
  if (rnd(3) == 0) p.push_back({MB200_SEG_SOFT_CLIP, 1 + rnd(10)});
  readLen = refLen = 0;
  for (const auto& s : p) {
    if (s.type != MB200_SEG_DELETE) readLen += s.length;
    if (s.type == MB200_SEG_SEQ_MATCH || s.type == MB200_SEG_DELETE) refLen += s.length;
  }
  return p;
}
static std::vector<uint32_t> pack(const mb200::path_t& p)
{
  std::vector<uint32_t> v;
  for (const auto& s : p) v.push_back(MB200_SEG_MAKE(s.length, s.type));
  return v;
}

#define REQUIRE(x)                                                          \
  do {                                                                      \
    if (!(x)) {                                                             \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #x);   \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { std::fprintf(stderr, "%s\n", dlerror()); return 2; }
  small_fn refSmall = reinterpret_cast<small_fn>(dlsym(h, "ref_sv_align_info_small"));
  jump_fn  refJump  = reinterpret_cast<jump_fn>(dlsym(h, "ref_sv_align_info_jump"));
  REQUIRE(refSmall && refJump);
  std::vector<char> cbuf(1 << 16), rbuf(1 << 16);
  int32_t           out[9];
  unsigned          nMinEdgeTrue = 0;
  for (int it = 0; it < 400; ++it) {
    // ---- complex candidate
    unsigned readLen, refLen, f, l;
    const mb200::path_t path(randPath(readLen, refLen, f, l));
    const int           beginPos(rnd(60));
    const std::string   ref(randSeq(beginPos + refLen + rnd(60))), contig(randSeq(readLen));
    mb200::AlignmentResult<int> aln;
    aln.align.beginPos = beginPos;
    aln.align.apath    = path;
    mb200_small_sv sv  = {};
    sv.pathFirst = f; sv.pathLast = l;
    const int hom(rnd(6));
    sv.bp1Begin = 5 + rnd(100); sv.bp1End = sv.bp1Begin + hom + 1;
    sv.bp2Begin = sv.bp1Begin + rnd(80); sv.bp2End = sv.bp2Begin + hom + 1;
    const unsigned minEdge(rnd(40));
    const std::vector<uint32_t> pk(pack(path));
    REQUIRE(refSmall(contig.c_str(), ref.c_str(), beginPos, pk.data(), pk.size(), f, l, sv.bp1Begin, sv.bp1End, sv.bp2Begin, sv.bp2End,
                     minEdge, out, cbuf.data(), rbuf.data(), cbuf.size()) == 0);
    const mb200::SVAlignmentInfo ai(sv, aln, contig, ref);
    REQUIRE(ai.bp1ContigOffset.begin == out[0] && ai.bp1ContigOffset.end == out[1] && ai.bp2ContigOffset.begin == out[2] && ai.bp2ContigOffset.end == out[3]);
    REQUIRE(ai.bp1RefOffset.begin == out[4] && ai.bp1RefOffset.end == out[5] && ai.bp2RefOffset.begin == out[6] && ai.bp2RefOffset.end == out[7]);
    REQUIRE(ai.bp1ContigSeq() == cbuf.data() && ai.bp2ContigSeq() == cbuf.data() && !ai.isSpanning());
    REQUIRE((ai.isMinBpEdge(minEdge) ? 1 : 0) == out[8]);
    nMinEdgeTrue += out[8];
    // ---- spanning candidate
    unsigned r1, g1, r2, g2, d0, d1;
    mb200::JumpAlignmentResult<int> ja;
    ja.align1.apath    = randPath(r1, g1, d0, d1);
    ja.align2.apath    = randPath(r2, g2, d0, d1);
    ja.align1.beginPos = rnd(60);
    ja.align2.beginPos = rnd(60);
    ja.jumpInsertSize  = rnd(3) ? 0 : rnd(30);
    ja.jumpRange       = rnd(2) ? 0 : rnd(8);
    ja.score           = 100;
    const std::string aref1(randSeq(ja.align1.beginPos + g1 + rnd(60))), aref2(randSeq(ja.align2.beginPos + g2 + rnd(60)));
    const std::string jcontig(randSeq(r1 + ja.jumpInsertSize + r2));
    mb200_jump_orient o = {};
    o.isBp2AlignedFirst = rnd(2);
    const unsigned revPick(rnd(3));  // never both (the reference asserts that)
    o.isBp1Reversed = (revPick == 1);
    o.isBp2Reversed = (revPick == 2);
    mb200_jump_sv jsv = {};
    jsv.valid = 1;
    jsv.bp1Begin = 5 + rnd(100); jsv.bp1End = jsv.bp1Begin + ja.jumpRange + 1;
    jsv.bp2Begin = 5 + rnd(100); jsv.bp2End = jsv.bp2Begin + ja.jumpRange + 1;
    const std::vector<uint32_t> p1(pack(ja.align1.apath)), p2(pack(ja.align2.apath));
    REQUIRE(refJump(jcontig.c_str(), aref1.c_str(), aref2.c_str(), ja.score, ja.jumpInsertSize, ja.jumpRange, ja.align1.beginPos, p1.data(),
                    p1.size(), ja.align2.beginPos, p2.data(), p2.size(), &o, jsv.bp1Begin, jsv.bp1End, jsv.bp2Begin, jsv.bp2End, minEdge, out,
                    cbuf.data(), rbuf.data(), cbuf.size()) == 0);
    const mb200::SVAlignmentInfo aj(jsv, ja, o, jcontig, aref1, aref2);
    REQUIRE(aj.bp1ContigOffset.begin == out[0] && aj.bp1ContigOffset.end == out[1] && aj.bp2ContigOffset.begin == out[2] && aj.bp2ContigOffset.end == out[3]);
    REQUIRE(aj.bp1RefOffset.begin == out[4] && aj.bp1RefOffset.end == out[5] && aj.bp2RefOffset.begin == out[6] && aj.bp2RefOffset.end == out[7]);
    REQUIRE((aj.isMinBpEdge(minEdge) ? 1 : 0) == out[8] && aj.isSpanning());
    const std::string& fwd(o.isBp1Reversed ? aj.bp2ContigSeq() : aj.bp1ContigSeq());
    REQUIRE(fwd == cbuf.data());
    if (o.isBp1Reversed) REQUIRE(aj.bp1ContigSeq() == rbuf.data());
    if (o.isBp2Reversed) REQUIRE(aj.bp2ContigSeq() == rbuf.data());
    REQUIRE((o.isBp2AlignedFirst ? aj.bp1ReferenceSeq() == aref2 : aj.bp1ReferenceSeq() == aref1));
  }
  REQUIRE(nMinEdgeTrue > 10 && nMinEdgeTrue < 390);
  std::printf("align_info_test ok\n");
  return 0;
}
