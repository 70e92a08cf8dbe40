// ref_trace.hpp -- TEST INFRASTRUCTURE ONLY.  Call recorder compiled INTO the reference's own GenerateSVCandidates
// (oracle/build_ref_full.sh --trace applies oracle/trace_calls.patch to a scratch copy of the reference sources): after each
// call of runIterativeAssembler, GlobalJumpAligner::align, GlobalLargeIndelAligner::align, GlobalAligner::align and
// splitReadAligner it appends one JSON line with the inputs and the reference's results to $MANTA_TRACE_CALLS.
// tests/golden/make_demo_calls.py turns the log of the reference's demo into tests/golden/demo_calls.json -- golden vectors
// on real reads, produced by the unmodified reference algorithms.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sstream>
#include <string>

namespace reftrace {

inline FILE* sink()
{
  static FILE* f = []() -> FILE* {
    const char* p = getenv("MANTA_TRACE_CALLS");
    return p ? fopen(p, "a") : nullptr;
  }();
  return f;
}
inline std::mutex& lock()
{
  static std::mutex m;
  return m;
}
inline void emit(const std::string& line)
{
  FILE* f = sink();
  if (!f) return;
  std::lock_guard<std::mutex> g(lock());
  fputs(line.c_str(), f);
  fputc('\n', f);
  fflush(f);
}
template <class It> std::string str(It b, It e) { return std::string(b, e); }
template <class Scores> std::string scores(const Scores& s)
{
  std::ostringstream o;
  o << "[" << s.match << "," << s.mismatch << "," << s.open << "," << s.extend << "," << s.offEdge << "," << (s.isAllowEdgeInsertion ? 1 : 0) << "]";
  return o.str();
}
template <class Aln> std::string aln(const Aln& a)
{
  std::ostringstream o;
  o << "{\"beginPos\":" << a.beginPos << ",\"cigar\":\"" << apath_to_cigar(a.apath) << "\"}";
  return o.str();
}
inline uint32_t bits(const float f)
{
  uint32_t b;
  memcpy(&b, &f, 4);
  return b;
}

template <class Aligner, class It, class Res>
void jump(const Aligner& a, It qb, It qe, It r1b, It r1e, It r2b, It r2e, const Res& r)
{
  if (!sink()) return;
  std::ostringstream o;
  o << "{\"kind\":\"jump\",\"scores\":" << scores(a.getScores()) << ",\"extra\":" << a.getJumpScore() << ",\"seqs\":[\"" << str(qb, qe) << "\",\""
    << str(r1b, r1e) << "\",\"" << str(r2b, r2e) << "\"],\"reference\":{\"score\":" << r.score << ",\"jumpInsertSize\":" << r.jumpInsertSize
    << ",\"jumpRange\":" << r.jumpRange << ",\"align1\":" << aln(r.align1) << ",\"align2\":" << aln(r.align2) << "}}";
  emit(o.str());
}
template <class Aligner, class It, class Res>
void single(const char* kind, const Aligner& a, const int extra, It qb, It qe, It rb, It re, const Res& r)
{
  if (!sink()) return;
  std::ostringstream o;
  o << "{\"kind\":\"" << kind << "\",\"scores\":" << scores(a.getScores()) << ",\"extra\":" << extra << ",\"seqs\":[\"" << str(qb, qe) << "\",\""
    << str(rb, re) << "\"],\"reference\":{\"score\":" << r.score << ",\"isJumped\":" << (r.isJumped ? "true" : "false")
    << ",\"align\":" << aln(r.align) << "}}";
  emit(o.str());
}
template <class Opt, class Reads, class Info, class Asm>
void assembly(const Opt& opt, const Reads& reads, const size_t normalReadCount, const Info& readInfo, const Asm& contigs)
{
  if (!sink()) return;
  std::ostringstream o;
  o << "{\"kind\":\"asm\",\"opts\":{\"minWordLength\":" << opt.minWordLength << ",\"maxWordLength\":" << opt.maxWordLength << ",\"wordStepSize\":"
    << opt.wordStepSize << ",\"minCoverage\":" << opt.minCoverage << ",\"minConservativeCoverage\":" << opt.minConservativeCoverage
    << ",\"maxAssemblyCount\":" << opt.maxAssemblyCount << ",\"minUnusedReads\":" << opt.minUnusedReads << ",\"minSupportReads\":"
    << opt.minSupportReads << "},\"reads\":[";
  for (size_t i(0); i < normalReadCount; ++i) o << (i ? "," : "") << "\"" << reads[i] << "\"";
  o << "],\"reference\":{\"contigs\":[";
  bool first(true);
  for (const auto& c : contigs) {
    o << (first ? "" : ",") << "{\"seq\":\"" << c.seq << "\",\"seedReadCount\":" << c.seedReadCount << ",\"supportReads\":[";
    first = false;
    bool f2(true);
    for (const unsigned r : c.supportReads) { o << (f2 ? "" : ",") << r; f2 = false; }
    o << "],\"rejectReads\":[";
    f2 = true;
    for (const unsigned r : c.rejectReads) { o << (f2 ? "" : ",") << r; f2 = false; }
    o << "],\"conservativeRange\":[" << c.conservativeRange.begin_pos() << "," << c.conservativeRange.end_pos() << "]}";
  }
  o << "],\"pseudoReads\":[";
  for (size_t i(normalReadCount); i < reads.size(); ++i) o << (i > normalReadCount ? "," : "") << "\"" << reads[i] << "\"";
  o << "],\"readInfo\":[";
  for (size_t i(0); i < readInfo.size(); ++i) {
    o << (i ? "," : "") << "{\"isUsed\":" << (readInfo[i].isUsed ? "true" : "false") << ",\"isPseudo\":" << (readInfo[i].isPseudo ? "true" : "false")
      << ",\"contigIds\":[";
    for (size_t k(0); k < readInfo[i].contigIds.size(); ++k) o << (k ? "," : "") << readInfo[i].contigIds[k];
    o << "]}";
  }
  o << "]}}";
  emit(o.str());
}
template <class Range, class SR>
void split(const unsigned flank, const std::string& query, const int table, const uint8_t* qual, const std::string& target, const Range& range,
           const SR& r)
{
  if (!sink()) return;
  std::ostringstream o;
  o << "{\"kind\":\"split\",\"flank\":" << flank << ",\"query\":\"" << query << "\",\"qual\":[";
  for (size_t i(0); i < query.size(); ++i) o << (i ? "," : "") << static_cast<unsigned>(qual[i]);
  o << "],\"target\":\"" << target << "\",\"bp\":[" << range.begin_pos() << "," << range.end_pos() << "],\"table\":" << table
    << ",\"reference\":{\"alignPos\":" << r.alignPos << ",\"leftSize\":" << r.leftSize << ",\"homSize\":" << r.homSize << ",\"rightSize\":" << r.rightSize
    << ",\"leftMismatches\":" << r.leftMismatches << ",\"homMismatches\":" << r.homMismatches << ",\"rightMismatches\":" << r.rightMismatches
    << ",\"alignScore\":" << r.alignScore << ",\"isEvidence\":" << (r.isEvidence ? 1 : 0) << ",\"isTier2Evidence\":" << (r.isTier2Evidence ? 1 : 0)
    << ",\"status\":0,\"alignLnLhood_bits\":" << bits(r.alignLnLhood) << ",\"evidence_bits\":" << bits(r.evidence) << "}}";
  emit(o.str());
}

}  // namespace reftrace
