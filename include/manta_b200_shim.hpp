// manta_b200_shim.hpp -- header-only C++ adapter from the reference's in-process call seams
// to the batched C ABI of manta_b200.h.  Each class/function below has the signature, argument meaning and error
// behaviour of the reference entity it replaces, so a Manta maintainer can swap the include and keep the call site:
//
//   mb200::GlobalJumpAligner<int>::align        <- lib/alignment/GlobalJumpAligner.hpp:45-53
//   mb200::GlobalLargeIndelAligner<int>::align  <- lib/alignment/GlobalLargeIndelAligner.hpp:48-54
//   mb200::GlobalAligner<int>::align            <- lib/alignment/GlobalAligner.hpp
//   mb200::splitReadAligner                     <- GSC/SplitReadAlignment.hpp:63-70
//   mb200::runIterativeAssembler                <- lib/assembly/IterativeAssembler.hpp:43-47
//
// The shim is deliberately free of reference headers: it ships minimal look-alike result types (same member names)
// so that this file compiles on its own, and every entry point is a template over the result / option / contig type,
// so the reference's own AlignmentResult<int>, JumpAlignmentResult<int>, SRAlignmentInfo, AssembledContig,
// IterativeAssemblerOptions ... pass straight through (integration/manta_b200.patch does exactly that; the patched
// GenerateSVCandidates reproduces the reference's demo VCFs, tests/test_vcf_demo.py).  Errors of the C ABI are thrown as
// mb200::GeneralException, which the integration maps onto illumina::common::GeneralException (same what() text).
//
// Batch size is 1 per call here (drop-in semantics); the staged driver described in INTEGRATION.md calls the
// batched C entry points directly.
#pragma once

#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "manta_b200.h"

namespace mb200 {

struct GeneralException : public std::runtime_error {
  GeneralException(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
  int status;
};

/// one device context per host thread (the reference keeps one aligner object per worker thread,
/// GSC/GenerateSVCandidates.cpp:232-250)
class Context {
public:
  explicit Context(int device = 0)
  {
    const int st(mb200_create(device, nullptr, &_ctx));
    if (st != MB200_OK) throw GeneralException(st, "mb200_create failed: no usable sm_100 CUDA device (no CPU fallback)");
  }
  ~Context() { mb200_destroy(_ctx); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  mb200_ctx* get() const { return _ctx; }
  void check(int st) const
  {
    if (st != MB200_OK) throw GeneralException(st, mb200_last_error(_ctx));
  }

private:
  mb200_ctx* _ctx = nullptr;
};

/// the calling thread's context, created on first use.  Aligners built without an explicit Context use it at CALL time,
/// so objects constructed on one thread and used by a worker (the refiner's const aligner members) stay race-free.
inline const Context& threadContext()
{
  static thread_local Context ctx(0);
  return ctx;
}

// ---- ALIGNPATH look-alikes (lib/blt_util/align_path.hpp:34-60) ----------------------------------
struct path_segment {
  unsigned type;  // ALIGNPATH::align_t value
  unsigned length;
};
typedef std::vector<path_segment> path_t;

inline std::string apath_to_cigar(const path_t& apath)
{
  static const char codes[] = "?MIDNSHP=X";
  std::string       s;
  for (const path_segment& ps : apath) s += std::to_string(ps.length) + codes[ps.type];
  return s;
}

struct Alignment {  // lib/alignment/Alignment.hpp:28-41
  bool   isAligned() const { return !apath.empty(); }
  int    beginPos = 0;
  path_t apath;
};

template <typename ScoreType>
struct AlignmentScores {  // lib/alignment/AlignmentScores.hpp:23-57
  AlignmentScores(ScoreType m, ScoreType mm, ScoreType o, ScoreType e, ScoreType oe, bool edgeIns = false)
    : match(m), mismatch(mm), open(o), extend(e), offEdge(oe), isAllowEdgeInsertion(edgeIns)
  {
  }
  const ScoreType match, mismatch, open, extend, offEdge;
  const bool      isAllowEdgeInsertion;
  /// from any type with the reference's member names (lib/alignment/AlignmentScores.hpp)
  template <typename S>
  explicit AlignmentScores(const S& s)
    : match(s.match), mismatch(s.mismatch), open(s.open), extend(s.extend), offEdge(s.offEdge), isAllowEdgeInsertion(s.isAllowEdgeInsertion)
  {
  }
  mb200_scores    c() const { return mb200_scores{match, mismatch, open, extend, offEdge, isAllowEdgeInsertion ? 1 : 0}; }
};

template <typename ScoreType>
struct AlignmentResult {  // lib/alignment/SingleRefAlignerShared.hpp:34-47
  ScoreType score    = 0;
  bool      isJumped = false;
  Alignment align;
};

template <typename ScoreType>
struct JumpAlignmentResult {  // lib/alignment/JumpAlignerBase.hpp:37-56
  ScoreType score          = 0;
  unsigned  jumpInsertSize = 0;
  unsigned  jumpRange      = 0;
  Alignment align1, align2;
};

namespace detail {
inline path_t toPath(const uint32_t* segs, uint64_t off, uint32_t n)
{
  path_t p(n);
  for (uint32_t i(0); i < n; ++i) p[i] = path_segment{MB200_SEG_TYPE(segs[off + i]), MB200_SEG_LEN(segs[off + i])};
  return p;
}
template <typename SymIter>
inline mb200_seqview append(std::vector<uint8_t>& pool, SymIter b, SymIter e)
{
  mb200_seqview v{pool.size(), 0, 0};
  for (; b != e; ++b) pool.push_back(static_cast<uint8_t>(*b));
  v.len = static_cast<uint32_t>(pool.size() - v.off);
  while (pool.size() % 16) pool.push_back(0);
  return v;
}
/// fill any path type whose segments have the reference's {type, length} members (ALIGNPATH::path_t, mb200::path_t)
template <typename PathT>
inline void fillPath(PathT& path, const uint32_t* segs, const uint64_t off, const uint32_t n)
{
  typedef typename PathT::value_type Seg;
  path.clear();
  path.reserve(n);
  for (uint32_t i(0); i < n; ++i) {
    Seg s;
    s.type   = static_cast<decltype(s.type)>(MB200_SEG_TYPE(segs[off + i]));
    s.length = MB200_SEG_LEN(segs[off + i]);
    path.push_back(s);
  }
}
}  // namespace detail

/// GlobalJumpAligner<int>::align  lib/alignment/GlobalJumpAligner.hpp:45-53 (throws on empty inputs, Impl:50-58).
/// `ResultT` is any type with JumpAlignmentResult's members (the reference's own or the look-alike above).
template <typename ScoreType>
struct GlobalJumpAligner {
  GlobalJumpAligner(const Context& ctx, const AlignmentScores<ScoreType>& scores, const ScoreType jumpScore)
    : _ctx(&ctx), _scores(scores), _jumpScore(jumpScore)
  {
  }
  /// uses threadContext() of the calling thread
  GlobalJumpAligner(const AlignmentScores<ScoreType>& scores, const ScoreType jumpScore) : _ctx(nullptr), _scores(scores), _jumpScore(jumpScore) {}
  template <typename SymIter, typename ResultT>
  void align(const SymIter queryBegin, const SymIter queryEnd, const SymIter ref1Begin, const SymIter ref1End,
             const SymIter ref2Begin, const SymIter ref2End, ResultT& result) const
  {
    const Context&       ctx(_ctx ? *_ctx : threadContext());
    std::vector<uint8_t> pool;
    const mb200_seqview  q(detail::append(pool, queryBegin, queryEnd));
    const mb200_seqview  r1(detail::append(pool, ref1Begin, ref1End));
    const mb200_seqview  r2(detail::append(pool, ref2Begin, ref2End));
    std::vector<uint32_t> cigar(q.len + r1.len + r2.len + 16);
    mb200_jump_result     res;
    uint64_t              used(0);
    const mb200_scores    sc(_scores.c());
    ctx.check(mb200_jump_align_batch(ctx.get(), &sc, _jumpScore, 1, pool.data(), pool.size(), &q, &r1, &r2, &res,
                                     cigar.data(), cigar.size(), &used));
    result                 = ResultT();
    result.score           = static_cast<ScoreType>(res.score);
    result.jumpInsertSize  = res.jumpInsertSize;
    result.jumpRange       = res.jumpRange;
    result.align1.beginPos = res.beginPos1;
    result.align2.beginPos = res.beginPos2;
    detail::fillPath(result.align1.apath, cigar.data(), res.cigarOff1, res.cigarLen1);
    detail::fillPath(result.align2.apath, cigar.data(), res.cigarOff2, res.cigarLen2);
  }
  const ScoreType&                  getJumpScore() const { return _jumpScore; }
  const AlignmentScores<ScoreType>& getScores() const { return _scores; }

private:
  const Context*                   _ctx;
  const AlignmentScores<ScoreType> _scores;
  const ScoreType                  _jumpScore;
};

/// GlobalLargeIndelAligner<int>::align (lib/alignment/GlobalLargeIndelAligner.hpp:48-54) when constructed with a
/// largeIndelScore, GlobalAligner<int>::align (lib/alignment/GlobalAligner.hpp) otherwise
template <typename ScoreType>
struct SingleRefAligner {
  SingleRefAligner(const Context* ctx, const AlignmentScores<ScoreType>& scores) : _ctx(ctx), _scores(scores), _isIndel(false), _L(0) {}
  SingleRefAligner(const Context* ctx, const AlignmentScores<ScoreType>& scores, const ScoreType largeIndelScore)
    : _ctx(ctx), _scores(scores), _isIndel(true), _L(largeIndelScore)
  {
  }
  template <typename SymIter, typename ResultT>
  void align(const SymIter queryBegin, const SymIter queryEnd, const SymIter refBegin, const SymIter refEnd, ResultT& result) const
  {
    const Context&        ctx(_ctx ? *_ctx : threadContext());
    std::vector<uint8_t>  pool;
    const mb200_seqview   q(detail::append(pool, queryBegin, queryEnd));
    const mb200_seqview   r(detail::append(pool, refBegin, refEnd));
    std::vector<uint32_t> cigar(q.len + r.len + 16);
    mb200_aln_result      res;
    uint64_t              used(0);
    const mb200_scores    sc(_scores.c());
    if (_isIndel)
      ctx.check(mb200_indel_align_batch(ctx.get(), &sc, _L, 1, pool.data(), pool.size(), &q, &r, &res, cigar.data(), cigar.size(), &used));
    else
      ctx.check(mb200_global_align_batch(ctx.get(), &sc, 1, pool.data(), pool.size(), &q, &r, &res, cigar.data(), cigar.size(), &used));
    result                = ResultT();
    result.score          = static_cast<ScoreType>(res.score);
    result.isJumped       = (res.isJumped != 0);
    result.align.beginPos = res.beginPos;
    detail::fillPath(result.align.apath, cigar.data(), res.cigarOff, res.cigarLen);
  }
  const AlignmentScores<ScoreType>& getScores() const { return _scores; }

private:
  const Context*                   _ctx;
  const AlignmentScores<ScoreType> _scores;
  const bool                       _isIndel;
  const ScoreType                  _L;
};
template <typename ScoreType>
struct GlobalLargeIndelAligner : public SingleRefAligner<ScoreType> {
  GlobalLargeIndelAligner(const Context& ctx, const AlignmentScores<ScoreType>& scores, const ScoreType largeIndelScore)
    : SingleRefAligner<ScoreType>(&ctx, scores, largeIndelScore)
  {
  }
  GlobalLargeIndelAligner(const AlignmentScores<ScoreType>& scores, const ScoreType largeIndelScore)
    : SingleRefAligner<ScoreType>(nullptr, scores, largeIndelScore)
  {
  }
};
template <typename ScoreType>
struct GlobalAligner : public SingleRefAligner<ScoreType> {
  GlobalAligner(const Context& ctx, const AlignmentScores<ScoreType>& scores) : SingleRefAligner<ScoreType>(&ctx, scores) {}
  explicit GlobalAligner(const AlignmentScores<ScoreType>& scores) : SingleRefAligner<ScoreType>(nullptr, scores) {}
};

// ---- split-read scan --------------------------------------------------------------------------------
struct SRAlignmentInfo {  // GSC/SplitReadAlignment.hpp:35-49
  unsigned alignPos = 0, leftSize = 0, homSize = 0, rightSize = 0;
  unsigned leftMismatches = 0, homMismatches = 0, rightMismatches = 0, alignScore = 0;
  float    alignLnLhood = 0;
  bool     isEvidence = false, isTier2Evidence = false;
  float    evidence = 0;
};

/// splitReadAligner  GSC/SplitReadAlignment.hpp:63-70.  `lnComp`/`lnErr` are the 71-entry tables of the caller's
/// qscore_snp (qphred_to_ln_comp_error_prob / qphred_to_ln_error_prob, lib/blt_util/qscore_snp.hpp:43-53);
/// targetBpBegin/End are targetBpOffsetRange.begin_pos()/end_pos().
template <typename SRAlignmentInfoT>
inline void splitReadAligner(const Context& ctx, const unsigned flankScoreSize, const std::string& querySeq,
                             const double* lnComp, const double* lnErr, const uint8_t* queryQual,
                             const std::string& targetSeq, const int targetBpBegin, const int targetBpEnd,
                             SRAlignmentInfoT& alignment)
{
  std::vector<uint8_t> pool;
  mb200_sr_job         job;
  std::memset(&job, 0, sizeof(job));
  job.query   = detail::append(pool, querySeq.begin(), querySeq.end());
  job.qualOff = detail::append(pool, queryQual, queryQual + querySeq.size()).off;
  job.target  = detail::append(pool, targetSeq.begin(), targetSeq.end());
  job.bpBegin = targetBpBegin;
  job.bpEnd   = targetBpEnd;
  mb200_qscore_tables t;
  for (int k(0); k < 2; ++k) {
    std::memcpy(t.lnComp[k], lnComp, sizeof(t.lnComp[k]));
    std::memcpy(t.lnErr[k], lnErr, sizeof(t.lnErr[k]));
  }
  mb200_sr_result r;
  ctx.check(mb200_split_align_batch(ctx.get(), &t, flankScoreSize, 1, pool.data(), pool.size(), &job, &r));
  if (r.status == MB200_ERR_SPLIT_INPUT) throw GeneralException(r.status, "Unexpected split read alignment input.");
  if (r.status == MB200_ERR_SPLIT_SCAN) throw GeneralException(r.status, "Unexpected split read alignment input condition: scanEnd < scanStart.");
  if (r.status == MB200_ERR_SPLIT_OUTCOME) throw GeneralException(r.status, "Unexpected split read alignment outcome. ");
  if (r.status != MB200_OK) throw GeneralException(r.status, "basecall quality out of range");
  alignment.alignPos = r.alignPos; alignment.leftSize = r.leftSize; alignment.homSize = r.homSize; alignment.rightSize = r.rightSize;
  alignment.leftMismatches = r.leftMismatches; alignment.homMismatches = r.homMismatches; alignment.rightMismatches = r.rightMismatches;
  alignment.alignScore = r.alignScore; alignment.alignLnLhood = r.alignLnLhood;
  alignment.isEvidence = (r.isEvidence != 0); alignment.isTier2Evidence = (r.isTier2Evidence != 0); alignment.evidence = r.evidence;
}

/// The reference's own argument list (GSC/SplitReadAlignment.hpp:63-70): `qualConvert` is the caller's qscore_snp (any
/// type with qphred_to_ln_comp_error_prob / qphred_to_ln_error_prob), `targetBpOffsetRange` its known_pos_range2.
/// Runs on the calling thread's context.
template <typename QscoreSnpT, typename PosRangeT, typename SRAlignmentInfoT>
inline void splitReadAligner(const unsigned flankScoreSize, const std::string& querySeq, const QscoreSnpT& qualConvert,
                             const uint8_t* queryQual, const std::string& targetSeq, const PosRangeT& targetBpOffsetRange,
                             SRAlignmentInfoT& alignment)
{
  double lnComp[71], lnErr[71];
  for (int q(0); q <= 70; ++q) {
    lnComp[q] = qualConvert.qphred_to_ln_comp_error_prob(q);
    lnErr[q]  = qualConvert.qphred_to_ln_error_prob(q);
  }
  splitReadAligner(threadContext(), flankScoreSize, querySeq, lnComp, lnErr, queryQual, targetSeq,
                   static_cast<int>(targetBpOffsetRange.begin_pos()), static_cast<int>(targetBpOffsetRange.end_pos()), alignment);
}

// ---- assembler ----------------------------------------------------------------------------------------
struct IterativeAssemblerOptions {  // lib/options/IterativeAssemblerOptions.hpp:26-59
  unsigned minWordLength = 41, maxWordLength = 76, wordStepSize = 5, minCoverage = 1, minConservativeCoverage = 2;
  unsigned minUnusedReads = 3, minSupportReads = 2, maxAssemblyCount = 10;
};
struct AssembledContig {  // lib/assembly/AssembledContig.hpp:38-50
  std::string        seq;
  unsigned           seedReadCount = 0;
  std::set<unsigned> supportReads, rejectReads;
  int                conservativeRangeBegin = 0, conservativeRangeEnd = 0;
};
struct AssemblyReadInfo {  // lib/assembly/AssemblyReadInfo.hpp:29-41
  bool                  isUsed = false, isFiltered = false, isPseudo = false;
  std::vector<unsigned> contigIds;
};
typedef std::vector<std::string>      AssemblyReadInput;
typedef std::vector<AssemblyReadInfo> AssemblyReadOutput;
typedef std::vector<AssembledContig>  Assembly;

/// runIterativeAssembler  lib/assembly/IterativeAssembler.hpp:43-47: `reads` is mutated (pseudo reads appended),
/// `readInfo` and `contigs` are filled exactly as the reference leaves them.
namespace detail {
// AssembledContig::conservativeRange is a known_pos_range2 in the reference (lib/assembly/AssembledContig.hpp:49) and
// two ints in the look-alike above
template <typename ContigT>
inline auto setConservativeRange(ContigT& c, const int b, const int e, int) -> decltype(c.conservativeRange.set_range(b, e), void())
{
  c.conservativeRange.set_range(b, e);
}
template <typename ContigT>
inline void setConservativeRange(ContigT& c, const int b, const int e, long)
{
  c.conservativeRangeBegin = b;
  c.conservativeRangeEnd   = e;
}
}  // namespace detail

/// `OptionsT`, the element of `ReadInfoVecT` and the element of `AssemblyT` are the reference's
/// IterativeAssemblerOptions / AssemblyReadInfo / AssembledContig or the look-alikes above.
template <typename OptionsT, typename ReadInfoVecT, typename AssemblyT>
inline void runIterativeAssembler(const Context& ctx, const OptionsT& opt, AssemblyReadInput& reads, ReadInfoVecT& readInfo,
                                  AssemblyT& contigs)
{
  typedef typename ReadInfoVecT::value_type AssemblyReadInfoT;
  typedef typename AssemblyT::value_type    AssembledContigT;
  std::vector<uint8_t>       pool;
  std::vector<mb200_seqview> views;
  for (const std::string& r : reads) views.push_back(detail::append(pool, r.begin(), r.end()));
  pool.resize(pool.size() + 64);
  const uint64_t       readOff[2] = {0, reads.size()};
  const mb200_asm_opts o{opt.minWordLength, opt.maxWordLength, opt.wordStepSize, opt.minCoverage, opt.minConservativeCoverage,
                         opt.maxAssemblyCount, opt.minUnusedReads, opt.minSupportReads};
  const uint32_t             A(opt.maxAssemblyCount);
  mb200_asm_locus            L;
  std::vector<mb200_contig>  cs(A);
  std::vector<mb200_seqview> ps(2 * A);
  size_t                     poolCap(1 << 16), setCap(4096);
  std::vector<uint8_t>       cpool;
  std::vector<uint32_t>      spool;
  for (;;) {
    cpool.resize(poolCap);
    spool.resize(setCap);
    const int st(mb200_assemble_batch(ctx.get(), &o, 1, readOff, views.data(), pool.data(), pool.size(), &L, cs.data(), A,
                                      ps.data(), 2 * A, cpool.data(), poolCap, spool.data(), setCap));
    if (st == MB200_ERR_POOL_OVERFLOW) {
      poolCap *= 4;
      setCap *= 4;
      continue;
    }
    ctx.check(st);
    break;
  }
  const unsigned normalReadCount(reads.size());
  for (uint32_t p(0); p < L.nPseudo; ++p)
    reads.emplace_back(reinterpret_cast<const char*>(cpool.data() + ps[L.firstPseudo + p].off), ps[L.firstPseudo + p].len);
  readInfo.assign(reads.size(), AssemblyReadInfoT());
  for (unsigned r(normalReadCount); r < reads.size(); ++r) readInfo[r].isPseudo = true;
  contigs.clear();
  for (uint32_t c(0); c < L.nContigs; ++c) {
    const mb200_contig& m(cs[L.firstContig + c]);
    AssembledContigT    ac;
    ac.seq.assign(reinterpret_cast<const char*>(cpool.data() + m.seqOff), m.seqLen);
    detail::setConservativeRange(ac, m.conservativeBegin, m.conservativeEnd, 0);
    for (uint32_t w(0); w < L.setWords; ++w) {
      for (uint32_t b(0); b < 32; ++b) {
        if (spool[m.supportOff + w] & (1u << b)) ac.supportReads.insert(w * 32 + b);
        if (spool[m.rejectOff + w] & (1u << b)) ac.rejectReads.insert(w * 32 + b);
      }
    }
    for (const unsigned rd : ac.supportReads) {  // IterativeAssembler.cpp:822-831
      if (rd >= readInfo.size()) continue;       // stale pseudo index after the last word length failed
      readInfo[rd].isUsed = true;
      readInfo[rd].contigIds.push_back(c);
    }
    contigs.push_back(ac);
  }
}

// ---- reference-window trim, contig QC and selection (refiner rows a5 / a9) ------------------------------------------

typedef std::vector<std::pair<unsigned, unsigned>> CandidateSegmentSetType;  // lib/manta/SVCandidateAssemblyData.hpp:136-137

namespace detail {
inline mb200_qc_opts qcOpts(const AlignmentScores<int>& contigFilterScores, const unsigned minCandidateVariantSize)
{
  mb200_qc_opts o;
  o.filterScores            = contigFilterScores.c();
  o.minCandidateVariantSize = minCandidateVariantSize;
  o.reserved                = 0;
  return o;
}
inline void appendPath(std::vector<uint32_t>& cigar, const path_t& apath, uint64_t& off, uint32_t& len)
{
  off = cigar.size();
  len = static_cast<uint32_t>(apath.size());
  for (const path_segment& ps : apath) cigar.push_back(MB200_SEG_MAKE(ps.length, ps.type));
}
/// the candidate indel runs of a path, as getLargeIndelSegments lists them (GSC/SVCandidateAssemblyRefiner.cpp:173)
inline CandidateSegmentSetType indelRuns(const path_t& apath, const unsigned minSize)
{
  CandidateSegmentSetType runs;
  const unsigned          n(apath.size());
  for (unsigned i(0); i < n;) {
    const bool isIndel((apath[i].type == MB200_SEG_INSERT) || (apath[i].type == MB200_SEG_DELETE));
    if (!isIndel) { ++i; continue; }
    const unsigned start(i);
    bool           isLarge(false);
    for (; i < n && ((apath[i].type == MB200_SEG_INSERT) || (apath[i].type == MB200_SEG_DELETE)); ++i) {
      if (apath[i].length >= minSize) isLarge = true;
    }
    if (isLarge) runs.emplace_back(start, i - 1);
  }
  return runs;
}
}  // namespace detail

/// The 10-mer anchored trim of a complex candidate's reference window (GSC/SVCandidateAssemblyRefiner.cpp:1984-2011):
/// returns adjustedLeadingCut / adjustedTrailingCut for align1RefStr
inline void trimReferenceWindow(const Context& ctx, const std::string& contigSeq, const std::string& align1RefStr,
                                const int leadingCut, const int trailingCut, const int maxLeadingCut, const int maxTrailingCut,
                                int& adjustedLeadingCut, int& adjustedTrailingCut)
{
  std::vector<uint8_t>  pool;
  const mb200_seqview   c(detail::append(pool, contigSeq.begin(), contigSeq.end()));
  const mb200_seqview   r(detail::append(pool, align1RefStr.begin(), align1RefStr.end()));
  const mb200_ref_cuts  cuts{leadingCut, trailingCut, maxLeadingCut, maxTrailingCut};
  mb200_trim_result     out{0, 0};
  ctx.check(mb200_ref_trim_batch(ctx.get(), 1, pool.data(), pool.size(), &c, &r, &cuts, &out));
  adjustedLeadingCut  = out.adjLeadingCut;
  adjustedTrailingCut = out.adjTrailingCut;
}

/// Contig QC and selection of one complex candidate: the spanSet loop and the rank1/rank2 rule of getSmallSVAssembly
/// (GSC/SVCandidateAssemblyRefiner.cpp:2046-2066, 2119-2227).  smallSVAlignments[i].align.beginPos is relative to
/// align1RefStr (after the += adjustedLeadingCut of :2038).  Fills smallSVSegments like assemblyData.smallSVSegments and
/// returns false when no contig qualifies (bestAlignmentIndex is left alone then, like the reference).
inline bool selectSmallSVContig(const Context& ctx, const AlignmentScores<int>& contigFilterScores,
                                const unsigned minCandidateVariantSize, const std::vector<std::string>& contigSeqs,
                                const std::vector<unsigned>& supportReadCounts,
                                const std::vector<AlignmentResult<int>>& smallSVAlignments, const std::string& align1RefStr,
                                std::vector<CandidateSegmentSetType>& smallSVSegments, unsigned& bestAlignmentIndex)
{
  const uint32_t                n(static_cast<uint32_t>(contigSeqs.size()));
  std::vector<uint8_t>          pool;
  std::vector<uint32_t>         cigar;
  std::vector<mb200_seqview>    views(n);
  std::vector<mb200_aln_result> aln(n);
  std::vector<mb200_jump_result> jump(n);
  std::vector<uint32_t>         support(supportReadCounts.begin(), supportReadCounts.end());
  for (uint32_t i(0); i < n; ++i) {
    views[i] = detail::append(pool, contigSeqs[i].begin(), contigSeqs[i].end());
    aln[i]   = mb200_aln_result{smallSVAlignments[i].score, static_cast<int32_t>(smallSVAlignments[i].align.beginPos),
                              smallSVAlignments[i].isJumped ? 1u : 0u, 0, 0};
    detail::appendPath(cigar, smallSVAlignments[i].align.apath, aln[i].cigarOff, aln[i].cigarLen);
  }
  const mb200_seqview ref(detail::append(pool, align1RefStr.begin(), align1RefStr.end()));
  const mb200_qc_opts o(detail::qcOpts(contigFilterScores, minCandidateVariantSize));
  const uint8_t       isJump(0);
  const uint32_t      slotFirst(0);
  std::vector<mb200_contig_qc> qc(n ? n : 1);
  int32_t                      best(-1);
  if (cigar.empty()) cigar.push_back(0);
  ctx.check(mb200_select_contigs_batch(ctx.get(), &o, 1, &isJump, &slotFirst, &n, n, aln.data(), jump.data(), cigar.data(),
                                       cigar.size(), pool.data(), pool.size(), views.data(), &ref, support.data(), qc.data(), &best));
  smallSVSegments.assign(n, CandidateSegmentSetType());
  for (uint32_t i(0); i < n; ++i) {
    if (!qc[i].isCandidate) continue;
    const CandidateSegmentSetType runs(detail::indelRuns(smallSVAlignments[i].align.apath, minCandidateVariantSize));
    smallSVSegments[i].assign(runs.begin() + qc[i].segOrdinal, runs.begin() + qc[i].segOrdinal + qc[i].nSegments);
  }
  if (best < 0) return false;
  bestAlignmentIndex = static_cast<unsigned>(best);
  return true;
}

/// selectJumpContigDNA (GSC/SVCandidateAssemblyRefiner.cpp:1364-1398)
inline bool selectJumpContigDNA(const Context& ctx, const AlignmentScores<int>& contigFilterScores,
                                const std::vector<JumpAlignmentResult<int>>& spanningAlignments, unsigned& bestAlignmentIndex)
{
  const uint32_t                 n(static_cast<uint32_t>(spanningAlignments.size()));
  std::vector<uint32_t>          cigar;
  std::vector<mb200_aln_result>  aln(n ? n : 1);
  std::vector<mb200_jump_result> jump(n ? n : 1);
  std::vector<mb200_seqview>     views(n ? n : 1, mb200_seqview{0, 0, 0});
  std::vector<uint32_t>          support(n ? n : 1, 0);
  for (uint32_t i(0); i < n; ++i) {
    const JumpAlignmentResult<int>& a(spanningAlignments[i]);
    jump[i] = mb200_jump_result{a.score, a.jumpInsertSize, a.jumpRange, static_cast<int32_t>(a.align1.beginPos),
                                static_cast<int32_t>(a.align2.beginPos), 0, 0, 0, 0, 0};
    detail::appendPath(cigar, a.align1.apath, jump[i].cigarOff1, jump[i].cigarLen1);
    detail::appendPath(cigar, a.align2.apath, jump[i].cigarOff2, jump[i].cigarLen2);
  }
  if (cigar.empty()) cigar.push_back(0);
  const std::vector<uint8_t> pool(16, 0);
  const mb200_seqview        ref{0, 1, 0};
  const mb200_qc_opts        o(detail::qcOpts(contigFilterScores, 10));
  const uint8_t              isJump(1);
  const uint32_t             slotFirst(0);
  std::vector<mb200_contig_qc> qc(n ? n : 1);
  int32_t                      best(-1);
  ctx.check(mb200_select_contigs_batch(ctx.get(), &o, 1, &isJump, &slotFirst, &n, n, aln.data(), jump.data(), cigar.data(),
                                       cigar.size(), pool.data(), pool.size(), views.data(), &ref, support.data(), qc.data(), &best));
  if (best < 0) return false;
  bestAlignmentIndex = static_cast<unsigned>(best);
  return true;
}

/// setSmallCandSV for every kept segment of the selected contig (GSC/SVCandidateAssemblyRefiner.cpp:2243-2251, 720-800).
/// Returns one record per segment of smallSVSegments[bestAlignmentIndex]; a record with status MB200_ERR_SV_RANGE is
/// where the reference throws ("Attempting to convert alignment to sv candidate", :768) -- the shim throws there too.
/// Coordinates are offsets into align1RefStr: add bp1ref.get_offset() like the reference does.
inline std::vector<mb200_small_sv> refineSmallSVs(const Context& ctx, const unsigned minCandidateVariantSize,
                                                  const std::string& bestContigSeq, const AlignmentResult<int>& bestAlign,
                                                  const CandidateSegmentSetType& candidateSegments,
                                                  const unsigned bestAlignmentIndex, const std::string& align1RefStr)
{
  std::vector<mb200_small_sv> out;
  if (candidateSegments.empty()) return out;
  const CandidateSegmentSetType runs(detail::indelRuns(bestAlign.align.apath, minCandidateVariantSize));
  unsigned                      ordinal(0);
  while (ordinal < runs.size() && runs[ordinal] != candidateSegments.front()) ordinal++;
  if (ordinal + candidateSegments.size() > runs.size()) throw GeneralException(MB200_ERR_BAD_ARG, "candidate segments do not belong to this alignment");
  std::vector<uint8_t>  pool;
  std::vector<uint32_t> cigar;
  const mb200_seqview   cv(detail::append(pool, bestContigSeq.begin(), bestContigSeq.end()));
  const mb200_seqview   rv(detail::append(pool, align1RefStr.begin(), align1RefStr.end()));
  mb200_aln_result      aln{bestAlign.score, static_cast<int32_t>(bestAlign.align.beginPos), bestAlign.isJumped ? 1u : 0u, 0, 0};
  detail::appendPath(cigar, bestAlign.align.apath, aln.cigarOff, aln.cigarLen);
  mb200_contig_qc qc{1, 0, ordinal, static_cast<uint32_t>(candidateSegments.size()), candidateSegments.front().first,
                     candidateSegments.back().second, 0, 0};
  const uint8_t  isJump(0);
  const uint32_t slotFirst(0);
  const int32_t  best(0);
  uint32_t       svFirst[2] = {0, 0};
  uint64_t       used(0);
  out.resize(candidateSegments.size());
  ctx.check(mb200_small_sv_batch(ctx.get(), minCandidateVariantSize, 1, &isJump, &slotFirst, 1, &best, &qc, &aln, cigar.data(),
                                 cigar.size(), pool.data(), pool.size(), &cv, &rv, svFirst, out.data(), out.size(), &used));
  out.resize(used);
  for (mb200_small_sv& v : out) {
    v.contig = bestAlignmentIndex;
    if (v.status == MB200_ERR_SV_RANGE) throw GeneralException(MB200_ERR_SV_RANGE, "Attempting to convert alignment to sv candidate.");
  }
  return out;
}

// ---- SVAlignmentInfo (row a10): host arithmetic on the refined SV records -------------------------------------------

struct known_pos_range2 {  // lib/blt_util/known_pos_range2.hpp (the two members this path uses)
  known_pos_range2() {}
  known_pos_range2(const int b, const int e) : begin(b), end(e) {}  // (C++11: no aggregate init with member initializers)
  int begin = 0, end = 0;
  int begin_pos() const { return begin; }
  int end_pos() const { return end; }
};

namespace detail {
inline std::string reverseCompCopy(const std::string& s)  // reverseCompCopyStr, lib/blt_util/seq_util.hpp:150-204
{
  std::string r(s.rbegin(), s.rend());
  for (char& c : r) {
    switch (c) {
    case 'A': c = 'T'; break;
    case 'C': c = 'G'; break;
    case 'G': c = 'C'; break;
    case 'T': c = 'A'; break;
    case 'N': break;
    default: throw GeneralException(MB200_ERR_BAD_BASE, "Invalid base in comp_base");
    }
  }
  return r;
}
inline unsigned pathLength(const path_t& p, const bool ofRead, const size_t upTo)
{
  unsigned v(0);
  for (size_t i(0); i < upTo && i < p.size(); ++i) {
    const unsigned t(p[i].type);
    const bool     isRead(t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH);
    const bool     isRef(t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH);
    if (ofRead ? isRead : isRef) v += p[i].length;
  }
  return v;
}
}  // namespace detail

/// The mapping from the refined breakends to the assembled contig and the reference windows that split-read scoring
/// aligns reads against: lib/manta/SVScoreInfo.hpp:35-83, SVScoreInfo.cpp:31-167.  Built from the records the library
/// returns (mb200_small_sv / mb200_jump_sv are in window coordinates, i.e. the reference's ranges minus get_offset()).
struct SVAlignmentInfo {
  /// complex candidate: one kept segment of the selected contig (SVScoreInfo.cpp:109-140)
  SVAlignmentInfo(const mb200_small_sv& sv, const AlignmentResult<int>& alignment, const std::string& contig, const std::string& bp1Ref)
    : _isSpanning(false), _bp1ContigReversed(false), _bp2ContigReversed(false)
  {
    const path_t&  apath(alignment.align.apath);
    const unsigned alignRefEnd(alignment.align.beginPos + detail::pathLength(apath, false, apath.size()));
    // getExtendedContig, lib/alignment/AlignmentUtil.cpp:112-127
    contigSeq = bp1Ref.substr(0, alignment.align.beginPos) + contig + bp1Ref.substr(alignRefEnd);
    bp1RefSeq = bp1Ref;
    const int bp1Hom(sv.bp1End - sv.bp1Begin - 1), bp2Hom(sv.bp2End - sv.bp2Begin - 1);
    bp1ContigOffset.begin = static_cast<int>(alignment.align.beginPos + detail::pathLength(apath, true, sv.pathFirst)) - 1;
    bp1ContigOffset.end   = bp1ContigOffset.begin + bp1Hom;
    bp2ContigOffset.begin = static_cast<int>(alignment.align.beginPos + detail::pathLength(apath, true, sv.pathLast + 1)) - 1;
    bp2ContigOffset.end   = bp2ContigOffset.begin + bp2Hom;
    bp1RefOffset = known_pos_range2{sv.bp1Begin, sv.bp1End};
    bp2RefOffset = known_pos_range2{sv.bp2Begin, sv.bp2End};
  }

  /// spanning candidate (SVScoreInfo.cpp:47-108); alignRef1 / alignRef2 are the windows align1 / align2 ran against
  SVAlignmentInfo(const mb200_jump_sv& sv, const JumpAlignmentResult<int>& alignment, const mb200_jump_orient& orient,
                  const std::string& contig, const std::string& alignRef1, const std::string& alignRef2)
    : _isSpanning(true), _bp1ContigReversed(orient.isBp1Reversed != 0), _bp2ContigReversed(orient.isBp2Reversed != 0)
  {
    const unsigned align2RefEnd(alignment.align2.beginPos + detail::pathLength(alignment.align2.apath, false, alignment.align2.apath.size()));
    // getExtendedContig, lib/alignment/AlignmentUtil.cpp:97-110
    contigSeq = alignRef1.substr(0, alignment.align1.beginPos) + contig + alignRef2.substr(align2RefEnd);
    const int bp1Hom(sv.bp1End - sv.bp1Begin - 1), bp2Hom(sv.bp2End - sv.bp2Begin - 1);
    const int align1Size(detail::pathLength(alignment.align1.apath, true, alignment.align1.apath.size()));
    bp1ContigOffset.begin = static_cast<int>(alignment.align1.beginPos) + align1Size - 1;
    bp2ContigOffset.begin = bp1ContigOffset.begin + static_cast<int>(alignment.jumpInsertSize);
    if (orient.isBp2AlignedFirst) std::swap(bp1ContigOffset, bp2ContigOffset);
    bp1ContigOffset.end = bp1ContigOffset.begin + bp1Hom;
    bp2ContigOffset.end = bp2ContigOffset.begin + bp2Hom;
    if (_bp1ContigReversed || _bp2ContigReversed) {
      revContigSeq = detail::reverseCompCopy(contigSeq);
      const int         revSize(static_cast<int>(contigSeq.size()) - 2);  // -2: the offset jumps to the other side of the breakend
      known_pos_range2& r(_bp1ContigReversed ? bp1ContigOffset : bp2ContigOffset);
      const known_pos_range2 tmp(r);
      r.begin = revSize - tmp.end;
      r.end   = revSize - tmp.begin;
    }
    bp1RefSeq = orient.isBp2AlignedFirst ? alignRef2 : alignRef1;
    bp2RefSeq = orient.isBp2AlignedFirst ? alignRef1 : alignRef2;
    bp1RefOffset = known_pos_range2{sv.bp1Begin, sv.bp1Begin + bp1Hom};
    bp2RefOffset = known_pos_range2{sv.bp2Begin, sv.bp2Begin + bp2Hom};
  }

  const std::string& bp1ContigSeq() const { return (_bp1ContigReversed ? revContigSeq : contigSeq); }
  const std::string& bp2ContigSeq() const { return (_bp2ContigReversed ? revContigSeq : contigSeq); }
  const std::string& bp1ReferenceSeq() const { return bp1RefSeq; }
  const std::string& bp2ReferenceSeq() const { return (_isSpanning ? bp2RefSeq : bp1RefSeq); }
  bool               isSpanning() const { return _isSpanning; }

  bool isMinBpEdge(const unsigned minEdge) const  // SVScoreInfo.cpp:142-164
  {
    const int iminEdge(minEdge);
    if ((bp1ContigOffset.begin + 1) < iminEdge) return false;
    if ((bp2ContigOffset.begin + 1) < iminEdge) return false;
    if ((bp1RefOffset.begin + 1) < iminEdge) return false;
    if ((bp2RefOffset.begin + 1) < iminEdge) return false;
    const int contigBpSize(static_cast<int>(contigSeq.size()) - 1);
    if ((contigBpSize - bp1ContigOffset.end) < iminEdge) return false;
    if ((contigBpSize - bp2ContigOffset.end) < iminEdge) return false;
    if ((static_cast<int>(bp1ReferenceSeq().size()) - 1 - bp1RefOffset.end) < iminEdge) return false;
    if ((static_cast<int>(bp2ReferenceSeq().size()) - 1 - bp2RefOffset.end) < iminEdge) return false;
    return true;
  }

  known_pos_range2 bp1ContigOffset, bp2ContigOffset, bp1RefOffset, bp2RefOffset;

private:
  std::string contigSeq, revContigSeq, bp1RefSeq, bp2RefSeq;
  const bool  _isSpanning, _bp1ContigReversed, _bp2ContigReversed;
};

}  // namespace mb200



