// ref_refiner_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Second adapter TU of oracle/_ref/libmanta_ref.so.  The contig QC functions of the refiner are file-static in
// GSC/SVCandidateAssemblyRefiner.cpp, so this TU #includes that UNMODIFIED source file from where it lies under
// /root/reference (nothing is copied) and calls the statics directly:
//   findCandidateVariantsFromComplexSVContigAlignment  GSC/SVCandidateAssemblyRefiner.cpp:430
//   getLargeIndelSegments / getLargestIndelSize        :173 / :210
//   isJumpAlignmentQCFail / isLowQualityJumpAlignment  :1266 / :1287
//   selectJumpContigDNA                                :1364
// Two pieces of the reference are inline in the member function getSmallSVAssembly (which needs BAM and FASTA files
// and therefore cannot be driven here): the 10-mer window trim (:1984-2011) and the rank1/rank2 contig selection
// (:2119-2227).  They are restated below around the reference's own types and marked RESTATED; everything they
// consume (QC flags, segments, variant sizes) comes from the reference statics.
// Unused members of the refiner class are dropped at link time (--gc-sections).

#include "applications/GenerateSVCandidates/SVCandidateAssemblyRefiner.cpp"

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <unordered_set>

#include "alignment/GlobalJumpAligner.hpp"
#include "alignment/GlobalLargeIndelAligner.hpp"
#include "assembly/IterativeAssembler.hpp"

#include "manta/SVScoreInfo.hpp"

#include "manta_b200.h"
#include "manta_oracle.h"  // checksum terms only (macros); no oracle code is linked here

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

AlignmentScores<int> toScores2(const mb200_scores* s)
{
  return AlignmentScores<int>(s->match, s->mismatch, s->open, s->extend, s->offEdge, s->isAllowEdgeInsertion != 0);
}

ALIGNPATH::path_t toPath(const uint32_t* seg, const uint32_t n)
{
  ALIGNPATH::path_t p;
  for (uint32_t i(0); i < n; ++i) {
    p.emplace_back(static_cast<ALIGNPATH::align_t>(MB200_SEG_TYPE(seg[i])), MB200_SEG_LEN(seg[i]));
  }
  return p;
}

typedef std::vector<std::pair<unsigned, unsigned>> SegSet;

/// QC of one complex-candidate contig: the spanSet loop of getSmallSVAssembly (:2046-2066) over the reference's
/// findCandidateVariantsFromComplexSVContigAlignment
void qcComplexContig(const mb200_qc_opts& o, const Alignment& align, const std::string& contigSeq,
                     const std::string& refSeq, mb200_contig_qc& q)
{
  const AlignmentScores<int> scores(toScores2(&o.filterScores));
  SegSet                     kept;
  bool                       isSmallSVCandidate(false);
  for (const unsigned maxQCRefSpan : {100u, 200u}) {
    SegSet     segments;
    const bool isCandidate(findCandidateVariantsFromComplexSVContigAlignment(
        maxQCRefSpan, scores, align, contigSeq, refSeq, o.minCandidateVariantSize, segments));
    if (!isCandidate) continue;
    if (segments.size() > kept.size()) kept = segments;
    isSmallSVCandidate = true;
  }
  std::memset(&q, 0, sizeof(q));
  if (!isSmallSVCandidate) return;
  SegSet all;
  getLargeIndelSegments(align.apath, o.minCandidateVariantSize, all);
  unsigned ordinal(0);
  while (ordinal < all.size() && all[ordinal] != kept.front()) ordinal++;
  q.isCandidate = 1;
  q.segOrdinal  = ordinal;
  q.nSegments   = static_cast<uint32_t>(kept.size());
  q.pathFirst   = kept.front().first;
  q.pathLast    = kept.back().second;
  q.variantSize = getLargestIndelSize(align.apath, kept);
}

/// RESTATED from the body of getSmallSVAssembly (:2119-2227): keep the two best QC'd contigs while walking the
/// contigs in order, then let the runner-up win on score ratio plus read support or variant size
int32_t selectComplex(const uint32_t nContigs, const mb200_aln_result* aln, const mb200_contig_qc* qc,
                      const uint32_t* support)
{
  ContigScoringInfo rank1, rank2;
  for (uint32_t c(0); c < nContigs; ++c) {
    if (!qc[c].isCandidate) continue;
    ContigScoringInfo cur;
    cur.isDefined   = true;
    cur.index       = c;
    cur.score       = aln[c].score;
    cur.variantSize = qc[c].variantSize;
    cur.isJumped    = (aln[c].isJumped != 0);
    const bool sameJumpState(cur.isJumped == rank1.isJumped);
    const bool takesRank1(
        (!rank1.isDefined) || (cur.isJumped && !rank1.isJumped) || (sameJumpState && (cur.score > rank1.score)));
    if (takesRank1) {
      if (rank1.isDefined) rank2 = rank1;
      rank1 = cur;
    } else if ((!rank2.isDefined) || (cur.score > rank2.score)) {
      rank2 = cur;
    }
  }
  if (rank2.isDefined && !(rank1.isJumped && !rank2.isJumped)) {
    const unsigned n1(support[rank1.index]);
    const unsigned n2(support[rank2.index]);
    static const float minScoreRatio(0.9f);
    static const float minSupportReadCountRatio(1.2f);
    static const float minVariantSizeRatio(1.1f);
    const bool rank2IsBest(
        (rank2.score > (rank1.score * minScoreRatio)) &&
        ((n2 > (n1 * minSupportReadCountRatio)) || (rank2.variantSize > (rank1.variantSize * minVariantSizeRatio))));
    if (rank2IsBest) rank1 = rank2;
  }
  return rank1.isDefined ? static_cast<int32_t>(rank1.index) : -1;
}

/// spanning candidate: QC flags per contig from the statics, selection by the reference's selectJumpContigDNA
int32_t selectSpanning(const mb200_qc_opts& o, const uint32_t nContigs, const mb200_jump_result* jump,
                       const uint32_t* cigarPool, mb200_contig_qc* qc)
{
  const AlignmentScores<int> scores(toScores2(&o.filterScores));
  SVCandidateAssemblyData    ad;
  ad.contigs.resize(nContigs);
  ad.spanningAlignments.resize(nContigs);
  for (uint32_t c(0); c < nContigs; ++c) {
    JumpAlignmentResult<int>& a(ad.spanningAlignments[c]);
    a.score           = jump[c].score;
    a.jumpInsertSize  = jump[c].jumpInsertSize;
    a.jumpRange       = jump[c].jumpRange;
    a.align1.beginPos = jump[c].beginPos1;
    a.align2.beginPos = jump[c].beginPos2;
    a.align1.apath    = toPath(cigarPool + jump[c].cigarOff1, jump[c].cigarLen1);
    a.align2.apath    = toPath(cigarPool + jump[c].cigarOff2, jump[c].cigarLen2);
    std::memset(&qc[c], 0, sizeof(qc[c]));
    qc[c].isCandidate = isJumpAlignmentQCFail(a) ? 0 : 1;
    // the reference only evaluates this for the selected contig; it reads nothing but the two paths
    bool low(true);
    if (a.align1.isAligned() || a.align2.isAligned()) low = isLowQualityJumpAlignment(a, scores, false);
    qc[c].isLowQuality = low ? 1 : 0;
  }
  const bool isSelected(selectJumpContigDNA(ad, scores));
  return isSelected ? static_cast<int32_t>(ad.bestAlignmentIndex) : -1;
}

/// RESTATED from getSmallSVAssembly (:1984-2011): walk inwards from both cut limits to the first reference 10-mer
/// that also occurs in the contig.  Same container as the reference (a hash set of std::string 10-mers) so that the
/// CPU baseline pays the reference's cost for this step.
void trimWindow(const std::string& contig, const std::string& ref, const mb200_ref_cuts& cut, mb200_trim_result& out)
{
  static const int                K(10);
  std::unordered_set<std::string> mers;
  for (unsigned p(0); p + K <= contig.size(); ++p) mers.insert(contig.substr(p, K));
  const pos_t refSize(ref.size());
  const pos_t lowest(cut.leadingCut);
  const pos_t highest(refSize - (cut.trailingCut + K));
  const pos_t fwdStop(std::min(static_cast<pos_t>(cut.maxLeadingCut), highest));
  pos_t       at(lowest);
  while (at <= fwdStop && mers.count(ref.substr(at, K)) == 0) at++;
  out.adjLeadingCut = at;
  const pos_t revStop(std::max(lowest, refSize - cut.maxTrailingCut));
  at = highest;
  while (at >= revStop && mers.count(ref.substr(at, K)) == 0) at--;
  out.adjTrailingCut = refSize - (at + K);
}

}  // namespace

REF_API int ref_ref_trim_batch(uint32_t n, const uint8_t* pool, const mb200_seqview* contig, const mb200_seqview* ref,
                               const mb200_ref_cuts* cuts, mb200_trim_result* out)
{
  for (uint32_t i(0); i < n; ++i) {
    if (contig[i].len < 10) return MB200_ERR_BAD_ARG;
    const std::string c(reinterpret_cast<const char*>(pool + contig[i].off), contig[i].len);
    const std::string r(reinterpret_cast<const char*>(pool + ref[i].off), ref[i].len);
    trimWindow(c, r, cuts[i], out[i]);
  }
  return MB200_OK;
}

REF_API int ref_select_contigs_batch(const mb200_qc_opts* opts, uint32_t nLoci, const uint8_t* isJump,
                                     const uint32_t* slotFirst, const uint32_t* nContigs, uint32_t /*nSlots*/,
                                     const mb200_aln_result* aln, const mb200_jump_result* jump,
                                     const uint32_t* cigarPool, const uint8_t* pool, const mb200_seqview* contigSeq,
                                     const mb200_seqview* ref1, const uint32_t* supportCount, mb200_contig_qc* qc,
                                     int32_t* best)
{
  for (uint32_t li(0); li < nLoci; ++li) {
    const uint32_t s0(slotFirst[li]);
    if (isJump[li]) {
      best[li] = selectSpanning(*opts, nContigs[li], jump + s0, cigarPool, qc + s0);
      continue;
    }
    const std::string refSeq(reinterpret_cast<const char*>(pool + ref1[li].off), ref1[li].len);
    for (uint32_t c(0); c < nContigs[li]; ++c) {
      const uint32_t s(s0 + c);
      Alignment      align;
      align.beginPos = aln[s].beginPos;
      align.apath    = toPath(cigarPool + aln[s].cigarOff, aln[s].cigarLen);
      const std::string contig(reinterpret_cast<const char*>(pool + contigSeq[s].off), contigSeq[s].len);
      qcComplexContig(*opts, align, contig, refSeq, qc[s]);
    }
    best[li] = selectComplex(nContigs[li], aln + s0, qc + s0, supportCount + s0);
  }
  return MB200_OK;
}

/// Refined small SVs through the reference's own setSmallCandSV (:720) for every kept segment of the selected contig
/// (the loop of getSmallSVAssembly :2243-2251); readRange / refRange / cipos are recomputed here from the SVCandidate
/// the static returns (bp1 = [refBegin - 1, +cipos.end + 1), bp2 = [refEnd, ...), insertSeq = contig[readRange]).
REF_API int ref_small_sv_batch(uint32_t minCandidateVariantSize, uint32_t nLoci, const uint8_t* isJump,
                               const uint32_t* slotFirst, uint32_t /*nSlots*/, const int32_t* best, const mb200_contig_qc* qc,
                               const mb200_aln_result* aln, const uint32_t* cigarPool, const uint8_t* pool,
                               const mb200_seqview* contigSeq, const mb200_seqview* ref1, uint32_t* svFirst,
                               mb200_small_sv* svs, uint64_t svCap, uint64_t* svUsed)
{
  uint64_t   used(0);
  GSCOptions opt;
  for (uint32_t li(0); li < nLoci; ++li) {
    svFirst[li] = static_cast<uint32_t>(used);
    if (isJump[li] || best[li] < 0) continue;
    const uint32_t s(slotFirst[li] + static_cast<uint32_t>(best[li]));
    Alignment      align;
    align.beginPos = aln[s].beginPos;
    align.apath    = toPath(cigarPool + aln[s].cigarOff, aln[s].cigarLen);
    const std::string        contig(reinterpret_cast<const char*>(pool + contigSeq[s].off), contigSeq[s].len);
    reference_contig_segment ref;
    ref.seq() = std::string(reinterpret_cast<const char*>(pool + ref1[li].off), ref1[li].len);
    ref.set_offset(0);
    SegSet all;
    getLargeIndelSegments(align.apath, minCandidateVariantSize, all);
    for (uint32_t k(0); k < qc[s].nSegments; ++k, ++used) {
      if (used >= svCap) continue;
      const std::pair<unsigned, unsigned> segRange(all[qc[s].segOrdinal + k]);
      mb200_small_sv                      v;
      std::memset(&v, 0, sizeof(v));
      v.locus = li; v.contig = static_cast<uint32_t>(best[li]); v.segment = k;
      v.pathFirst = segRange.first; v.pathLast = segRange.second;
      // the two ranges, like the head of setSmallCandSV (they are not returned by it)
      {
        pos_t readPos(0), refPos(align.beginPos);
        for (unsigned i(0); i < align.apath.size(); ++i) {
          const ALIGNPATH::path_segment& ps(align.apath[i]);
          if (i == segRange.first) { v.refBegin = refPos; v.readBegin = readPos; }
          if (is_segment_type_ref_length(ps.type)) refPos += ps.length;
          if (is_segment_type_read_length(ps.type)) readPos += ps.length;
          if (i == segRange.second) { v.refEnd = refPos; v.readEnd = readPos; }
        }
      }
      const known_pos_range2 cipos(getVariantRange(
          ref.seq(), known_pos_range2(v.refBegin, v.refEnd), contig, known_pos_range2(v.readBegin, v.readEnd)));
      v.ciposBegin = cipos.begin_pos();
      v.ciposEnd   = cipos.end_pos();
      SVCandidate sv;
      try {
        setSmallCandSV(ref, contig, align, segRange, sv, opt);
        v.bp1Begin = sv.bp1.interval.range.begin_pos();
        v.bp1End   = sv.bp1.interval.range.end_pos();
        v.bp2Begin = sv.bp2.interval.range.begin_pos();
        v.bp2End   = sv.bp2.interval.range.end_pos();
        if (sv.insertSeq != contig.substr(v.readBegin, v.readEnd - v.readBegin)) v.status = -1;  // would be a shim bug
        if (sv.insertAlignment.size() != (segRange.second - segRange.first + 1)) v.status = -2;
      } catch (const illumina::common::GeneralException&) {
        v.status = MB200_ERR_SV_RANGE;
      }
      svs[used] = v;
    }
  }
  svFirst[nLoci] = static_cast<uint32_t>(used);
  *svUsed        = used;
  return used > svCap ? MB200_ERR_POOL_OVERFLOW : MB200_OK;
}

/// Large insertions with an incompletely assembled insert, through the reference's own statics: the per-contig
/// candidate test of getSmallSVAssembly (:2079-2112, restated here around isLargeInsertAlignment :612 because the loop is
/// inline in a member function that needs BAM/FASTA files) and processLargeInsertion (:833-1007) itself on an
/// SVCandidateAssemblyData filled from the flat arrays.  The set of excluded positions is the rule of :2253-2256
/// (getExtendedSVType == INSERT) applied to the refined small SVs.  Where processLargeInsertion returns early the exit is
/// recovered from what it leaves behind; the two QC exits at :955 and :966-973 leave the same state, so they are
/// reported as MB200_LI_FLANK_QC both.
REF_API int ref_large_insertion_batch(const mb200_scores* edgeScores, const mb200_scores* completeScores, uint32_t nLoci,
                                      const uint8_t* isJump, const uint32_t* slotFirst, const uint32_t* nContigs,
                                      uint32_t nSlots, const mb200_aln_result* aln, const uint32_t* cigarPool,
                                      const uint8_t* pool, const mb200_seqview* contigSeq, const int32_t* consRange,
                                      const mb200_seqview* ref1, const mb200_ref_cuts* cuts, const uint32_t* svFirst,
                                      const mb200_small_sv* svs, mb200_large_insert_info* info, mb200_large_insertion* out,
                                      uint32_t* outCigar, uint64_t outCigarCap, uint64_t* outCigarUsed)
{
  const GlobalAligner<int> edgeAligner(toScores2(edgeScores));
  const GlobalAligner<int> completeAligner(toScores2(completeScores));
  GSCOptions               opt;
  uint64_t                 used(0);
  bool                     overflow(false);
  std::memset(info, 0, sizeof(*info) * nSlots);
  std::memset(out, 0, sizeof(*out) * nLoci);
  for (uint32_t li(0); li < nLoci; ++li) {
    if (isJump[li]) continue;
    const uint32_t          s0(slotFirst[li]), nc(nContigs[li]);
    SVCandidateAssemblyData ad;
    ad.isSpanning = false;
    ad.bp1ref.seq() = std::string(reinterpret_cast<const char*>(pool + ref1[li].off), ref1[li].len);
    ad.bp1ref.set_offset(0);
    ad.contigs.resize(nc);
    ad.smallSVAlignments.resize(nc);
    ad.smallSVSegments.resize(nc);
    ad.largeInsertInfo.resize(nc);
    ad.extendedContigs.resize(nc);
    std::vector<unsigned> candidateIndex;
    for (uint32_t k(0); k < nc; ++k) {
      const uint32_t s(s0 + k);
      ad.contigs[k].seq = std::string(reinterpret_cast<const char*>(pool + contigSeq[s].off), contigSeq[s].len);
      ad.contigs[k].conservativeRange.set_range(consRange[2 * s], consRange[2 * s + 1]);
      ad.smallSVAlignments[k].score          = aln[s].score;
      ad.smallSVAlignments[k].align.beginPos = aln[s].beginPos;
      ad.smallSVAlignments[k].align.apath    = toPath(cigarPool + aln[s].cigarOff, aln[s].cigarLen);
      // ---- :2079-2112
      LargeInsertionInfo& candidateInsertInfo(ad.largeInsertInfo[k]);
      candidateInsertInfo.clear();
      LargeInsertionInfo insertInfo;
      ALIGNPATH::path_t  apath_conservative(ad.smallSVAlignments[k].align.apath);
      apath_limit_read_length(ad.contigs[k].conservativeRange, apath_conservative);
      bool isCandidate(isLargeInsertAlignment(edgeAligner, apath_conservative, insertInfo));
      if (isCandidate) {
        LargeInsertionInfo insertInfo2;
        isCandidate = isLargeInsertAlignment(edgeAligner, ad.smallSVAlignments[k].align.apath, insertInfo2);
        if ((insertInfo.isLeftCandidate != insertInfo2.isLeftCandidate) ||
            (insertInfo.isRightCandidate != insertInfo2.isRightCandidate)) {
          isCandidate = false;
        }
        insertInfo.contigOffset = insertInfo2.contigOffset;
        insertInfo.refOffset    = insertInfo2.refOffset;
      }
      if (isCandidate) {
        candidateInsertInfo = insertInfo;
        candidateIndex.push_back(k);
        info[s].isLeftCandidate  = insertInfo.isLeftCandidate ? 1 : 0;
        info[s].isRightCandidate = insertInfo.isRightCandidate ? 1 : 0;
        info[s].contigOffset     = insertInfo.contigOffset;
        info[s].refOffset        = insertInfo.refOffset;
        info[s].score            = insertInfo.score;
      }
    }
    std::set<pos_t> excludedPos;
    if (svFirst && svs) {
      for (uint32_t v(svFirst[li]); v < svFirst[li + 1]; ++v) {
        if (svs[v].status != 0) continue;
        const unsigned deleteSize(svs[v].bp2Begin - svs[v].bp1Begin);
        const unsigned insertSize(svs[v].readEnd - svs[v].readBegin);
        if (deleteSize < insertSize) excludedPos.insert(svs[v].bp1Begin);
      }
    }
    SVCandidate sv;
    const pos_t leadingCut(cuts ? cuts[li].leadingCut : 0), trailingCut(cuts ? cuts[li].trailingCut : 0);
    mb200_large_insertion& o(out[li]);
    bool                   threw(false);
    try {
      processLargeInsertion(sv, leadingCut, trailingCut, completeAligner, candidateIndex, excludedPos, ad, opt);
    } catch (const illumina::common::GeneralException&) {
      threw = true;
    }
    if (ad.contigs.size() == nc) continue;  // no pair: MB200_LI_NONE
    // which pair was joined: the joined contig is left + 100 N + right
    const std::string& fake(ad.contigs[nc].seq);
    bool               found(false);
    for (uint32_t a(0); a < nc && !found; ++a) {
      for (uint32_t b(0); b < nc && !found; ++b) {
        if (a == b || !(ad.largeInsertInfo[a].isLeftCandidate && ad.largeInsertInfo[b].isRightCandidate)) continue;
        if (fake == ad.contigs[a].seq + std::string(100, 'N') + ad.contigs[b].seq) {
          // several pairs can spell the same joined contig only if two contigs are equal, which the assembler never returns
          o.leftContig = a; o.rightContig = b;
          found = true;
        }
      }
    }
    if (!found) return MB200_ERR_BAD_ARG;
    {
      const LargeInsertionInfo& i1(ad.largeInsertInfo[o.leftContig]);
      const LargeInsertionInfo& i2(ad.largeInsertInfo[o.rightContig]);
      o.breakDist  = std::abs(static_cast<long>(aln[s0 + o.leftContig].beginPos + i1.refOffset) -
                              static_cast<long>(aln[s0 + o.rightContig].beginPos + i2.refOffset));
      o.breakScore = i1.score + i2.score;
    }
    const SVCandidateAssemblyData::SmallAlignmentResultType& fa(ad.smallSVAlignments[nc]);
    o.score = fa.score; o.beginPos = fa.align.beginPos; o.cigarLen = fa.align.apath.size(); o.cigarOff = used;
    for (const ALIGNPATH::path_segment& ps : fa.align.apath) {
      if (used < outCigarCap) outCigar[used] = MB200_SEG_MAKE(ps.length, static_cast<uint32_t>(ps.type));
      else overflow = true;
      ++used;
    }
    o.status = MB200_LI_SEGMENT_QC;
    const std::vector<std::pair<unsigned, unsigned>>& segs(ad.smallSVSegments[nc]);
    if ((1 != segs.size()) || (segs[0].second < segs[0].first)) continue;
    o.segFirst = segs[0].first; o.segLast = segs[0].second;
    o.status   = MB200_LI_FLANK_QC;
    if (threw) {
      o.status = MB200_LI_RANGE_ERROR;
    } else if (ad.extendedContigs[nc].empty()) {
      continue;  // :955 or :966-973
    } else if (ad.svs.empty()) {
      o.status = MB200_LI_DUPLICATE;
    } else {
      o.status = MB200_LI_ACCEPTED;
    }
    const known_pos_range2 trim(getInsertTrim(fa.align.apath, segs[0]));
    o.insertTrimBegin = trim.begin_pos(); o.insertTrimEnd = trim.end_pos();
    {
      pos_t readPos(0), refPos(fa.align.beginPos);
      for (unsigned i(0); i < fa.align.apath.size(); ++i) {
        const ALIGNPATH::path_segment& ps(fa.align.apath[i]);
        if (i == segs[0].first) { o.refBegin = refPos; o.readBegin = readPos; }
        if (is_segment_type_ref_length(ps.type)) refPos += ps.length;
        if (is_segment_type_read_length(ps.type)) readPos += ps.length;
        if (i == segs[0].second) { o.refEnd = refPos; o.readEnd = readPos; }
      }
    }
    const known_pos_range2 cipos(getVariantRange(ad.bp1ref.seq(), known_pos_range2(o.refBegin, o.refEnd), fake,
                                                 known_pos_range2(o.readBegin, o.readEnd)));
    o.ciposBegin = cipos.begin_pos(); o.ciposEnd = cipos.end_pos();
    if (o.status == MB200_LI_RANGE_ERROR) continue;
    if (o.status == MB200_LI_ACCEPTED) {
      const SVCandidate& nsv(ad.svs.back());
      o.bp1Begin = nsv.bp1.interval.range.begin_pos(); o.bp1End = nsv.bp1.interval.range.end_pos();
      o.bp2Begin = nsv.bp2.interval.range.begin_pos(); o.bp2End = nsv.bp2.interval.range.end_pos();
      o.leftSeqBegin = ad.contigs[o.leftContig].seq.size() - nsv.unknownSizeInsertionLeftSeq.size();
      o.rightSeqLen  = nsv.unknownSizeInsertionRightSeq.size();
      if (!nsv.isUnknownSizeInsertion) return MB200_ERR_BAD_ARG;
    } else {  // duplicate: the breakends the static computed before it gave up (same formula as setSmallCandSV)
      o.bp1Begin = o.refBegin - 1; o.bp1End = o.bp1Begin + o.ciposEnd + 1;
      o.bp2Begin = o.refEnd; o.bp2End = o.bp2Begin + o.ciposEnd + 1;
    }
  }
  *outCigarUsed = used;
  return overflow ? MB200_ERR_POOL_OVERFLOW : MB200_OK;
}

/// Refined spanning breakends through the reference's own generateRefinedSVCandidateFromJumpAlignment (:1190), which
/// applies adjustAssembledBreakend (:59) to both breakends; only the window sizes of bp1ref / bp2ref matter to it.
REF_API int ref_jump_sv_batch(uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst, uint32_t /*nSlots*/,
                              const int32_t* best, const mb200_jump_result* jump, const uint32_t* cigarPool,
                              const mb200_seqview* ref1, const mb200_seqview* ref2, const mb200_jump_orient* orient,
                              mb200_jump_sv* out)
{
  for (uint32_t li(0); li < nLoci; ++li) {
    mb200_jump_sv v;
    std::memset(&v, 0, sizeof(v));
    if (isJump[li] && best[li] >= 0) {
      const mb200_jump_result& r(jump[slotFirst[li] + static_cast<uint32_t>(best[li])]);
      const mb200_jump_orient& o(orient[li]);
      SVCandidateAssemblyData  ad;
      ad.bestAlignmentIndex = static_cast<unsigned>(best[li]);
      ad.spanningAlignments.resize(ad.bestAlignmentIndex + 1);
      JumpAlignmentResult<int>& a(ad.spanningAlignments[ad.bestAlignmentIndex]);
      a.score           = r.score;
      a.jumpInsertSize  = r.jumpInsertSize;
      a.jumpRange       = r.jumpRange;
      a.align1.beginPos = r.beginPos1;
      a.align2.beginPos = r.beginPos2;
      a.align1.apath    = toPath(cigarPool + r.cigarOff1, r.cigarLen1);
      a.align2.apath    = toPath(cigarPool + r.cigarOff2, r.cigarLen2);
      ad.bporient.isBp2AlignedFirst = (o.isBp2AlignedFirst != 0);
      ad.bporient.isBp1Reversed     = (o.isBp1Reversed != 0);
      ad.bporient.isBp2Reversed     = (o.isBp2Reversed != 0);
      // align1 was aligned to bp2's window when bp2 is aligned first (:1544-1547)
      ad.bp1ref.seq() = std::string(o.isBp2AlignedFirst ? ref2[li].len : ref1[li].len, 'N');
      ad.bp2ref.seq() = std::string(o.isBp2AlignedFirst ? ref1[li].len : ref2[li].len, 'N');
      ad.bp1ref.set_offset(0);
      ad.bp2ref.set_offset(0);
      SVCandidate sv;
      sv.bp1.state = o.bp1RightOpen ? SVBreakendState::RIGHT_OPEN : SVBreakendState::LEFT_OPEN;
      sv.bp2.state = o.bp2RightOpen ? SVBreakendState::RIGHT_OPEN : SVBreakendState::LEFT_OPEN;
      generateRefinedSVCandidateFromJumpAlignment(ad, sv);
      v.valid       = 1;
      v.contig      = sv.assemblyAlignIndex;
      v.bp1Begin    = sv.bp1.interval.range.begin_pos();
      v.bp1End      = sv.bp1.interval.range.end_pos();
      v.bp2Begin    = sv.bp2.interval.range.begin_pos();
      v.bp2End      = sv.bp2.interval.range.end_pos();
      v.insertBegin = static_cast<int32_t>(apath_read_length(a.align1.apath));  // getFwdStrandQuerySegments, AlignmentUtil.cpp
      v.insertEnd   = v.insertBegin + static_cast<int32_t>(a.jumpInsertSize);
    }
    out[li] = v;
  }
  return MB200_OK;
}

/// SVAlignmentInfo (lib/manta/SVScoreInfo.cpp:31-140) of one refined SV, flattened: out[0..7] = bp1ContigOffset, bp2ContigOffset,
/// bp1RefOffset, bp2RefOffset (begin, end each); out[8] = isMinBpEdge(minEdge); the extended contig (getExtendedContig,
/// lib/alignment/AlignmentUtil.cpp:97-127) and its reverse complement go to contigOut / revOut (NUL-terminated, cap bytes).
/// Pins the host-side shim class mb200::SVAlignmentInfo (tests/cpp/align_info_test.cpp).
REF_API int ref_sv_align_info_small(const char* contig, const char* ref1, int32_t beginPos, const uint32_t* path, uint32_t nPath,
                                    uint32_t segFirst, uint32_t segLast, int32_t bp1Begin, int32_t bp1End, int32_t bp2Begin,
                                    int32_t bp2End, uint32_t minEdge, int32_t* out, char* contigOut, char* revOut, uint32_t cap)
{
  SVCandidate sv;
  sv.setPrecise();
  sv.bp1.interval.range.set_range(bp1Begin, bp1End);
  sv.bp2.interval.range.set_range(bp2Begin, bp2End);
  sv.assemblyAlignIndex   = 0;
  sv.assemblySegmentIndex = 0;
  SVCandidateAssemblyData ad;
  ad.isSpanning = false;
  ad.smallSVAlignments.resize(1);
  ad.smallSVAlignments[0].align.beginPos = beginPos;
  ad.smallSVAlignments[0].align.apath    = toPath(path, nPath);
  ad.smallSVSegments.resize(1);
  ad.smallSVSegments[0].push_back(std::make_pair(segFirst, segLast));
  ad.extendedContigs.resize(1);
  ad.bp1ref.seq() = ref1;
  ad.bp1ref.set_offset(0);
  getExtendedContig(ad.smallSVAlignments[0], std::string(contig), ad.bp1ref.seq(), ad.extendedContigs[0]);
  const SVAlignmentInfo ai(sv, ad);
  out[0] = ai.bp1ContigOffset.begin_pos(); out[1] = ai.bp1ContigOffset.end_pos();
  out[2] = ai.bp2ContigOffset.begin_pos(); out[3] = ai.bp2ContigOffset.end_pos();
  out[4] = ai.bp1RefOffset.begin_pos(); out[5] = ai.bp1RefOffset.end_pos();
  out[6] = ai.bp2RefOffset.begin_pos(); out[7] = ai.bp2RefOffset.end_pos();
  out[8] = ai.isMinBpEdge(minEdge) ? 1 : 0;
  const std::string& c(ai.bp1ContigSeq());
  if (c.size() + 1 > cap) return MB200_ERR_POOL_OVERFLOW;
  std::memcpy(contigOut, c.c_str(), c.size() + 1);
  revOut[0] = 0;
  return MB200_OK;
}

REF_API int ref_sv_align_info_jump(const char* contig, const char* alignRef1, const char* alignRef2, int32_t score, uint32_t jumpInsertSize,
                                   uint32_t jumpRange, int32_t beginPos1, const uint32_t* path1, uint32_t nPath1, int32_t beginPos2,
                                   const uint32_t* path2, uint32_t nPath2, const mb200_jump_orient* orient, int32_t bp1Begin,
                                   int32_t bp1End, int32_t bp2Begin, int32_t bp2End, uint32_t minEdge, int32_t* out, char* contigOut,
                                   char* revOut, uint32_t cap)
{
  SVCandidate sv;
  sv.setPrecise();
  sv.bp1.interval.range.set_range(bp1Begin, bp1End);
  sv.bp2.interval.range.set_range(bp2Begin, bp2End);
  sv.assemblyAlignIndex = 0;
  SVCandidateAssemblyData ad;
  ad.isSpanning = true;
  ad.spanningAlignments.resize(1);
  JumpAlignmentResult<int>& a(ad.spanningAlignments[0]);
  a.score = score; a.jumpInsertSize = jumpInsertSize; a.jumpRange = jumpRange;
  a.align1.beginPos = beginPos1; a.align1.apath = toPath(path1, nPath1);
  a.align2.beginPos = beginPos2; a.align2.apath = toPath(path2, nPath2);
  ad.bporient.isBp2AlignedFirst = (orient->isBp2AlignedFirst != 0);
  ad.bporient.isBp1Reversed     = (orient->isBp1Reversed != 0);
  ad.bporient.isBp2Reversed     = (orient->isBp2Reversed != 0);
  // align1 ran against bp2's window when bp2 is aligned first (:1544-1547)
  ad.bp1ref.seq() = orient->isBp2AlignedFirst ? alignRef2 : alignRef1;
  ad.bp2ref.seq() = orient->isBp2AlignedFirst ? alignRef1 : alignRef2;
  ad.bp1ref.set_offset(0);
  ad.bp2ref.set_offset(0);
  ad.extendedContigs.resize(1);
  getExtendedContig(a, std::string(contig), std::string(alignRef1), std::string(alignRef2), ad.extendedContigs[0]);
  const SVAlignmentInfo ai(sv, ad);
  out[0] = ai.bp1ContigOffset.begin_pos(); out[1] = ai.bp1ContigOffset.end_pos();
  out[2] = ai.bp2ContigOffset.begin_pos(); out[3] = ai.bp2ContigOffset.end_pos();
  out[4] = ai.bp1RefOffset.begin_pos(); out[5] = ai.bp1RefOffset.end_pos();
  out[6] = ai.bp2RefOffset.begin_pos(); out[7] = ai.bp2RefOffset.end_pos();
  out[8] = ai.isMinBpEdge(minEdge) ? 1 : 0;
  // the forward contig is the one a non-reversed breakend reads; the reversed one (if any) the other
  const bool         rev1(orient->isBp1Reversed != 0), rev2(orient->isBp2Reversed != 0);
  const std::string& fwd(rev1 ? ai.bp2ContigSeq() : ai.bp1ContigSeq());
  const std::string  rev((rev1 || rev2) ? (rev1 ? ai.bp1ContigSeq() : ai.bp2ContigSeq()) : std::string());
  if (fwd.size() + 1 > cap || rev.size() + 1 > cap) return MB200_ERR_POOL_OVERFLOW;
  std::memcpy(contigOut, fwd.c_str(), fwd.size() + 1);
  std::memcpy(revOut, rev.c_str(), rev.size() + 1);
  return MB200_OK;
}

/// The reference's per-candidate work on `threads` host threads: runIterativeAssembler on the locus' reads, then
/// every contig through [the window trim and] GlobalLargeIndelAligner (isJump == 0) or GlobalJumpAligner
/// (isJump == 1), then [contig QC + selection].  Loci are handed out dynamically (like the reference's ctpl pool
/// hands out edges, GSC/GenerateSVCandidates.cpp:252-266); each worker owns its aligner objects (they carry mutable
/// scratch, GlobalJumpAligner.hpp:117-124).  scoreSum / qcSum are checksums the GPU path must reproduce.
REF_API int ref_candidates_mt(const mb200_candidate_opts* copts, uint32_t nLoci, const uint64_t* readOff,
                              const mb200_seqview* readViews, const uint8_t* isJump, const mb200_seqview* ref1,
                              const mb200_seqview* ref2, const uint8_t* pool, const mb200_ref_cuts* cuts, int doSelect,
                              int threads, uint64_t* nContigsOut, uint64_t* cellsOut, int64_t* scoreSumOut,
                              int64_t* qcSumOut)
{
  IterativeAssemblerOptions opt;
  opt.minWordLength           = copts->asmOpts.minWordLength;
  opt.maxWordLength           = copts->asmOpts.maxWordLength;
  opt.wordStepSize            = copts->asmOpts.wordStepSize;
  opt.minCoverage             = copts->asmOpts.minCoverage;
  opt.minConservativeCoverage = copts->asmOpts.minConservativeCoverage;
  opt.maxAssemblyCount        = copts->asmOpts.maxAssemblyCount;
  opt.minUnusedReads          = copts->asmOpts.minUnusedReads;
  opt.minSupportReads         = copts->asmOpts.minSupportReads;
  if (threads < 1) threads = 1;
  std::atomic<uint32_t> next(0);
  std::atomic<uint64_t> nContigs(0), cells(0);
  std::atomic<int64_t>  scoreSum(0), qcSum(0);
  auto worker = [&]() {
    const GlobalLargeIndelAligner<int> indelAligner(toScores2(&copts->indelScores), copts->largeIndelScore);
    const GlobalJumpAligner<int>       jumpAligner(toScores2(&copts->jumpScores), copts->jumpScore);
    const AlignmentScores<int>         filterScores(toScores2(&copts->qcOpts.filterScores));
    uint64_t                           myContigs(0), myCells(0);
    int64_t                            mySum(0), myQc(0);
    for (;;) {
      const uint32_t li(next.fetch_add(1));
      if (li >= nLoci) break;
      AssemblyReadInput reads;
      for (uint64_t r(readOff[li]); r < readOff[li + 1]; ++r) {
        reads.emplace_back(reinterpret_cast<const char*>(pool + readViews[r].off), readViews[r].len);
      }
      AssemblyReadOutput      readInfo;
      SVCandidateAssemblyData ad;
      runIterativeAssembler(opt, reads, readInfo, ad.contigs);
      const unsigned    contigCount(ad.contigs.size());
      const std::string r1(reinterpret_cast<const char*>(pool + ref1[li].off), ref1[li].len);
      const char*       r2(reinterpret_cast<const char*>(pool + ref2[li].off));
      ad.smallSVAlignments.resize(contigCount);
      ad.spanningAlignments.resize(contigCount);
      for (unsigned c(0); c < contigCount; ++c) {
        const std::string& seq(ad.contigs[c].seq);
        const char*        qb(seq.data());
        const char*        qe(qb + seq.size());
        if (isJump[li]) {
          jumpAligner.align(qb, qe, r1.data(), r1.data() + r1.size(), r2, r2 + ref2[li].len, ad.spanningAlignments[c]);
          mySum += ad.spanningAlignments[c].score;
          myCells += static_cast<uint64_t>(seq.size()) * (ref1[li].len + ref2[li].len);
        } else {
          mb200_trim_result tr = {0, 0};
          if (cuts) {
            trimWindow(seq, r1, cuts[li], tr);
            myQc += ORC_QC_TERM_TRIM(tr);
          }
          AlignmentResult<int>& res(ad.smallSVAlignments[c]);
          indelAligner.align(qb, qe, r1.data() + tr.adjLeadingCut, r1.data() + r1.size() - tr.adjTrailingCut, res);
          res.align.beginPos += tr.adjLeadingCut;
          mySum += res.score;
          myCells += static_cast<uint64_t>(seq.size()) * (r1.size() - tr.adjLeadingCut - tr.adjTrailingCut);
        }
        myContigs++;
      }
      if (!doSelect) continue;
      int32_t best(-1);
      if (isJump[li]) {
        // the statics, in the order selectJumpContigDNA uses them (:1364-1398); the per-contig flags of the checksum
        // need the low-quality test on every contig
        for (unsigned c(0); c < contigCount; ++c) {
          const JumpAlignmentResult<int>& a(ad.spanningAlignments[c]);
          mb200_contig_qc                 q;
          std::memset(&q, 0, sizeof(q));
          q.isCandidate = isJumpAlignmentQCFail(a) ? 0 : 1;
          bool low(true);
          if (a.align1.isAligned() || a.align2.isAligned()) low = isLowQualityJumpAlignment(a, filterScores, false);
          q.isLowQuality = low ? 1 : 0;
          myQc += ORC_QC_TERM_CONTIG(q);
        }
        if (selectJumpContigDNA(ad, filterScores)) best = static_cast<int32_t>(ad.bestAlignmentIndex);
      } else {
        std::vector<mb200_contig_qc>  qc(contigCount);
        std::vector<mb200_aln_result> flat(contigCount);
        std::vector<uint32_t>         support(contigCount);
        for (unsigned c(0); c < contigCount; ++c) {
          const AlignmentResult<int>& res(ad.smallSVAlignments[c]);
          qcComplexContig(copts->qcOpts, res.align, ad.contigs[c].seq, r1, qc[c]);
          std::memset(&flat[c], 0, sizeof(flat[c]));
          flat[c].score    = res.score;
          flat[c].isJumped = res.isJumped ? 1 : 0;
          support[c]       = static_cast<uint32_t>(ad.contigs[c].supportReads.size());
          myQc += ORC_QC_TERM_CONTIG(qc[c]);
        }
        best = selectComplex(contigCount, flat.data(), qc.data(), support.data());
      }
      myQc += ORC_QC_TERM_LOCUS(best);
    }
    nContigs += myContigs;
    cells += myCells;
    scoreSum += mySum;
    qcSum += myQc;
  };
  std::vector<std::thread> pool_;
  for (int t(0); t < threads; ++t) pool_.emplace_back(worker);
  for (std::thread& t : pool_) t.join();
  *nContigsOut = nContigs;
  *cellsOut    = cells;
  *scoreSumOut = scoreSum;
  if (qcSumOut) *qcSumOut = qcSum;
  return MB200_OK;
}
