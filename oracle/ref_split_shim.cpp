// ref_split_shim.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libmanta_ref.so, built by oracle/build_ref.sh).
//
// Row a11 through the reference's own code: this TU #includes the UNMODIFIED GSC/SVScorerSplit.cpp where it lies and calls
// its file-static incrementSplitReadEvidence (:88-128, with incrementAlleleEvidence :42-86) on the four SRAlignmentInfo that
// the reference's splitReadAligner (GSC/SplitReadAlignment.cpp) returns for the read, writing into the reference's own
// SVFragmentEvidence / SVSampleInfo objects.  What surrounds those calls in getReadSplitScore (:131-267) needs a bam_record
// and is restated here and marked: the isSplitEvaluated test, setReadEvidence (called for real, GSC/SVScorerShared.cpp) and
// the order of the four aligner calls.  The BAM-reading members of the included file are dropped by --gc-sections.
#include "applications/GenerateSVCandidates/SVScorerSplit.cpp"

#include <cstdint>
#include <cstring>
#include <map>
#include <string>

#include "manta_b200.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
void toResult(const SRAlignmentInfo& a, mb200_sr_result& r)
{
  std::memset(&r, 0, sizeof(r));
  r.alignPos = a.alignPos; r.leftSize = a.leftSize; r.homSize = a.homSize; r.rightSize = a.rightSize;
  r.leftMismatches = a.leftMismatches; r.homMismatches = a.homMismatches; r.rightMismatches = a.rightMismatches;
  r.alignScore = a.alignScore; r.alignLnLhood = a.alignLnLhood;
  r.isEvidence = a.isEvidence; r.isTier2Evidence = a.isTier2Evidence; r.evidence = a.evidence;
}
}  // namespace

REF_API int ref_split_score_batch(const mb200_qscore_tables* /*tables: the reference builds its own*/, uint32_t flankScoreSize,
                                  uint32_t minMapQ, uint32_t minTier2MapQ, uint32_t nCandidates,
                                  const mb200_split_candidate* candidates, const uint32_t* readFirst, const mb200_split_read* reads,
                                  const uint8_t* pool, mb200_sr_result* sr, mb200_split_read_evidence* ev,
                                  mb200_split_sample_counts* counts)
{
  const qscore_snp altQ(0.), refQ(1e-3);  // CallOptionsSharedDeriv: refQ(opt.snpPrior = 1e-3), altQ(0)  GSC/SVScorer.hpp / CallOptionsShared
  for (uint32_t c(0); c < nCandidates; ++c) {
    SVEvidence::evidenceTrack_t sampleEvidence;  // std::map<std::string, SVFragmentEvidence>
    SVSampleInfo                sample;
    const mb200_split_target*   tg[4] = {&candidates[c].altBp1, &candidates[c].altBp2, &candidates[c].refBp1, &candidates[c].refBp2};
    std::string                 tseq[4];
    known_pos_range2            trange[4];
    for (int k(0); k < 4; ++k) {
      tseq[k] = std::string(reinterpret_cast<const char*>(pool + tg[k]->seq.off), tg[k]->seq.len);
      trange[k].set_range(tg[k]->bpBegin, tg[k]->bpEnd);
    }
    for (uint32_t r(readFirst[c]); r < readFirst[c + 1]; ++r) {
      mb200_split_read_evidence& e(ev[r]);
      std::memset(&e, 0, sizeof(e));
      std::memset(&sr[4 * static_cast<size_t>(r)], 0, 4 * sizeof(mb200_sr_result));
      const bool          isRead1(reads[r].isRead1 != 0);
      SVFragmentEvidence& fragment(sampleEvidence[std::to_string(reads[r].fragment)]);
      // ---- restated from getReadSplitScore :147-159
      SVFragmentEvidenceAlleleBreakendPerRead& altBp1ReadSupport(fragment.alt.bp1.getRead(isRead1));
      if (altBp1ReadSupport.isSplitEvaluated) { e.status = MB200_SPLIT_DUPLICATE; continue; }
      SVFragmentEvidenceAlleleBreakendPerRead& refBp1ReadSupport(fragment.ref.bp1.getRead(isRead1));
      SVFragmentEvidenceAlleleBreakendPerRead& altBp2ReadSupport(fragment.alt.bp2.getRead(isRead1));
      SVFragmentEvidenceAlleleBreakendPerRead& refBp2ReadSupport(fragment.ref.bp2.getRead(isRead1));
      altBp1ReadSupport.isSplitEvaluated = true;
      refBp1ReadSupport.isSplitEvaluated = true;
      altBp2ReadSupport.isSplitEvaluated = true;
      refBp2ReadSupport.isSplitEvaluated = true;
      const std::string readSeq(reinterpret_cast<const char*>(pool + reads[r].seq.off), reads[r].seq.len);
      const uint8_t*    qual(pool + reads[r].qualOff);
      SVFragmentEvidenceRead& evidenceRead(fragment.getRead(isRead1));
      setReadEvidence(minMapQ, minTier2MapQ, reads[r].mapq, readSeq.size(), false, evidenceRead);
      e.isAnchored      = evidenceRead.isAnchored(false) ? 1 : 0;
      e.isTier2Anchored = evidenceRead.isAnchored(true) ? 1 : 0;
      SRAlignmentInfo a[4];
      try {
        for (int k(0); k < 4; ++k) splitReadAligner(flankScoreSize, readSeq, (k < 2) ? altQ : refQ, qual, tseq[k], trange[k], a[k]);
      } catch (const illumina::common::GeneralException&) {
        e.status = MB200_ERR_SPLIT_INPUT;  // which of the aligner's three checks fired is compared on the oracle side
        continue;
      }
      for (int k(0); k < 4; ++k) toResult(a[k], sr[4 * static_cast<size_t>(r) + k]);
      // ---- the reference's static (:88-128)
      const SVSampleAlleleInfo refBefore(sample.ref), altBefore(sample.alt);
      incrementSplitReadEvidence(a[2], a[3], a[0], a[1], reads[r].mapq, false, sample.ref, sample.alt, refBp1ReadSupport,
                                 refBp2ReadSupport, altBp1ReadSupport, altBp2ReadSupport);
      const SVFragmentEvidenceAlleleBreakendPerRead* sup[4] = {&altBp1ReadSupport, &altBp2ReadSupport, &refBp1ReadSupport, &refBp2ReadSupport};
      for (int k(0); k < 4; ++k) {
        e.splitLnLhood[k]        = sup[k]->splitLnLhood;
        e.splitEvidence[k]       = sup[k]->splitEvidence;
        e.isSplitSupport[k]      = sup[k]->isSplitSupport ? 1 : 0;
        e.isTier2SplitSupport[k] = sup[k]->isTier2SplitSupport ? 1 : 0;
      }
      e.countedRef = (sample.ref.splitReadCount != refBefore.splitReadCount) ? 1 : 0;
      e.countedAlt = (sample.alt.splitReadCount != altBefore.splitReadCount) ? 1 : 0;
      // the value added: recomputed the way the static computes it (the accumulated sums are compared as well)
      if (e.countedRef) e.refEvidence = std::max(a[2].isEvidence ? a[2].evidence : 0.f, a[3].isEvidence ? a[3].evidence : 0.f);
      if (e.countedAlt) e.altEvidence = std::max(a[0].isEvidence ? a[0].evidence : 0.f, a[1].isEvidence ? a[1].evidence : 0.f);
    }
    counts[c].ref.splitReadCount = sample.ref.splitReadCount; counts[c].ref.splitReadEvidence = sample.ref.splitReadEvidence;
    counts[c].ref.splitReadMapQ = sample.ref.splitReadMapQ; counts[c].ref.reserved = 0;
    counts[c].alt.splitReadCount = sample.alt.splitReadCount; counts[c].alt.splitReadEvidence = sample.alt.splitReadEvidence;
    counts[c].alt.splitReadMapQ = sample.alt.splitReadMapQ; counts[c].alt.reserved = 0;
  }
  return MB200_OK;
}
