/*
 * orc_split.c -- TEST INFRASTRUCTURE ONLY (see manta_oracle.h).
 *
 * CPU restatement of the reference's split-read realigner: an UNGAPPED sliding-window scan that keeps the
 * first offset with the strictly greatest float32 log-likelihood, scoring only bases within
 * +-flankScoreSize of the breakpoint range.
 *
 *   splitReadAligner      GSC/SplitReadAlignment.cpp:223-350
 *   getLnLhood            GSC/SplitReadAlignment.cpp:50-92   (float accumulator, double table entries)
 *   calculateAlignScore   GSC/SplitReadAlignment.cpp:94-121
 *   isEvidenceCheck       GSC/SplitReadAlignment.cpp:123-135
 *   setEvidence           GSC/SplitReadAlignment.cpp:137-154
 *   qscore_snp ctor       lib/blt_util/qscore_snp.cpp:29-41, log1p_switch lib/blt_util/math_util.hpp:35-45
 */
#include <math.h>
#include <string.h>

#include "manta_oracle.h"

void orc_qscore_tables(double snpPrior0, double snpPrior1, mb200_qscore_tables* out)
{
  const double priors[2] = {snpPrior0, snpPrior1};
  for (int t = 0; t < 2; ++t) {
    const double snp_prob  = priors[t];
    const double comp_snp3 = 1. - (snp_prob / 3.);
    for (int i = 0; i <= MB200_MAX_QSCORE; ++i) {
      const double qerr = pow(10., -((double)i) / 10.); /* phred_to_error_prob, qscore.hpp:71-74 */
      const double q2p  = (qerr * comp_snp3) + ((1 - qerr) * snp_prob);
      const double x    = -q2p;
      out->lnComp[t][i] = (fabs(x) < 0.01) ? log1p(x) : log(1 + x);
      out->lnErr[t][i]  = log(q2p);
    }
  }
}

static int is_evidence_check(const mb200_sr_result* a, uint32_t minFlankSize)
{
  if (a->leftSize < minFlankSize) return 0;
  if (a->rightSize < minFlankSize) return 0;
  /* float / then compare against the double literal 0.25 (SplitReadAlignment.cpp:128-129) */
  if ((a->leftMismatches / (float)a->leftSize) >= 0.25) return 0;
  if ((a->rightMismatches / (float)a->rightSize) >= 0.25) return 0;
  const float size = (float)(a->leftSize + a->rightSize);
  if ((a->alignScore / size) < 0.9) return 0;
  return 1;
}

static int split_one(const mb200_qscore_tables* tb, uint32_t flank, const uint8_t* pool, const mb200_sr_job* job,
                     mb200_sr_result* o)
{
  const uint8_t* q     = pool + job->query.off;
  const uint8_t* qual  = pool + job->qualOff;
  const uint8_t* t     = pool + job->target.off;
  const uint32_t Q     = job->query.len;
  const uint32_t T     = job->target.len;
  const int32_t  bpB   = job->bpBegin;
  const int32_t  bpE   = job->bpEnd;
  const double*  lnC   = tb->lnComp[job->table ? 1 : 0];
  const double*  lnE   = tb->lnErr[job->table ? 1 : 0];

  if (Q >= T) return MB200_ERR_SPLIT_INPUT;
  /* qscore_check (lib/blt_util/qscore_cache.hpp:48-52) throws for phred > 70 when a table is indexed; this
   * restatement (and the CUDA kernel) rejects the whole read up front instead -- documented restriction. */
  for (uint32_t j = 0; j < Q; ++j)
    if (qual[j] > MB200_MAX_QSCORE) return MB200_ERR_BAD_QUAL;

  int32_t v = bpB - (int32_t)Q + 2;
  const uint32_t scanStart = (uint32_t)(v > 0 ? v : 0);
  v = (int32_t)(T - Q);
  if (bpE < v) v = bpE;
  const uint32_t scanEnd = (uint32_t)(v > 0 ? v : 0);
  if (scanEnd < scanStart) return MB200_ERR_SPLIT_SCAN;

  const int32_t srBegin = bpB - (int32_t)flank;
  const int32_t srEnd   = bpE + (int32_t)flank;

  const float ln_one_third = logf(1 / 3.f);
  const float lnRandomBase = -logf(4.f);

  float    bestLnLhood = 0;
  uint32_t bestPos     = 0;
  int      isBest      = 0;
  for (uint32_t i = scanStart; i <= scanEnd; ++i) {
    float lnLhood = 0;
    for (uint32_t j = 0; j < Q; ++j) {
      int baseQual = qual[j];
      if (baseQual < 2) baseQual = 2;
      const int32_t p = (int32_t)i + (int32_t)j;
      if (p > srEnd) break;
      if (p <= srBegin) continue;
      const uint8_t tb_ = t[i + j];
      if (q[j] != tb_ || q[j] == 'N') {
        if (q[j] == 'N' || tb_ == 'N') {
          lnLhood += lnRandomBase;
        } else {
          lnLhood += lnE[baseQual] + ln_one_third; /* double + float -> double; += rounds to float */
        }
      } else {
        lnLhood += lnC[baseQual];
      }
      /* the reference's early exit (:88) cannot change the winner: every term is <= 0 */
      if (isBest && (lnLhood < bestLnLhood)) break;
    }
    if (!isBest || lnLhood > bestLnLhood) {
      bestLnLhood = lnLhood;
      bestPos     = i;
      isBest      = 1;
    }
  }

  if ((int32_t)bestPos <= bpB + 1) {
    o->leftSize = (uint32_t)((bpB + 1) - (int32_t)bestPos);
  } else {
    o->leftSize = 0;
  }
  if (o->leftSize > Q) return MB200_ERR_SPLIT_OUTCOME;
  {
    /* std::min(unsigned, pos_t-derived) is evaluated in unsigned arithmetic (SplitReadAlignment.cpp:320-322) */
    const uint32_t a = Q - o->leftSize;
    const uint32_t b = (uint32_t)((bpE + 1) - (int32_t)bestPos) - o->leftSize;
    o->homSize       = a < b ? a : b;
  }
  o->rightSize    = (o->leftSize + o->homSize < Q) ? Q - (o->leftSize + o->homSize) : 0;
  o->alignLnLhood = bestLnLhood;
  o->alignPos     = bestPos;

  o->leftMismatches = o->homMismatches = o->rightMismatches = 0;
  for (uint32_t i = 0; i < Q; ++i) {
    if (q[i] != t[bestPos + i] || q[i] == 'N') {
      if (i <= o->leftSize)
        o->leftMismatches++;
      else if (i <= o->leftSize + o->homSize)
        o->homMismatches++;
      else
        o->rightMismatches++;
    }
  }
  o->alignScore = Q - (o->leftMismatches + o->homMismatches + o->rightMismatches);

  o->isEvidence      = (uint32_t)is_evidence_check(o, 16);
  o->isTier2Evidence = (uint32_t)is_evidence_check(o, 8);
  o->evidence        = 0;
  if (o->isEvidence || o->isTier2Evidence) {
    const float    size = (float)(o->leftSize + o->rightSize);
    const uint32_t mn   = o->leftSize < o->rightSize ? o->leftSize : o->rightSize;
    o->evidence         = 2 * mn / size;
  }
  return MB200_OK;
}

int orc_split_align_batch(const mb200_qscore_tables* tables, uint32_t flankScoreSize, uint32_t n,
                          const uint8_t* pool, const mb200_sr_job* jobs, mb200_sr_result* results)
{
  for (uint32_t i = 0; i < n; ++i) {
    mb200_sr_result* o = &results[i];
    memset(o, 0, sizeof(*o));
    const int st = split_one(tables, flankScoreSize, pool, &jobs[i], o);
    if (st != MB200_OK) {
      memset(o, 0, sizeof(*o));
      o->status = st;
    }
  }
  return MB200_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * getReadSplitScore (GSC/SVScorerSplit.cpp:131-267, DNA branch) with incrementSplitReadEvidence (:88-128) and
 * incrementAlleleEvidence (:42-86); setReadEvidence GSC/SVScorerShared.cpp:26-42.
 * ------------------------------------------------------------------------------------------------------------- */
#include <stdlib.h>

static void allele_evidence(const mb200_sr_result* bp1, const mb200_sr_result* bp2, unsigned mapq, float evidence,
                            mb200_split_allele_counts* allele, mb200_split_read_evidence* e, int k1, int k2, uint8_t* counted,
                            float* added)
{
  if (bp1->isEvidence) { e->isSplitSupport[k1] = 1; e->splitEvidence[k1] = bp1->evidence; }
  if (bp1->isTier2Evidence) e->isTier2SplitSupport[k1] = 1;
  if (bp2->isEvidence) { e->isSplitSupport[k2] = 1; e->splitEvidence[k2] = bp2->evidence; }
  if (bp2->isTier2Evidence) e->isTier2SplitSupport[k2] = 1;
  if (bp1->isEvidence || bp2->isEvidence) {
    allele->splitReadCount++;
    allele->splitReadEvidence += evidence;
    allele->splitReadMapQ += mapq * mapq; /* unsigned product, then float += (:72) */
    *counted = 1;
    *added   = evidence;
  }
}

int orc_split_score_batch(const mb200_qscore_tables* tables, uint32_t flankScoreSize, uint32_t minMapQ, uint32_t minTier2MapQ,
                          uint32_t nCandidates, const mb200_split_candidate* candidates, const uint32_t* readFirst,
                          const mb200_split_read* reads, const uint8_t* pool, mb200_sr_result* sr,
                          mb200_split_read_evidence* ev, mb200_split_sample_counts* counts)
{
  for (uint32_t c = 0; c < nCandidates; ++c) {
    memset(&counts[c], 0, sizeof(counts[c]));
    const uint32_t r0 = readFirst[c], r1 = readFirst[c + 1];
    for (uint32_t r = r0; r < r1; ++r) {
      mb200_split_read_evidence* e = &ev[r];
      memset(e, 0, sizeof(*e));
      memset(&sr[4 * (size_t)r], 0, 4 * sizeof(mb200_sr_result));
      /* isSplitEvaluated (:150-159): one evaluation per (fragment, read number) and candidate */
      int dup = 0;
      for (uint32_t p = r0; p < r && !dup; ++p)
        if (reads[p].fragment == reads[r].fragment && (reads[p].isRead1 != 0) == (reads[r].isRead1 != 0)) dup = 1;
      if (dup) { e->status = MB200_SPLIT_DUPLICATE; continue; }
      e->isAnchored      = reads[r].mapq >= minMapQ;
      e->isTier2Anchored = reads[r].mapq >= minTier2MapQ;
      const mb200_split_target* tg[4] = {&candidates[c].altBp1, &candidates[c].altBp2, &candidates[c].refBp1, &candidates[c].refBp2};
      mb200_sr_job jobs[4];
      for (int k = 0; k < 4; ++k) {
        memset(&jobs[k], 0, sizeof(jobs[k]));
        jobs[k].query = reads[r].seq; jobs[k].qualOff = reads[r].qualOff; jobs[k].target = tg[k]->seq;
        jobs[k].bpBegin = tg[k]->bpBegin; jobs[k].bpEnd = tg[k]->bpEnd;
        jobs[k].table = k < 2 ? 1u : 0u; /* dopt.altQ for the alt contig, dopt.refQ for the reference (:190-214) */
      }
      mb200_sr_result* a = &sr[4 * (size_t)r];
      orc_split_align_batch(tables, flankScoreSize, 4, pool, jobs, a);
      for (int k = 0; k < 4; ++k)
        if (a[k].status != MB200_OK && e->status == 0) e->status = a[k].status;
      if (e->status != 0) continue; /* the reference throws out of the candidate here */
      /* incrementSplitReadEvidence :88-128 (isRNA == false) */
      for (int k = 0; k < 4; ++k) e->splitLnLhood[k] = a[k].alignLnLhood;
      const float altLlh = a[0].alignLnLhood > a[1].alignLnLhood ? a[0].alignLnLhood : a[1].alignLnLhood; /* std::max */
      const float refLlh = a[2].alignLnLhood > a[3].alignLnLhood ? a[2].alignLnLhood : a[3].alignLnLhood;
      if (refLlh > altLlh) {
        const float e1 = a[2].isEvidence ? a[2].evidence : 0, e2 = a[3].isEvidence ? a[3].evidence : 0;
        allele_evidence(&a[2], &a[3], reads[r].mapq, e1 > e2 ? e1 : e2, &counts[c].ref, e, 2, 3, &e->countedRef, &e->refEvidence);
      }
      if (altLlh > refLlh) {
        const float e1 = a[0].isEvidence ? a[0].evidence : 0, e2 = a[1].isEvidence ? a[1].evidence : 0;
        allele_evidence(&a[0], &a[1], reads[r].mapq, e1 > e2 ? e1 : e2, &counts[c].alt, e, 0, 1, &e->countedAlt, &e->altEvidence);
      }
    }
  }
  return MB200_OK;
}
