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
