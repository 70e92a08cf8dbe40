// split_scan.cu -- batched split-read realignment for sm_100a.
//
// Replaces (bit-exactly) splitReadAligner, GSC/SplitReadAlignment.cpp:223-350: an UNGAPPED sliding-window scan
// that keeps the first offset with the strictly greatest log-likelihood, where only bases within
// +-flankScoreSize of the breakpoint range are scored (getLnLhood :50-92), followed by the per-segment mismatch
// counts (calculateAlignScore :94-121) and the evidence flags (setEvidence :137-154).
//
// Numerics: the reference accumulates in a `float` while the table entries are `double`
// (lib/blt_util/qscore_snp.hpp:43-53): each step is  L = (float)((double)L + term)  for the table terms and a
// plain float add for the N case.  The kernel performs exactly these operations in the same order
// (ascending read position), so the result is bit-identical; the reference's early exit (:88) cannot change
// the winner because every term is <= 0.
//
// Mapping: one warp per (read, target) job; one LANE per candidate offset, 32 offsets per round.  All lanes of
// a round walk the same target position p per iteration (the target byte is a warp-wide broadcast load) and
// read their own read position j = p - offset from shared memory (consecutive lanes -> consecutive bytes).
#include <cmath>
#include <cstdint>
#include <vector>

#include "mb200_internal.h"

namespace {

constexpr int      SR_WARPS    = 4;
constexpr int      SR_MAXQ_SM  = 1024;  // read + qualities staged in shared memory up to this length
constexpr int      SR_MAXQ_TERM = 256;  // per-base double terms staged in shared memory up to this length
constexpr uint32_t SR_NONE     = 0xffffffffu;
constexpr uint32_t SR_SKIP     = 0xffffffffu;  // mb200_sr_job.table of a job that is not to be run (duplicate read)

struct SrParams {
  const uint8_t*      pool;
  const mb200_sr_job* jobs;
  mb200_sr_result*    results;
  const double*       tables;  // lnComp[2][71], lnErr[2][71]
  uint32_t*           oStar;   // [n] offset whose exact likelihood bounds the scan (SR_NONE: job finished in the first pass)
  float*              bound;   // [n]
  uint32_t            n;
  int                 flank;
  float               lnOneThird;    // std::log(1 / 3.f)  (SplitReadAlignment.cpp:60), evaluated by the host libm
  float               lnRandomBase;  // -std::log(4.f)     (SplitReadAlignment.cpp:78)
};

__device__ __forceinline__ bool isEvidenceCheck(uint32_t leftSize, uint32_t rightSize, uint32_t leftMis, uint32_t rightMis,
                                                uint32_t alignScore, uint32_t minFlankSize)
{
  if (leftSize < minFlankSize) return false;
  if (rightSize < minFlankSize) return false;
  if ((double)__fdiv_rn((float)leftMis, (float)leftSize) >= 0.25) return false;
  if ((double)__fdiv_rn((float)rightMis, (float)rightSize) >= 0.25) return false;
  const float size = (float)(leftSize + rightSize);
  if ((double)__fdiv_rn((float)alignScore, size) < 0.9) return false;
  return true;
}

__device__ __forceinline__ void scanRange(const mb200_sr_job& J, uint32_t& scanStart, uint32_t& scanEnd)
{
  const uint32_t Q = J.query.len, T = J.target.len;
  int v = J.bpBegin - (int)Q + 2;
  scanStart = (uint32_t)(v > 0 ? v : 0);
  v = (int)(T - Q);
  if (J.bpEnd < v) v = J.bpEnd;
  scanEnd = (uint32_t)(v > 0 ? v : 0);
}

// one likelihood step: L = (float)((double)L + term) for the table terms, a float add for the N case (getLnLhood :50-92)
__device__ __forceinline__ float lnStep(float L, uint8_t qb, uint8_t tb, double termMatch, double termMismatch, float lnRandomBase)
{
  if (qb != tb || qb == 'N') {
    if (qb == 'N' || tb == 'N') return __fadd_rn(L, lnRandomBase);
    return __double2float_rn(__dadd_rn((double)L, termMismatch));
  }
  return __double2float_rn(__dadd_rn((double)L, termMatch));
}

// ---- pass 1 (one warp per job): input checks, and a guess of the best offset from 16 sampled read positions.  The guess
// only decides how much the scan below can skip; any guess gives the same result.
__global__ void __launch_bounds__(SR_WARPS * 32) splitGuessKernel(const SrParams P)
{
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  for (uint32_t job = blockIdx.x * SR_WARPS + warp; job < P.n; job += gridDim.x * SR_WARPS) {
    const mb200_sr_job J = P.jobs[job];
    if (J.table == SR_SKIP) {
      if (lane == 0) {
        mb200_sr_result r;
        memset(&r, 0, sizeof(r));
        P.results[job] = r;
        P.oStar[job]   = SR_NONE;
      }
      continue;
    }
    const uint8_t* q    = P.pool + J.query.off;
    const uint8_t* qual = P.pool + J.qualOff;
    const uint8_t* t    = P.pool + J.target.off;
    const uint32_t Q = J.query.len, T = J.target.len;
    int status = MB200_OK;
    if (Q >= T) status = MB200_ERR_SPLIT_INPUT;
    bool badQual = false;  // lib/blt_util/qscore_cache.hpp:48-52 (checked for the whole read: DESIGN.md "divergences")
    for (uint32_t j = lane; j < Q; j += 32)
      if (qual[j] > MB200_MAX_QSCORE) badQual = true;
    badQual = __any_sync(0xffffffffu, badQual);
    if (status == MB200_OK && badQual) status = MB200_ERR_BAD_QUAL;
    uint32_t scanStart, scanEnd;
    scanRange(J, scanStart, scanEnd);
    if (status == MB200_OK && scanEnd < scanStart) status = MB200_ERR_SPLIT_SCAN;
    if (status != MB200_OK) {
      if (lane == 0) {
        mb200_sr_result r;
        memset(&r, 0, sizeof(r));
        r.status       = status;
        P.results[job] = r;
        P.oStar[job]   = SR_NONE;
      }
      continue;
    }
    // sampled read positions: 16 spread over the part of the read that the scoring window can cover
    uint32_t bestM = 0, bestI = scanStart;
    for (uint32_t base = scanStart; base <= scanEnd; base += 32) {
      const uint32_t i = base + lane;
      uint32_t       m = 0;
      if (i <= scanEnd) {
#pragma unroll 4
        for (uint32_t s = 0; s < 16; ++s) {
          const uint32_t j  = (uint32_t)(((uint64_t)s * Q) >> 4) + (Q >> 5);
          const uint8_t  qb = __ldg(q + j), tb = __ldg(t + i + j);
          m += (qb == tb && qb != 'N') ? 1u : 0u;
        }
        if (m > bestM) { bestM = m; bestI = i; }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint32_t om = __shfl_xor_sync(0xffffffffu, bestM, off);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, bestI, off);
      if (om > bestM || (om == bestM && oi < bestI)) { bestM = om; bestI = oi; }
    }
    if (lane == 0) P.oStar[job] = bestI;
  }
}

// ---- pass 2 (one THREAD per job): the exact likelihood of the guessed offset -- a chain of ~100 dependent
// float <- double + double steps that cannot be split (every step rounds to float), so 32 jobs share a warp here.
__global__ void __launch_bounds__(128) splitBoundKernel(const SrParams P)
{
  __shared__ double sTab[4][MB200_MAX_QSCORE + 1];
  for (int i = threadIdx.x; i < 4 * (MB200_MAX_QSCORE + 1); i += blockDim.x) (&sTab[0][0])[i] = P.tables[i];
  __syncthreads();
  const uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
  if (job >= P.n) return;
  const uint32_t i = P.oStar[job];
  if (i == SR_NONE) return;
  const mb200_sr_job J = P.jobs[job];
  const uint8_t*     q    = P.pool + J.query.off;
  const uint8_t*     qual = P.pool + J.qualOff;
  const uint8_t*     t    = P.pool + J.target.off;
  const uint32_t     Q = J.query.len, T = J.target.len;
  const double*      lnC = sTab[J.table ? 1 : 0];
  const double*      lnE = sTab[J.table ? 3 : 2];
  const double       third = (double)P.lnOneThird;
  int pLo = J.bpBegin - P.flank + 1;
  if (pLo < (int)i) pLo = (int)i;
  long long pHi = (long long)J.bpEnd + P.flank;
  if (pHi > (long long)i + Q - 1) pHi = (long long)i + Q - 1;
  if (pHi > (long long)T - 1) pHi = (long long)T - 1;
  float L = 0.f;
  for (int p = pLo; p <= (int)pHi; ++p) {
    const uint32_t j  = (uint32_t)(p - (int)i);
    const uint8_t  qb = __ldg(q + j), tb = __ldg(t + p);
    uint8_t        qv = __ldg(qual + j);
    qv                = qv < 2 ? 2 : qv;
    L                 = lnStep(L, qb, tb, lnC[qv], __dadd_rn(lnE[qv], third), P.lnRandomBase);
  }
  P.bound[job] = L;
}

// ---- pass 3 (one warp per job, one lane per offset): the scan itself.  An offset is dropped as soon as its running
// likelihood falls below the bound (every term is <= 0, so it can only fall further): the offsets that finish are exactly
// those whose likelihood is >= the bound, the guessed one among them, so the first strictly greatest one is found as
// by the full scan.  Then the per-segment mismatch counts and evidence flags (calculateAlignScore :94-121, setEvidence
// :137-154).
__global__ void __launch_bounds__(SR_WARPS * 32) splitScanKernel(const SrParams P)
{
  __shared__ double  sTab[4][MB200_MAX_QSCORE + 1];  // [0]=lnComp t0, [1]=lnComp t1, [2]=lnErr t0, [3]=lnErr t1
  __shared__ uint8_t sQ[SR_WARPS][SR_MAXQ_SM];
  __shared__ uint8_t sQual[SR_WARPS][SR_MAXQ_SM];
  __shared__ double  sTerm[SR_WARPS][2 * SR_MAXQ_TERM];  // per read base j: [2j] = lnComp[q_j], [2j+1] = lnErr[q_j] + ln(1/3)

  for (int i = threadIdx.x; i < 4 * (MB200_MAX_QSCORE + 1); i += blockDim.x)
    (&sTab[0][0])[i] = P.tables[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float lnOneThird   = P.lnOneThird;
  const float lnRandomBase = P.lnRandomBase;

  for (uint32_t job = blockIdx.x * SR_WARPS + warp; job < P.n; job += gridDim.x * SR_WARPS) {
    const uint32_t oStar = P.oStar[job];
    if (oStar == SR_NONE) continue;  // finished (status / skipped) in the first pass
    const mb200_sr_job J = P.jobs[job];
    const uint8_t*     q    = P.pool + J.query.off;
    const uint8_t*     qual = P.pool + J.qualOff;
    const uint8_t*     t    = P.pool + J.target.off;
    const uint32_t     Q = J.query.len, T = J.target.len;
    const int          bpB = J.bpBegin, bpE = J.bpEnd;
    mb200_sr_result*   out = P.results + job;
    const float        B   = P.bound[job];

    const double* lnC = sTab[J.table ? 1 : 0];
    const double* lnE = sTab[J.table ? 3 : 2];
    const bool inSm   = (Q <= SR_MAXQ_SM);
    const bool inTerm = (Q <= SR_MAXQ_TERM);
    __syncwarp();
    for (uint32_t j = lane; j < Q; j += 32) {
      uint8_t qv = qual[j];
      qv         = qv < 2 ? 2 : qv;
      if (inSm) {
        sQ[warp][j]    = q[j];
        sQual[warp][j] = qv;
      }
      if (inTerm) {
        sTerm[warp][2 * j]     = lnC[qv];
        sTerm[warp][2 * j + 1] = __dadd_rn(lnE[qv], (double)lnOneThird);
      }
    }
    __syncwarp();

    uint32_t scanStart, scanEnd;
    scanRange(J, scanStart, scanEnd);
    const int srBegin = bpB - P.flank, srEnd = bpE + P.flank;

    // the guessed offset is known exactly (pass 2): it enters the arg-max as it is and is not walked again
    float    bestL   = B;
    uint32_t bestPos = (lane == 0) ? oStar : SR_NONE;
    for (uint32_t base = scanStart; base <= scanEnd; base += 32) {
      const uint32_t i = base + lane;
      bool           alive = (i <= scanEnd) && (i != oStar);
      // every lane walks its own scored positions p in (srBegin, srEnd] with 0 <= p - i < Q, p < T, from ITS first one:
      // all lanes score a base in every iteration from the first on, so a hopeless offset is gone after a few bases
      int pCur = srBegin + 1;
      if (pCur < (int)i) pCur = (int)i;
      long long pe = (long long)srEnd;
      if (pe > (long long)i + Q - 1) pe = (long long)i + Q - 1;
      if (pe > (long long)T - 1) pe = (long long)T - 1;
      const int pEnd = (int)pe;
      float     L    = 0.f;
      if (inTerm) {
        // branch-free body (lanes that are done repeat a harmless load): ~20 instructions per scored base
        const int pSafe = (pCur <= pEnd) ? pCur : (int)i;  // i <= T - Q: always a valid target / read position pair
        while (__any_sync(0xffffffffu, alive && pCur <= pEnd)) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool    act  = alive && pCur <= pEnd;
            const int     pp   = act ? pCur : pSafe;
            const uint8_t tb   = __ldg(t + pp);
            const int     j    = pp - (int)i;
            const uint8_t qb   = sQ[warp][j];
            const bool    mism = (qb != tb);
            const double  term = sTerm[warp][2 * j + (mism ? 1 : 0)];
            const bool    isN  = (qb == 'N') || (mism && tb == 'N');  // getLnLhood :72-79
            const float   Ld   = __double2float_rn(__dadd_rn((double)L, term));
            const float   Ln   = __fadd_rn(L, lnRandomBase);
            const float   Lx   = isN ? Ln : Ld;
            L                  = act ? Lx : L;
            alive              = alive && !(act && Lx < B);  // below the bound: cannot win any more
            pCur += act ? 1 : 0;
          }
        }
      } else {
        while (__any_sync(0xffffffffu, alive && pCur <= pEnd)) {
          if (alive && pCur <= pEnd) {
            const uint8_t tb = __ldg(t + pCur);
            const int     j  = pCur - (int)i;
            uint8_t       qb, qv;
            if (inSm) { qb = sQ[warp][j]; qv = sQual[warp][j]; }
            else { qb = __ldg(q + j); qv = __ldg(qual + j); qv = qv < 2 ? 2 : qv; }
            L = lnStep(L, qb, tb, lnC[qv], __dadd_rn(lnE[qv], (double)lnOneThird), lnRandomBase);
            if (L < B) alive = false;
            ++pCur;
          }
        }
      }
      if (alive && (bestPos == SR_NONE || L > bestL || (L == bestL && i < bestPos))) {
        bestL   = L;
        bestPos = i;
      }
    }
    // warp arg-max: strictly greatest likelihood, ties -> smallest offset (first in scan order)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float    oL = __shfl_xor_sync(0xffffffffu, bestL, off);
      const uint32_t oP = __shfl_xor_sync(0xffffffffu, bestPos, off);
      if (oP != SR_NONE && (bestPos == SR_NONE || oL > bestL || (oL == bestL && oP < bestPos))) {
        bestL   = oL;
        bestPos = oP;
      }
    }

    uint32_t leftSize = 0;
    if ((int)bestPos <= bpB + 1) leftSize = (uint32_t)((bpB + 1) - (int)bestPos);
    if (leftSize > Q) {
      if (lane == 0) {
        mb200_sr_result r;
        memset(&r, 0, sizeof(r));
        r.status = MB200_ERR_SPLIT_OUTCOME;
        *out     = r;
      }
      __syncwarp();
      continue;
    }
    const uint32_t a       = Q - leftSize;
    const uint32_t b       = (uint32_t)((bpE + 1) - (int)bestPos) - leftSize;  // unsigned arithmetic, :320-322
    const uint32_t homSize = a < b ? a : b;
    const uint32_t rightSize = (leftSize + homSize < Q) ? Q - (leftSize + homSize) : 0;

    uint32_t lm = 0, hm = 0, rm = 0;
    for (uint32_t j = lane; j < Q; j += 32) {
      const uint8_t qb = inSm ? sQ[warp][j] : __ldg(q + j);
      const uint8_t tb = __ldg(t + bestPos + j);
      if (qb != tb || qb == 'N') {
        if (j <= leftSize) lm++;
        else if (j <= leftSize + homSize) hm++;
        else rm++;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lm += __shfl_xor_sync(0xffffffffu, lm, off);
      hm += __shfl_xor_sync(0xffffffffu, hm, off);
      rm += __shfl_xor_sync(0xffffffffu, rm, off);
    }
    if (lane == 0) {
      mb200_sr_result r;
      memset(&r, 0, sizeof(r));
      r.alignPos = bestPos; r.leftSize = leftSize; r.homSize = homSize; r.rightSize = rightSize;
      r.leftMismatches = lm; r.homMismatches = hm; r.rightMismatches = rm;
      r.alignScore   = Q - (lm + hm + rm);
      r.alignLnLhood = bestL;
      const bool e1 = isEvidenceCheck(leftSize, rightSize, lm, rm, r.alignScore, 16);
      const bool e2 = isEvidenceCheck(leftSize, rightSize, lm, rm, r.alignScore, 8);
      r.isEvidence = e1; r.isTier2Evidence = e2;
      r.evidence = 0.f;
      if (e1 || e2) {
        const float    size = (float)(leftSize + rightSize);
        const uint32_t mn   = leftSize < rightSize ? leftSize : rightSize;
        r.evidence          = __fdiv_rn((float)(2 * mn), size);
      }
      *out = r;
    }
    __syncwarp();
  }
}

// ---- row a11: what getReadSplitScore leaves behind for one read (incrementSplitReadEvidence, GSC/SVScorerSplit.cpp:88-128)
struct EvParams {
  const mb200_sr_result*     sr;       // 4 per read: altBp1, altBp2, refBp1, refBp2
  const mb200_split_read*    reads;
  const uint8_t*             dup;      // 1: the fragment's read was evaluated before (isSplitEvaluated)
  const uint32_t*            readFirst;
  mb200_split_read_evidence* ev;
  mb200_split_sample_counts* counts;
  uint32_t                   nReads, nCand, minMapQ, minTier2MapQ;
};

__global__ void splitEvidenceKernel(const EvParams P)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.nReads) return;
  mb200_split_read_evidence e;
  memset(&e, 0, sizeof(e));
  if (P.dup[r]) {
    e.status = MB200_SPLIT_DUPLICATE;
    P.ev[r]  = e;
    return;
  }
  const mb200_split_read rd = P.reads[r];
  e.isAnchored      = rd.mapq >= P.minMapQ;
  e.isTier2Anchored = rd.mapq >= P.minTier2MapQ;
  const mb200_sr_result* a = P.sr + 4 * (size_t)r;
  for (int k = 0; k < 4; ++k)
    if (a[k].status != MB200_OK && e.status == 0) e.status = a[k].status;
  if (e.status == 0) {
    for (int k = 0; k < 4; ++k) e.splitLnLhood[k] = a[k].alignLnLhood;
    const float altLlh = fmaxf(a[0].alignLnLhood, a[1].alignLnLhood), refLlh = fmaxf(a[2].alignLnLhood, a[3].alignLnLhood);
    for (int allele = 0; allele < 2; ++allele) {  // 0: ref (targets 2,3), 1: alt (targets 0,1)
      if (allele == 0 ? !(refLlh > altLlh) : !(altLlh > refLlh)) continue;
      const int k1 = allele == 0 ? 2 : 0, k2 = k1 + 1;
      if (a[k1].isEvidence) { e.isSplitSupport[k1] = 1; e.splitEvidence[k1] = a[k1].evidence; }
      if (a[k1].isTier2Evidence) e.isTier2SplitSupport[k1] = 1;
      if (a[k2].isEvidence) { e.isSplitSupport[k2] = 1; e.splitEvidence[k2] = a[k2].evidence; }
      if (a[k2].isTier2Evidence) e.isTier2SplitSupport[k2] = 1;
      if (a[k1].isEvidence || a[k2].isEvidence) {
        const float e1 = a[k1].isEvidence ? a[k1].evidence : 0.f, e2 = a[k2].isEvidence ? a[k2].evidence : 0.f;
        const float v  = fmaxf(e1, e2);
        if (allele == 0) { e.countedRef = 1; e.refEvidence = v; }
        else { e.countedAlt = 1; e.altEvidence = v; }
      }
    }
  }
  P.ev[r] = e;
}

// SVSampleAlleleInfo sums in read order (float accumulation is order dependent): one thread per candidate
__global__ void splitCountsKernel(const EvParams P)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.nCand) return;
  mb200_split_sample_counts s;
  memset(&s, 0, sizeof(s));
  for (uint32_t r = P.readFirst[c]; r < P.readFirst[c + 1]; ++r) {
    const mb200_split_read_evidence e = P.ev[r];
    if (e.status != 0) continue;
    const uint32_t mq = P.reads[r].mapq;
    if (e.countedRef) {
      s.ref.splitReadCount++;
      s.ref.splitReadEvidence = __fadd_rn(s.ref.splitReadEvidence, e.refEvidence);
      s.ref.splitReadMapQ     = __fadd_rn(s.ref.splitReadMapQ, (float)(mq * mq));
    }
    if (e.countedAlt) {
      s.alt.splitReadCount++;
      s.alt.splitReadEvidence = __fadd_rn(s.alt.splitReadEvidence, e.altEvidence);
      s.alt.splitReadMapQ     = __fadd_rn(s.alt.splitReadMapQ, (float)(mq * mq));
    }
  }
  P.counts[c] = s;
}

// the three passes over a job list that already sits on the device
int splitRunJobs(mb200_ctx* ctx, const uint8_t* dPool, const mb200_sr_job* dJobs, mb200_sr_result* dResults, uint32_t n,
                 uint32_t flankScoreSize)
{
  int rc;
  if ((rc = mb200_reserve(ctx, ctx->dMisc, 256)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSegScratch, (size_t)n * 8 + 64)) != MB200_OK) return rc;  // free outside the DP
  SrParams P;
  P.pool    = dPool;
  P.jobs    = dJobs;
  P.results = dResults;
  P.tables  = reinterpret_cast<const double*>(ctx->dTables.ptr);
  P.oStar   = reinterpret_cast<uint32_t*>(ctx->dSegScratch.ptr);
  P.bound   = reinterpret_cast<float*>(P.oStar + n);
  P.n       = n;
  P.flank   = (int)flankScoreSize;
  {
    // same expressions and (float) overloads as the reference; like there, the compiler folds them to the
    // correctly rounded constants (0xBF8C9F54, 0xBFB17218)
    static const float lnOneThird(std::log(1 / 3.f));
    static const float lnRandomBase(-std::log(4.f));
    P.lnOneThird   = lnOneThird;
    P.lnRandomBase = lnRandomBase;
  }
  uint32_t       grid    = (n + SR_WARPS - 1) / SR_WARPS;
  const uint32_t maxGrid = (uint32_t)ctx->smCount * 16;
  if (grid > maxGrid) grid = maxGrid;
  splitGuessKernel<<<grid, SR_WARPS * 32, 0, ctx->stream>>>(P);
  splitBoundKernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(P);
  uint32_t gridC = (n + SR_WARPS - 1) / SR_WARPS;
  if (gridC > (uint32_t)ctx->smCount * 8) gridC = (uint32_t)ctx->smCount * 8;
  splitScanKernel<<<gridC, SR_WARPS * 32, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches += 3;
  return MB200_OK;
}

}  // namespace

extern "C" int mb200_split_score_batch(mb200_ctx* ctx, const mb200_qscore_tables* tables, uint32_t flankScoreSize, uint32_t minMapQ,
                                       uint32_t minTier2MapQ, uint32_t nCandidates, const mb200_split_candidate* candidates,
                                       const uint32_t* readFirst, const mb200_split_read* reads, const uint8_t* pool,
                                       uint64_t poolBytes, mb200_sr_result* sr, mb200_split_read_evidence* ev,
                                       mb200_split_sample_counts* counts)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  ctx->timing = mb200_timing();
  if (nCandidates == 0) return MB200_OK;
  if (!tables || !candidates || !readFirst || !counts) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  const uint32_t nReads = readFirst[nCandidates];
  if (nReads && (!reads || !pool || !sr || !ev)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  // the job list (4 per read, in the reference's call order) and the isSplitEvaluated rule, on the host
  std::vector<mb200_sr_job> jobs(4 * (size_t)nReads);
  std::vector<uint8_t>      dup(nReads ? nReads : 1, 0);
  std::vector<uint64_t>     seen;
  for (uint32_t c = 0; c < nCandidates; ++c) {
    if (readFirst[c + 1] < readFirst[c] || readFirst[c + 1] > nReads) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "readFirst not monotone");
    const mb200_split_target* tg[4] = {&candidates[c].altBp1, &candidates[c].altBp2, &candidates[c].refBp1, &candidates[c].refBp2};
    for (int k = 0; k < 4; ++k)
      if (tg[k]->seq.off + tg[k]->seq.len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "split-read target outside the pool");
    seen.clear();
    for (uint32_t r = readFirst[c]; r < readFirst[c + 1]; ++r) {
      const mb200_split_read& rd = reads[r];
      if (rd.seq.off + rd.seq.len > poolBytes || rd.qualOff + rd.seq.len > poolBytes)
        return mb200_fail(ctx, MB200_ERR_BAD_ARG, "split-read read outside the pool");
      const uint64_t key = ((uint64_t)rd.fragment << 1) | (rd.isRead1 ? 1u : 0u);
      for (uint64_t s : seen)
        if (s == key) dup[r] = 1;
      if (!dup[r]) seen.push_back(key);
      for (int k = 0; k < 4; ++k) {
        mb200_sr_job& j = jobs[4 * (size_t)r + k];
        memset(&j, 0, sizeof(j));
        j.query = rd.seq; j.qualOff = rd.qualOff; j.target = tg[k]->seq; j.bpBegin = tg[k]->bpBegin; j.bpEnd = tg[k]->bpEnd;
        j.table = dup[r] ? SR_SKIP : (k < 2 ? 1u : 0u);  // dopt.altQ for the alt contig, dopt.refQ for the reference (:190-214)
        if (!dup[r]) {
          const uint64_t win = (uint64_t)j.query.len + (uint64_t)(j.bpEnd - j.bpBegin) + 2;
          ctx->timing.algoBytes += 2ull * j.query.len + (win < j.target.len ? win : j.target.len) + sizeof(mb200_sr_result);
          ctx->timing.cells += (uint64_t)j.query.len * (uint64_t)(2 * flankScoreSize + (j.bpEnd - j.bpBegin) + 1);
        }
      }
    }
  }
  int          rc;
  const size_t nr = nReads ? nReads : 1;
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dJobs, 4 * nr * sizeof(mb200_sr_job))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dResults, 4 * nr * sizeof(mb200_sr_result))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dTables, sizeof(mb200_qscore_tables))) != MB200_OK) return rc;
  DeviceBuffer* B = ctx->dLi;  // shared with the large-insertion call (calls on a ctx are serialised)
  if ((rc = mb200_reserve(ctx, B[0], nr * sizeof(mb200_split_read))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[1], nr + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[2], 4 * (size_t)(nCandidates + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[3], nr * sizeof(mb200_split_read_evidence))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[4], (size_t)nCandidates * sizeof(mb200_split_sample_counts))) != MB200_OK) return rc;
  cudaStream_t st = ctx->stream;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], st));
  if (nReads) {
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dPool.ptr, pool, poolBytes, cudaMemcpyHostToDevice, st));
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobs.ptr, jobs.data(), jobs.size() * sizeof(mb200_sr_job), cudaMemcpyHostToDevice, st));
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(B[0].ptr, reads, (size_t)nReads * sizeof(mb200_split_read), cudaMemcpyHostToDevice, st));
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(B[1].ptr, dup.data(), nReads, cudaMemcpyHostToDevice, st));
  }
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dTables.ptr, tables, sizeof(mb200_qscore_tables), cudaMemcpyHostToDevice, st));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(B[2].ptr, readFirst, 4 * (size_t)(nCandidates + 1), cudaMemcpyHostToDevice, st));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], st));
  if (nReads) {
    rc = splitRunJobs(ctx, reinterpret_cast<const uint8_t*>(ctx->dPool.ptr), reinterpret_cast<const mb200_sr_job*>(ctx->dJobs.ptr),
                      reinterpret_cast<mb200_sr_result*>(ctx->dResults.ptr), 4 * nReads, flankScoreSize);
    if (rc != MB200_OK) return rc;
  }
  EvParams E;
  E.sr = reinterpret_cast<const mb200_sr_result*>(ctx->dResults.ptr);
  E.reads = reinterpret_cast<const mb200_split_read*>(B[0].ptr);
  E.dup = reinterpret_cast<const uint8_t*>(B[1].ptr);
  E.readFirst = reinterpret_cast<const uint32_t*>(B[2].ptr);
  E.ev = reinterpret_cast<mb200_split_read_evidence*>(B[3].ptr);
  E.counts = reinterpret_cast<mb200_split_sample_counts*>(B[4].ptr);
  E.nReads = nReads; E.nCand = nCandidates; E.minMapQ = minMapQ; E.minTier2MapQ = minTier2MapQ;
  if (nReads) splitEvidenceKernel<<<(nReads + 127) / 128, 128, 0, st>>>(E);
  splitCountsKernel<<<(nCandidates + 127) / 128, 128, 0, st>>>(E);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches += 2;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], st));
  if (nReads) {
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(sr, ctx->dResults.ptr, 4 * (size_t)nReads * sizeof(mb200_sr_result), cudaMemcpyDeviceToHost, st));
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ev, B[3].ptr, (size_t)nReads * sizeof(mb200_split_read_evidence), cudaMemcpyDeviceToHost, st));
  }
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(counts, B[4].ptr, (size_t)nCandidates * sizeof(mb200_split_sample_counts), cudaMemcpyDeviceToHost, st));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], st));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}

extern "C" int mb200_split_align_batch(mb200_ctx* ctx, const mb200_qscore_tables* tables, uint32_t flankScoreSize,
                                       uint32_t n, const uint8_t* pool, uint64_t poolBytes,
                                       const mb200_sr_job* jobs, mb200_sr_result* results)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!tables || (n && (!pool || !jobs || !results))) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  ctx->timing = mb200_timing();
  if (n == 0) return MB200_OK;
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  for (uint32_t i = 0; i < n; ++i) {
    const mb200_sr_job& j = jobs[i];
    if (j.query.off + j.query.len > poolBytes || j.qualOff + j.query.len > poolBytes ||
        j.target.off + j.target.len > poolBytes)
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "split-read job view outside the pool");
    // read + quals + scanned target window + 48-byte result
    const uint64_t win = (uint64_t)j.query.len + (uint64_t)(j.bpEnd - j.bpBegin) + 2;
    ctx->timing.algoBytes += 2ull * j.query.len + (win < j.target.len ? win : j.target.len) + sizeof(mb200_sr_result);
    ctx->timing.cells += (uint64_t)j.query.len * (uint64_t)(2 * flankScoreSize + (j.bpEnd - j.bpBegin) + 1);
  }
  int rc;
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dJobs, (size_t)n * sizeof(mb200_sr_job))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dResults, (size_t)n * sizeof(mb200_sr_result))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dTables, sizeof(mb200_qscore_tables))) != MB200_OK) return rc;

  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dPool.ptr, pool, poolBytes, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobs.ptr, jobs, (size_t)n * sizeof(mb200_sr_job), cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dTables.ptr, tables, sizeof(mb200_qscore_tables), cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

  rc = splitRunJobs(ctx, reinterpret_cast<const uint8_t*>(ctx->dPool.ptr), reinterpret_cast<const mb200_sr_job*>(ctx->dJobs.ptr),
                    reinterpret_cast<mb200_sr_result*>(ctx->dResults.ptr), n, flankScoreSize);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(results, ctx->dResults.ptr, (size_t)n * sizeof(mb200_sr_result), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}
