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

#include "mb200_internal.h"

namespace {

constexpr int SR_WARPS   = 4;
constexpr int SR_MAXQ_SM = 1024;  // read + qualities staged in shared memory up to this length

struct SrParams {
  const uint8_t*      pool;
  const mb200_sr_job* jobs;
  mb200_sr_result*    results;
  const double*       tables;  // lnComp[2][71], lnErr[2][71]
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

__global__ void __launch_bounds__(SR_WARPS * 32) splitScanKernel(const SrParams P)
{
  __shared__ double  sTab[4][MB200_MAX_QSCORE + 1];  // [0]=lnComp t0, [1]=lnComp t1, [2]=lnErr t0, [3]=lnErr t1
  __shared__ uint8_t sQ[SR_WARPS][SR_MAXQ_SM];
  __shared__ uint8_t sQual[SR_WARPS][SR_MAXQ_SM];

  for (int i = threadIdx.x; i < 4 * (MB200_MAX_QSCORE + 1); i += blockDim.x)
    (&sTab[0][0])[i] = P.tables[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float lnOneThird   = P.lnOneThird;
  const float lnRandomBase = P.lnRandomBase;

  for (uint32_t job = blockIdx.x * SR_WARPS + warp; job < P.n; job += gridDim.x * SR_WARPS) {
    const mb200_sr_job J = P.jobs[job];
    const uint8_t*     q    = P.pool + J.query.off;
    const uint8_t*     qual = P.pool + J.qualOff;
    const uint8_t*     t    = P.pool + J.target.off;
    const uint32_t     Q = J.query.len, T = J.target.len;
    const int          bpB = J.bpBegin, bpE = J.bpEnd;
    mb200_sr_result*   out = P.results + job;

    int status = MB200_OK;
    if (Q >= T) status = MB200_ERR_SPLIT_INPUT;
    // stage read + qualities (also validates qualities, lib/blt_util/qscore_cache.hpp:48-52)
    const bool inSm = (Q <= SR_MAXQ_SM);
    bool       badQual = false;
    for (uint32_t j = lane; j < Q; j += 32) {
      const uint8_t qv = qual[j];
      if (qv > MB200_MAX_QSCORE) badQual = true;
      if (inSm) {
        sQ[warp][j]    = q[j];
        sQual[warp][j] = qv < 2 ? 2 : qv;
      }
    }
    badQual = __any_sync(0xffffffffu, badQual);
    __syncwarp();
    if (status == MB200_OK && badQual) status = MB200_ERR_BAD_QUAL;

    int v = bpB - (int)Q + 2;
    const uint32_t scanStart = (uint32_t)(v > 0 ? v : 0);
    v = (int)(T - Q);
    if (bpE < v) v = bpE;
    const uint32_t scanEnd = (uint32_t)(v > 0 ? v : 0);
    if (status == MB200_OK && scanEnd < scanStart) status = MB200_ERR_SPLIT_SCAN;

    if (status != MB200_OK) {
      if (lane == 0) {
        mb200_sr_result r;
        memset(&r, 0, sizeof(r));
        r.status = status;
        *out     = r;
      }
      __syncwarp();
      continue;
    }

    const int     srBegin = bpB - P.flank, srEnd = bpE + P.flank;
    const double* lnC     = sTab[J.table ? 1 : 0];
    const double* lnE     = sTab[J.table ? 3 : 2];

    float    bestL   = 0.f;
    uint32_t bestPos = 0xffffffffu;  // none yet
    for (uint32_t base = scanStart; base <= scanEnd; base += 32) {
      const uint32_t i      = base + lane;
      const bool     inScan = (i <= scanEnd);
      // positions scored by any lane of this round: p in (srBegin, srEnd], p >= base, p < base+31+Q, p < T
      int pLo = srBegin + 1;
      if (pLo < (int)base) pLo = (int)base;
      long long pHiL = (long long)srEnd;
      const long long lim = (long long)base + 31 + (long long)Q - 1;
      if (pHiL > lim) pHiL = lim;
      if (pHiL > (long long)T - 1) pHiL = (long long)T - 1;
      const int pHi = (int)pHiL;
      float L = 0.f;
      for (int p = pLo; p <= pHi; ++p) {
        const uint8_t tb = __ldg(t + p);
        const int     j  = p - (int)i;
        if (inScan && j >= 0 && j < (int)Q) {
          uint8_t qb, qv;
          if (inSm) {
            qb = sQ[warp][j];
            qv = sQual[warp][j];
          } else {
            qb = __ldg(q + j);
            qv = __ldg(qual + j);
            qv = qv < 2 ? 2 : qv;
          }
          if (qb != tb || qb == 'N') {
            if (qb == 'N' || tb == 'N') {
              L = __fadd_rn(L, lnRandomBase);
            } else {
              L = __double2float_rn(__dadd_rn((double)L, __dadd_rn(lnE[qv], (double)lnOneThird)));
            }
          } else {
            L = __double2float_rn(__dadd_rn((double)L, lnC[qv]));
          }
        }
      }
      if (inScan && (bestPos == 0xffffffffu || L > bestL)) {
        bestL   = L;
        bestPos = i;
      }
    }
    // warp arg-max: strictly greatest likelihood, ties -> smallest offset (first in scan order)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float    oL = __shfl_xor_sync(0xffffffffu, bestL, off);
      const uint32_t oP = __shfl_xor_sync(0xffffffffu, bestPos, off);
      if (oP != 0xffffffffu && (bestPos == 0xffffffffu || oL > bestL || (oL == bestL && oP < bestPos))) {
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

}  // namespace

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

  SrParams P;
  P.pool    = reinterpret_cast<const uint8_t*>(ctx->dPool.ptr);
  P.jobs    = reinterpret_cast<const mb200_sr_job*>(ctx->dJobs.ptr);
  P.results = reinterpret_cast<mb200_sr_result*>(ctx->dResults.ptr);
  P.tables  = reinterpret_cast<const double*>(ctx->dTables.ptr);
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
  uint32_t grid = (n + SR_WARPS - 1) / SR_WARPS;
  const uint32_t maxGrid = (uint32_t)ctx->smCount * 16;
  if (grid > maxGrid) grid = maxGrid;
  splitScanKernel<<<grid, SR_WARPS * 32, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches = 1;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(results, ctx->dResults.ptr, (size_t)n * sizeof(mb200_sr_result), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}
