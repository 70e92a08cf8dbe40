// dp_align.cu -- C ABI entry points of the three aligners and the kind dispatch (kernels: dp_kernels.cuh).
#include <algorithm>
#include <cstdint>

#include "mb200_internal.h"

template <int KIND>
int mb200_dp_run_kind(mb200_ctx* ctx, const mb200_scores* scores, int extra, std::vector<DpTask>& tasks, void* dResults,
                      uint32_t* dCigar, uint64_t cigarCap, unsigned long long* dCigarUsed, int* dErrFlag,
                      uint32_t* dWorkCounters, uint64_t* cellsOut, uint64_t* algoBytesOut);
extern template int mb200_dp_run_kind<0>(mb200_ctx*, const mb200_scores*, int, std::vector<DpTask>&, void*, uint32_t*, uint64_t, unsigned long long*, int*, uint32_t*, uint64_t*, uint64_t*);
extern template int mb200_dp_run_kind<1>(mb200_ctx*, const mb200_scores*, int, std::vector<DpTask>&, void*, uint32_t*, uint64_t, unsigned long long*, int*, uint32_t*, uint64_t*, uint64_t*);
extern template int mb200_dp_run_kind<2>(mb200_ctx*, const mb200_scores*, int, std::vector<DpTask>&, void*, uint32_t*, uint64_t, unsigned long long*, int*, uint32_t*, uint64_t*, uint64_t*);

namespace {
enum { KIND_GLOBAL = MB200_KIND_GLOBAL, KIND_JUMP = MB200_KIND_JUMP, KIND_INDEL = MB200_KIND_INDEL };

template <int KIND>
int alignBatch(mb200_ctx* ctx, const mb200_scores* scores, int extra, uint32_t n, const uint8_t* pool,
               uint64_t poolBytes, const mb200_seqview* query, const mb200_seqview* ref1, const mb200_seqview* ref2,
               void* results, size_t resultBytes, uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!scores || (n && (!pool || !query || !ref1 || !results)) || !cigarUsed) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (KIND == KIND_JUMP && scores->isAllowEdgeInsertion)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "GlobalJumpAligner does not support isAllowEdgeInsertion (GlobalJumpAligner.hpp:41)");
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;  // shares device buffers with the fused path's resident results
  *cigarUsed  = 0;
  if (n == 0) return MB200_OK;
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));

  // input checks of the reference (exceptions there, status codes here)
  for (uint32_t i = 0; i < n; ++i) {
    if (query[i].len == 0) return mb200_fail(ctx, MB200_ERR_EMPTY_QUERY, "Unexpected empty query sequence");
    if (KIND == KIND_JUMP) {
      if (ref1[i].len == 0) return mb200_fail(ctx, MB200_ERR_EMPTY_REF1, "Unexpected empty reference1 sequence");
      if (ref2[i].len == 0) return mb200_fail(ctx, MB200_ERR_EMPTY_REF2, "Unexpected empty reference2 sequence");
    } else if (ref1[i].len == 0) {
      return mb200_fail(ctx, MB200_ERR_EMPTY_REF, "Unexpected empty reference sequence");
    }
    if (query[i].off + query[i].len > poolBytes || ref1[i].off + ref1[i].len > poolBytes ||
        (KIND == KIND_JUMP && ref2[i].off + ref2[i].len > poolBytes))
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "sequence view outside the pool");
  }
  int rc;
  constexpr uint64_t FRONT = MB200_POOL_FRONT;  // aligned-down TMA sources stay inside the allocation
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + FRONT + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dResults, (size_t)n * resultBytes)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dCigar, (size_t)(cigarCap ? cigarCap : 1) * sizeof(uint32_t))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dMisc, 256)) != MB200_OK) return rc;
  const uint8_t* dPool = reinterpret_cast<const uint8_t*>(ctx->dPool.ptr) + FRONT;
  std::vector<DpTask> tasks(n);
  for (uint32_t i = 0; i < n; ++i) {
    DpTask& t  = tasks[i];
    t.q        = dPool + query[i].off;
    t.r1       = dPool + ref1[i].off;
    t.r2       = (KIND == KIND_JUMP) ? dPool + ref2[i].off : t.r1;
    t.Q        = query[i].len;
    t.R1       = ref1[i].len;
    t.R2       = (KIND == KIND_JUMP) ? ref2[i].len : 0;
    t.outIndex = i;
  }
  // dMisc: [4] errFlag, [8..9] cigarUsed (u64), [32..39] class work counters
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dMisc.ptr, 0, 256, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dPool.ptr, 0, FRONT, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ctx->dPool.ptr) + FRONT, pool, poolBytes,
                                        cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(reinterpret_cast<uint8_t*>(ctx->dPool.ptr) + FRONT + poolBytes, 0, 64, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  uint32_t* misc = reinterpret_cast<uint32_t*>(ctx->dMisc.ptr);
  rc = mb200_dp_run_kind<KIND>(ctx, scores, extra, tasks, ctx->dResults.ptr, reinterpret_cast<uint32_t*>(ctx->dCigar.ptr), cigarCap,
                       reinterpret_cast<unsigned long long*>(misc + 8), reinterpret_cast<int*>(misc + 4), misc + 32,
                       &ctx->timing.cells, &ctx->timing.algoBytes);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));

  uint32_t hMisc[16];
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hMisc, ctx->dMisc.ptr, sizeof(hMisc), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(results, ctx->dResults.ptr, (size_t)n * resultBytes, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long used;
  memcpy(&used, hMisc + 8, sizeof(used));
  *cigarUsed = used;
  if (hMisc[4] != 0) return mb200_fail(ctx, (int)hMisc[4], "device-side failure in dpAlignKernel (TMA wait timeout)");
  if (used > cigarCap) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "cigar pool too small");
  if (used) {
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(cigarPool, ctx->dCigar.ptr, used * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}

}  // namespace

// device-level entry used by the fused candidates path (candidates.cu): results stay on the device
int mb200_dp_run(mb200_ctx* ctx, int kind, const mb200_scores* scores, int extra, std::vector<DpTask>& tasks,
                 void* dResults, uint32_t* dCigar, uint64_t cigarCap, unsigned long long* dCigarUsed, int* dErrFlag,
                 uint32_t* dWorkCounters, uint64_t* cellsOut, uint64_t* algoBytesOut)
{
  switch (kind) {
  case KIND_GLOBAL:
    return mb200_dp_run_kind<KIND_GLOBAL>(ctx, scores, extra, tasks, dResults, dCigar, cigarCap, dCigarUsed, dErrFlag, dWorkCounters, cellsOut, algoBytesOut);
  case KIND_JUMP:
    return mb200_dp_run_kind<KIND_JUMP>(ctx, scores, extra, tasks, dResults, dCigar, cigarCap, dCigarUsed, dErrFlag, dWorkCounters, cellsOut, algoBytesOut);
  default:
    return mb200_dp_run_kind<KIND_INDEL>(ctx, scores, extra, tasks, dResults, dCigar, cigarCap, dCigarUsed, dErrFlag, dWorkCounters, cellsOut, algoBytesOut);
  }
}

extern "C" {

int mb200_jump_align_batch(mb200_ctx* ctx, const mb200_scores* scores, int32_t jumpScore, uint32_t n,
                           const uint8_t* pool, uint64_t poolBytes, const mb200_seqview* query,
                           const mb200_seqview* ref1, const mb200_seqview* ref2, mb200_jump_result* results,
                           uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed)
{
  if (n && !ref2) return MB200_ERR_BAD_ARG;
  return alignBatch<KIND_JUMP>(ctx, scores, jumpScore, n, pool, poolBytes, query, ref1, ref2, results,
                               sizeof(mb200_jump_result), cigarPool, cigarCap, cigarUsed);
}

int mb200_indel_align_batch(mb200_ctx* ctx, const mb200_scores* scores, int32_t largeIndelScore, uint32_t n,
                            const uint8_t* pool, uint64_t poolBytes, const mb200_seqview* query,
                            const mb200_seqview* ref, mb200_aln_result* results, uint32_t* cigarPool,
                            uint64_t cigarCap, uint64_t* cigarUsed)
{
  return alignBatch<KIND_INDEL>(ctx, scores, largeIndelScore, n, pool, poolBytes, query, ref, nullptr, results,
                                sizeof(mb200_aln_result), cigarPool, cigarCap, cigarUsed);
}

int mb200_global_align_batch(mb200_ctx* ctx, const mb200_scores* scores, uint32_t n, const uint8_t* pool,
                             uint64_t poolBytes, const mb200_seqview* query, const mb200_seqview* ref,
                             mb200_aln_result* results, uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed)
{
  return alignBatch<KIND_GLOBAL>(ctx, scores, 0, n, pool, poolBytes, query, ref, nullptr, results,
                                 sizeof(mb200_aln_result), cigarPool, cigarCap, cigarUsed);
}

}  // extern "C"
