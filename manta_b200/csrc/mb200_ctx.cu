// mb200_ctx.cu -- context lifecycle of libmanta_b200.so (include/manta_b200.h).
//
// One ctx == one device + one stream + grow-only device buffers.  There is deliberately no CPU path:
// if no CUDA device of compute capability 10.x is usable, mb200_create fails (MB200_ERR_NO_DEVICE).
#include <cstdlib>

#include "mb200_internal.h"

int mb200_fail(mb200_ctx* ctx, int status, const char* msg)
{
  if (ctx) ctx->lastError = msg;
  return status;
}

int mb200_reserve(mb200_ctx* ctx, DeviceBuffer& buf, size_t bytes)
{
  if (bytes <= buf.bytes) return MB200_OK;
  if (buf.ptr) {
    MB200_CUDA_CHECK(ctx, cudaDeviceSynchronize());  // kernels on the class streams may still use it
    MB200_CUDA_CHECK(ctx, cudaFree(buf.ptr));
    buf.ptr   = nullptr;
    buf.bytes = 0;
  }
  // round up so that slowly growing batches do not reallocate every call
  size_t want = bytes + bytes / 8 + 256;
  MB200_CUDA_CHECK(ctx, cudaMalloc(&buf.ptr, want));
  buf.bytes = want;
  return MB200_OK;
}

int mb200_reserve_pinned(mb200_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->hPinnedBytes) return MB200_OK;
  if (ctx->hPinned) {
    MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    MB200_CUDA_CHECK(ctx, cudaFreeHost(ctx->hPinned));
    ctx->hPinned      = nullptr;
    ctx->hPinnedBytes = 0;
  }
  size_t want = bytes + bytes / 8 + 4096;
  MB200_CUDA_CHECK(ctx, cudaMallocHost(&ctx->hPinned, want));
  ctx->hPinnedBytes = want;
  return MB200_OK;
}

int mb200_reserve_pinned_buf(mb200_ctx* ctx, PinnedBuffer& buf, size_t bytes)
{
  if (bytes <= buf.bytes) return MB200_OK;
  if (buf.ptr) {
    MB200_CUDA_CHECK(ctx, cudaDeviceSynchronize());
    MB200_CUDA_CHECK(ctx, cudaFreeHost(buf.ptr));
    buf.ptr   = nullptr;
    buf.bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  MB200_CUDA_CHECK(ctx, cudaMallocHost(&buf.ptr, want));
  buf.bytes = want;
  return MB200_OK;
}

extern "C" {

int mb200_create(int device, const mb200_config* cfg, mb200_ctx** out)
{
  if (!out) return MB200_ERR_BAD_ARG;
  *out = nullptr;
  int         nDev = 0;
  cudaError_t e    = cudaGetDeviceCount(&nDev);
  if (e != cudaSuccess || nDev <= 0 || device < 0 || device >= nDev) return MB200_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MB200_ERR_NO_DEVICE;
  // kernels are compiled for sm_100a only
  if (prop.major != 10) return MB200_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return MB200_ERR_NO_DEVICE;

  mb200_ctx* ctx = new mb200_ctx();
  ctx->device    = device;
  ctx->smCount   = prop.multiProcessorCount;
  if (cfg) ctx->cfg = *cfg;
  ctx->forceInt32 = (ctx->cfg.reserved[0] & 1u) != 0;
  ctx->noTagKernel = (ctx->cfg.reserved[0] & 2u) != 0;
  ctx->legacyAssembler = (ctx->cfg.reserved[0] & 4u) != 0;
  ctx->forceChipAssembler = (ctx->cfg.reserved[0] & 8u) != 0;
  if (const char* e = getenv("MB200_ASM_ONCHIP")) {  // A/B switch for measurements: 0 = never, 1 = always
    ctx->legacyAssembler    = (atoi(e) == 0);
    ctx->forceChipAssembler = (atoi(e) != 0);
  }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return MB200_ERR_CUDA;
  }
  for (int i = 0; i < 4; ++i) {
    if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) {
      delete ctx;
      return MB200_ERR_CUDA;
    }
  }
  for (int i = 0; i < 8; ++i) {
    if (cudaEventCreate(&ctx->cev[i]) != cudaSuccess) {
      delete ctx;
      return MB200_ERR_CUDA;
    }
  }
  if (cudaStreamCreateWithFlags(&ctx->uploadStream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return MB200_ERR_CUDA;
  }
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreate(&ctx->cand.slot[i].upBegin) != cudaSuccess || cudaEventCreate(&ctx->cand.slot[i].upEnd) != cudaSuccess) {
      delete ctx;
      return MB200_ERR_CUDA;
    }
  }
  if (cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return MB200_ERR_CUDA;
  }
  for (int i = 0; i < 3; ++i) {
    if (cudaEventCreateWithFlags(&ctx->copyDone[i], cudaEventDisableTiming) != cudaSuccess) {
      delete ctx;
      return MB200_ERR_CUDA;
    }
  }
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreate(&ctx->cevX[i]) != cudaSuccess) {
      delete ctx;
      return MB200_ERR_CUDA;
    }
  }
  for (int i = 0; i < mb200_ctx::kDpClasses; ++i) {
    if (cudaStreamCreateWithFlags(&ctx->dpStream[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->dpDone[i], cudaEventDisableTiming) != cudaSuccess) {
      delete ctx;
      return MB200_ERR_CUDA;
    }
  }
  if (cudaEventCreateWithFlags(&ctx->dpFork, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return MB200_ERR_CUDA;
  }
  *out = ctx;
  return MB200_OK;
}

void mb200_destroy(mb200_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < mb200_ctx::kDpClasses; ++i) {
    if (ctx->dPtrC[i].ptr) cudaFree(ctx->dPtrC[i].ptr);
    if (ctx->dSegC[i].ptr) cudaFree(ctx->dSegC[i].ptr);
    if (ctx->dBndC[i].ptr) cudaFree(ctx->dBndC[i].ptr);
    if (ctx->dpStream[i]) cudaStreamDestroy(ctx->dpStream[i]);
    if (ctx->dpDone[i]) cudaEventDestroy(ctx->dpDone[i]);
  }
  for (int i = 0; i < 3; ++i) {
    if (ctx->dTasksK[i].ptr) cudaFree(ctx->dTasksK[i].ptr);
    if (ctx->hTasksK[i].ptr) cudaFreeHost(ctx->hTasksK[i].ptr);
    if (ctx->copyDone[i]) cudaEventDestroy(ctx->copyDone[i]);
  }
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  for (int i = 0; i < 2; ++i) {
    auto& S = ctx->cand.slot[i];
    DeviceBuffer* sb[] = {&S.dPool, &S.dReadOff, &S.dViews, &S.dIsJump, &S.dRef1, &S.dRef2, &S.dCuts};
    for (DeviceBuffer* b : sb)
      if (b->ptr) cudaFree(b->ptr);
    if (S.upBegin) cudaEventDestroy(S.upBegin);
    if (S.upEnd) cudaEventDestroy(S.upEnd);
  }
  if (ctx->uploadStream) cudaStreamDestroy(ctx->uploadStream);
  if (ctx->dpFork) cudaEventDestroy(ctx->dpFork);
  DeviceBuffer* bufs[] = {&ctx->dPool,    &ctx->dTasks,   &ctx->dResults, &ctx->dCigar,   &ctx->dPtrScratch,
                          &ctx->dSegScratch, &ctx->dBndScratch, &ctx->dMisc, &ctx->dJobs, &ctx->dTables,
                          &ctx->dAsmWork, &ctx->dAsmChip, &ctx->dAsmOut0, &ctx->dAsmOut1, &ctx->dAsmOut2, &ctx->dAsmOut3,
                          &ctx->dAsmOut4, &ctx->dViews,   &ctx->dReadOff, &ctx->dAln, &ctx->dJump,
                          &ctx->dQcScratch, &ctx->dQcSupport, &ctx->dQcSlotLocus, &ctx->dQcSlotFirst, &ctx->dQcNContigs,
                          &ctx->dQcViews, &ctx->dQcRef, &ctx->dQcCuts, &ctx->dQcTrim, &ctx->dQcIsJump, &ctx->dQc, &ctx->dQcBest, &ctx->dSvFirst, &ctx->dSvs};
  for (DeviceBuffer* b : bufs)
    if (b->ptr) cudaFree(b->ptr);
  for (DeviceBuffer& b : ctx->dLi)
    if (b.ptr) cudaFree(b.ptr);
  if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 8; ++i)
    if (ctx->cev[i]) cudaEventDestroy(ctx->cev[i]);
  for (int i = 0; i < 2; ++i)
    if (ctx->cevX[i]) cudaEventDestroy(ctx->cevX[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* mb200_last_error(const mb200_ctx* ctx)
{
  return ctx ? ctx->lastError.c_str() : "null ctx";
}

const char* mb200_version(void)
{
  return "manta_b200 0.1 (sm_100a; GenerateSVCandidates hot path of Illumina/manta v1.6.0)";
}

int mb200_get_timing(const mb200_ctx* ctx, mb200_timing* out)
{
  if (!ctx || !out) return MB200_ERR_BAD_ARG;
  *out = ctx->timing;
  return MB200_OK;
}

}  // extern "C"
