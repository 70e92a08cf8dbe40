// gather_nccl.cu -- the one collective of the path: the per-rank candidate/result buffers are merged on
// one rank over NCCL (NVLink 5 / NVSwitch on a B200 box).  Candidate loci are independent, so nothing else crosses GPUs.
//
// Variable-length gather: every rank contributes (its byte count, the capacity it offers as root) through one
// ncclAllGather, so that all ranks take the same decision when the merged buffer would not fit; then, in one NCCL group,
// every other rank ncclSends its buffer and the root ncclRecvs each one straight into its slot of the merged buffer
// (rank order), while the root's own part is a device-to-device copy.  Buffers are DEVICE memory (the result records of
// the fused path are resident in HBM); the caller's host never touches the payload.
//
// NCCL is bound lazily with dlopen("libnccl.so.2"): a host that never gathers (1 GPU) does not need NCCL installed, and
// inside a process that already carries NCCL (torch) the same instance is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <mutex>
#include <vector>

#include "mb200_internal.h"

namespace {

struct NcclApi {
  ncclResult_t (*getUniqueId)(ncclUniqueId*)                                                    = nullptr;
  ncclResult_t (*commInitRank)(ncclComm_t*, int, ncclUniqueId, int)                             = nullptr;
  ncclResult_t (*commDestroy)(ncclComm_t)                                                       = nullptr;
  ncclResult_t (*commCount)(const ncclComm_t, int*)                                             = nullptr;
  ncclResult_t (*commUserRank)(const ncclComm_t, int*)                                          = nullptr;
  ncclResult_t (*allGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)      = nullptr;
  ncclResult_t (*recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)            = nullptr;
  ncclResult_t (*groupStart)()                                                                  = nullptr;
  ncclResult_t (*groupEnd)()                                                                    = nullptr;
  const char* (*getErrorString)(ncclResult_t)                                                   = nullptr;
  bool ok = false;
};

const NcclApi& nccl()
{
  static NcclApi        api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
#define MB200_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
    MB200_SYM(getUniqueId, "ncclGetUniqueId");
    MB200_SYM(commInitRank, "ncclCommInitRank");
    MB200_SYM(commDestroy, "ncclCommDestroy");
    MB200_SYM(commCount, "ncclCommCount");
    MB200_SYM(commUserRank, "ncclCommUserRank");
    MB200_SYM(allGather, "ncclAllGather");
    MB200_SYM(send, "ncclSend");
    MB200_SYM(recv, "ncclRecv");
    MB200_SYM(groupStart, "ncclGroupStart");
    MB200_SYM(groupEnd, "ncclGroupEnd");
    MB200_SYM(getErrorString, "ncclGetErrorString");
#undef MB200_SYM
    api.ok = api.getUniqueId && api.commInitRank && api.commDestroy && api.commCount && api.commUserRank && api.allGather &&
             api.send && api.recv && api.groupStart && api.groupEnd && api.getErrorString;
  });
  return api;
}

int ncclFail(mb200_ctx* ctx, const char* what, ncclResult_t r)
{
  if (ctx) ctx->lastError = std::string(what) + ": " + nccl().getErrorString(r);
  return MB200_ERR_CUDA;
}

}  // namespace

#define MB200_NCCL_CHECK(ctx, expr)                                  \
  do {                                                               \
    ncclResult_t _r = (expr);                                        \
    if (_r != ncclSuccess) return ncclFail(ctx, #expr, _r);          \
  } while (0)

extern "C" {

int mb200_nccl_unique_id(uint8_t id[128])
{
  if (!id) return MB200_ERR_BAD_ARG;
  if (!nccl().ok) return MB200_ERR_UNSUPPORTED;
  ncclUniqueId u;
  if (nccl().getUniqueId(&u) != ncclSuccess) return MB200_ERR_CUDA;
  static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, sizeof(u));
  return MB200_OK;
}

int mb200_nccl_comm_init(mb200_ctx* ctx, int nRanks, const uint8_t id[128], int rank, void** comm)
{
  if (!ctx || !id || !comm || nRanks < 1 || rank < 0 || rank >= nRanks) return MB200_ERR_BAD_ARG;
  if (!nccl().ok) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "libnccl.so.2 not found (needed for more than one GPU only)");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t c = nullptr;
  MB200_NCCL_CHECK(ctx, nccl().commInitRank(&c, nRanks, u, rank));
  *comm = c;
  return MB200_OK;
}

int mb200_nccl_comm_destroy(void* comm)
{
  if (!comm) return MB200_OK;
  if (!nccl().ok) return MB200_ERR_UNSUPPORTED;
  return nccl().commDestroy(static_cast<ncclComm_t>(comm)) == ncclSuccess ? MB200_OK : MB200_ERR_CUDA;
}

int mb200_gather_results(mb200_ctx* ctx, void* ncclComm, int root, const void* localDev, uint64_t localBytes, void* rootDev,
                         uint64_t rootCap, uint64_t* rankBytes)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!ncclComm || !rankBytes || (localBytes && !localDev)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (!nccl().ok) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "libnccl.so.2 not found");
  ncclComm_t comm = static_cast<ncclComm_t>(ncclComm);
  int        n = 0, me = 0;
  MB200_NCCL_CHECK(ctx, nccl().commCount(comm, &n));
  MB200_NCCL_CHECK(ctx, nccl().commUserRank(comm, &me));
  if (root < 0 || root >= n) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "root outside the communicator");
  if (me == root && rootCap && !rootDev) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int rc;
  // lengths (and the capacity each rank would offer as root) to everybody: dJobs is free outside the split-read call
  if ((rc = mb200_reserve(ctx, ctx->dJobs, sizeof(uint64_t) * 2 * (size_t)(n + 1))) != MB200_OK) return rc;
  uint64_t*             dMine = reinterpret_cast<uint64_t*>(ctx->dJobs.ptr);
  uint64_t*             dAll  = dMine + 2;
  const uint64_t        mine[2] = {localBytes, me == root ? rootCap : 0};
  std::vector<uint64_t> all(2 * (size_t)n);
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(dMine, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  MB200_NCCL_CHECK(ctx, nccl().allGather(dMine, dAll, 2, ncclUint64, comm, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(all.data(), dAll, sizeof(uint64_t) * all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t total = 0;
  for (int r = 0; r < n; ++r) {
    rankBytes[r] = all[2 * (size_t)r];
    total += rankBytes[r];
  }
  // every rank sees the same numbers, so every rank takes this exit together (no rank is left inside a send)
  if (total > all[2 * (size_t)root + 1]) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "merged result buffer larger than the root's capacity");
  MB200_NCCL_CHECK(ctx, nccl().groupStart());
  if (me != root) {
    if (localBytes) MB200_NCCL_CHECK(ctx, nccl().send(localDev, (size_t)localBytes, ncclUint8, root, comm, ctx->stream));
  } else {
    uint64_t off = 0;
    for (int r = 0; r < n; ++r) {
      uint8_t* dst = static_cast<uint8_t*>(rootDev) + off;
      if (r != root && rankBytes[r]) MB200_NCCL_CHECK(ctx, nccl().recv(dst, (size_t)rankBytes[r], ncclUint8, r, comm, ctx->stream));
      off += rankBytes[r];
    }
  }
  MB200_NCCL_CHECK(ctx, nccl().groupEnd());
  if (me == root && localBytes) {
    uint64_t off = 0;
    for (int r = 0; r < root; ++r) off += rankBytes[r];
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(rootDev) + off, localDev, (size_t)localBytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  // complete before returning: a sender that went on to launch the next batch's persistent kernels would fill every SM and
  // keep NCCL's own kernel -- and with it the root -- waiting until they drain (measured in round 1: 83 % -> 99 % scaling)
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

}  // extern "C"
