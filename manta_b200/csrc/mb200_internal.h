// mb200_internal.h -- shared host-side plumbing of libmanta_b200.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "manta_b200.h"

struct DeviceBuffer {
  void*  ptr   = nullptr;
  size_t bytes = 0;
};

struct mb200_ctx {
  int          device     = 0;
  int          smCount    = 0;
  cudaStream_t stream     = nullptr;
  cudaEvent_t  ev[4]      = {nullptr, nullptr, nullptr, nullptr};  // h2d begin, kernel begin, kernel end, d2h end
  mb200_config cfg        = {};
  mb200_timing timing     = {};
  std::string  lastError;
  // growable device buffers, reused across calls (library owns device memory)
  DeviceBuffer dPool, dTasks, dResults, dCigar, dPtrScratch, dSegScratch, dBndScratch, dMisc, dJobs, dTables;
  DeviceBuffer dAsmWork, dAsmOut0, dAsmOut1, dAsmOut2, dAsmOut3, dAsmOut4, dViews, dReadOff;
  // pinned staging
  void*  hPinned      = nullptr;
  size_t hPinnedBytes = 0;
};

#define MB200_CUDA_CHECK(ctx, expr)                                                                   \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      (ctx)->lastError = std::string(#expr) + ": " + cudaGetErrorString(_e);                          \
      return MB200_ERR_CUDA;                                                                          \
    }                                                                                                 \
  } while (0)

// grow-only device allocation
int mb200_reserve(mb200_ctx* ctx, DeviceBuffer& buf, size_t bytes);
int mb200_reserve_pinned(mb200_ctx* ctx, size_t bytes);
int mb200_fail(mb200_ctx* ctx, int status, const char* msg);
