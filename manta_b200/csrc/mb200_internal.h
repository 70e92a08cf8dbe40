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

// host scan of an assembler input batch (assemble.cu: mb200_asm_scan)
struct mb200_asm_input_stats {
  uint32_t minWordLength;
  uint32_t valid;
  uint64_t maxReads, maxPackWords, maxMaskWords, maxWin, totalBases, totalWindows;
};

struct PinnedBuffer {
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
  bool         forceInt32 = false;  // cfg.reserved[0] & 1: always use the int32 one-alignment-per-warp DP kernel
  bool         noTagKernel = false;  // cfg.reserved[0] & 2: int16x2 pairs on the predicate/SEL kernel, never the tagged-score one
  // growable device buffers, reused across calls (library owns device memory)
  DeviceBuffer dPool, dTasks, dResults, dCigar, dPtrScratch, dSegScratch, dBndScratch, dMisc, dJobs, dTables;
  DeviceBuffer dLi[14];  // large_insert.cu
  bool         legacyAssembler = false;  // cfg.reserved[0] & 4: window-indexed global-memory assembler only (no on-chip kernel)
  bool         forceChipAssembler = false;  // cfg.reserved[0] & 8: on-chip assembler for every batch size
  uint32_t     asmChipMaxLoci = 0;       // batches up to this many loci take the on-chip kernel (0 = default, 3 per SM: one wave of the on-chip grid)
  DeviceBuffer dAsmChip;  // on-chip assembler: per-CTA global workspace + retry list + counters
  DeviceBuffer dAsmWork, dAsmOut0, dAsmOut1, dAsmOut2, dAsmOut3, dAsmOut4, dViews, dReadOff;
  // pinned staging
  void*  hPinned      = nullptr;
  size_t hPinnedBytes = 0;
  // staged candidate batches (mb200_candidates_upload): two slots, so that the copy of the next batch (on
  // uploadStream) overlaps the run on the current one.  Staged batches are consumed first-in first-out by
  // mb200_candidates_run; with nothing staged a run repeats on `cur`, the batch consumed last.
  struct CandSlot {
    uint32_t                    nLoci = 0;
    uint64_t                    poolBytes = 0, nReads = 0, h2dBytes = 0;
    std::vector<uint64_t>       readOff;
    std::vector<mb200_seqview>  ref1, ref2;
    std::vector<uint8_t>        isJump;
    std::vector<mb200_ref_cuts> cuts;  // empty: no window trim
    mb200_asm_input_stats       scan = {};  // of the staged reads, for word length scan.minWordLength
    DeviceBuffer                dPool, dReadOff, dViews, dIsJump, dRef1, dRef2, dCuts;
    cudaEvent_t                 upBegin = nullptr, upEnd = nullptr;
    bool                        everUploaded = false;
  };
  struct {
    CandSlot slot[2];
    int      cur = -1;
    int      queue[2] = {-1, -1};  // staged, not yet consumed (oldest first)
    int      nQueued  = 0;
    uint32_t scanWordLength = 41;  // word length the next upload scans for (the last one a run used)
  } cand;
  cudaStream_t uploadStream = nullptr;
  DeviceBuffer           dAln, dJump;
  // reference-window trim / contig QC + selection (contig_qc.cu)
  DeviceBuffer dQcScratch, dQcSupport, dQcSlotLocus, dQcSlotFirst, dQcNContigs, dQcViews, dQcRef, dQcCuts, dQcTrim, dQcIsJump,
      dQc, dQcBest, dSvFirst, dSvs;
  // what mb200_candidates_small_svs needs to know about the last fused run
  struct {
    bool     valid = false;
    int      slot = -1;
    uint32_t nLoci = 0, minSize = 0;
  } candSv;
  // DP row classes run concurrently on their own streams with their own scratch (hides the per-launch tail)
  static constexpr int   kDpClasses = 8;
  cudaStream_t           dpStream[kDpClasses] = {};
  cudaEvent_t            dpDone[kDpClasses]   = {};
  cudaEvent_t            dpFork               = nullptr;
  DeviceBuffer           dPtrC[kDpClasses], dSegC[kDpClasses], dBndC[kDpClasses];
  DeviceBuffer           dTasksK[3];
  // DP work lists are built in pinned memory and copied on their own stream, so that planning the next aligner kind
  // (and enqueueing its kernels) is not serialised behind the kernels already queued on the ctx stream
  PinnedBuffer           hTasksK[3];
  cudaStream_t           copyStream  = nullptr;
  cudaEvent_t            copyDone[3] = {nullptr, nullptr, nullptr};
  mb200_candidate_timing candTiming = {};
  cudaEvent_t            cev[8]     = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t            cevX[2]    = {nullptr, nullptr};  // end of trim, end of QC
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
int mb200_reserve_pinned_buf(mb200_ctx* ctx, PinnedBuffer& buf, size_t bytes);
int mb200_fail(mb200_ctx* ctx, int status, const char* msg);

// ---- device-level pieces shared between translation units ---------------------------------------------
#define MB200_POOL_FRONT 64  // bytes of padding in front of device sequence pools (aligned-down TMA sources)

enum { MB200_KIND_GLOBAL = 0, MB200_KIND_JUMP = 1, MB200_KIND_INDEL = 2 };

struct DpTask {
  const uint8_t *q, *r1, *r2;  // device pointers
  uint32_t       Q, R1, R2;
  uint32_t       outIndex;  // index of the result record
};

// dp_align.cu: run one aligner kind over `tasks`; results / cigar segments stay on the device
int mb200_dp_run(mb200_ctx* ctx, int kind, const mb200_scores* scores, int extra, std::vector<DpTask>& tasks,
                 void* dResults, uint32_t* dCigar, uint64_t cigarCap, unsigned long long* dCigarUsed, int* dErrFlag,
                 uint32_t* dWorkCounters, uint64_t* cellsOut, uint64_t* algoBytesOut);

// assemble.cu: host scan of the read views, and the run on device-resident inputs (outputs stay on the device)
int mb200_asm_scan(mb200_ctx* ctx, uint32_t minWordLength, uint32_t nLoci, const uint64_t* hReadOff, const mb200_seqview* hViews,
                   uint64_t poolBytes, mb200_asm_input_stats* st);
int mb200_asm_run(mb200_ctx* ctx, const mb200_asm_opts* opts, uint32_t nLoci, const mb200_asm_input_stats* st,
                  const uint8_t* dPool, const uint64_t* dReadOff,
                  const mb200_seqview* dViews, mb200_asm_locus* dLoci, mb200_contig* dContigs, mb200_seqview* dPseudo,
                  uint8_t* dContigPool, uint64_t contigPoolCap, unsigned long long* dContigPoolUsed, uint32_t* dSetPool,
                  uint64_t setPoolCap, unsigned long long* dSetPoolUsed, uint32_t* dLociCounter, int* dErrFlag,
                  uint64_t* cellsOut, uint64_t* algoBytesOut);

// contig_qc.cu: device-level trim / QC / selection (all pointers are device pointers)
int mb200_trim_run(mb200_ctx* ctx, uint32_t n, uint32_t perLocus, const uint8_t* dContigPool, const mb200_seqview* dContigView,
                   const mb200_contig* dContigRec, const uint8_t* dRefPool, const mb200_seqview* dRefView,
                   const mb200_ref_cuts* dCuts, const uint8_t* dIsJump, uint32_t maxContigLen, mb200_trim_result* dOut,
                   int* dErrFlag);
// fills ctx->dQcSupport / dQcSlotLocus / dQcSlotFirst / dQcNContigs / dQcViews from the assembler's device outputs
int mb200_qc_prep_fused(mb200_ctx* ctx, uint32_t nLoci, uint32_t A, const mb200_asm_locus* dLoci, const mb200_contig* dContigs,
                        const uint32_t* dSetPool);
int mb200_small_sv_run(mb200_ctx* ctx, uint32_t minSize, uint32_t nLoci, const uint8_t* dIsJump, const uint32_t* dSlotFirst,
                       const int32_t* dBest, const mb200_contig_qc* dQc, const mb200_aln_result* dAln, const uint32_t* dCigar,
                       const uint8_t* dContigPool, const mb200_seqview* dContigView, const uint8_t* dRefPool,
                       const mb200_seqview* dRef1, uint32_t* dSvFirst, mb200_small_sv* dSvs, uint64_t svCap);
int mb200_jump_sv_run(mb200_ctx* ctx, uint32_t nLoci, const uint8_t* dIsJump, const uint32_t* dSlotFirst, const int32_t* dBest,
                      const mb200_jump_result* dJump, const uint32_t* dCigar, const mb200_seqview* dRef1, const mb200_seqview* dRef2,
                      const mb200_jump_orient* dOrient, mb200_jump_sv* dOut);
int mb200_apply_trim(mb200_ctx* ctx, uint32_t nSlots, const uint32_t* dSlotLocus, const uint8_t* dIsJump,
                     const mb200_trim_result* dTrim, mb200_aln_result* dAln);
int mb200_qc_run(mb200_ctx* ctx, const mb200_qc_opts* opts, uint32_t nLoci, uint32_t nSlots, const uint8_t* dIsJump,
                 const uint32_t* dSlotFirst, const uint32_t* dNContigs, const uint32_t* dSlotLocus, const mb200_aln_result* dAln,
                 const mb200_jump_result* dJump, const uint32_t* dCigar, const uint8_t* dContigPool,
                 const mb200_seqview* dContigView, const uint8_t* dRefPool, const mb200_seqview* dRef1, const uint32_t* dSupport,
                 mb200_contig_qc* dQc, int32_t* dBest, const unsigned long long* dCigarUsed, uint64_t cigarCap, const int* dErrFlag);
