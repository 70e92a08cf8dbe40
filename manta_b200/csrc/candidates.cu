// candidates.cu -- the fused "candidate assembled + aligned" path (BASELINE.json metric).
//
// Composition of the reference's per-candidate work in GSC/SVCandidateAssemblyRefiner.cpp: runIterativeAssembler on
// the candidate's reads (:1925 / lib/manta/SVCandidateAssembler.cpp:674,697), then every contig is aligned with
// GlobalLargeIndelAligner (complex candidates, :2032) or GlobalJumpAligner (spanning candidates, :1668).
// The contigs never leave the device between the two stages: the assembler writes them into a device pool and the
// DP tasks point straight at it; only the 40-byte contig records cross PCIe to let the host bin the DP tasks by
// query length.
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "mb200_internal.h"

extern "C" {

int mb200_candidates_upload(mb200_ctx* ctx, uint32_t nLoci, const uint64_t* readOff, const mb200_seqview* readViews,
                            const uint8_t* isJump, const mb200_seqview* ref1, const mb200_seqview* ref2,
                            const uint8_t* pool, uint64_t poolBytes, const mb200_ref_cuts* cuts)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (nLoci == 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "empty batch");
  if (!readOff || !isJump || !ref1 || !ref2 || !pool) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto& C = ctx->cand;
  if (C.nQueued == 2) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "two batches are already staged (run one first)");
  // with one batch staged take the other slot; with none, the slot the last consumed batch does not use (it stays
  // available for repeated runs until a second upload overwrites it)
  const int tgt = (C.nQueued == 1) ? 1 - C.queue[0] : (C.cur == 0 ? 1 : 0);
  if (C.cur == tgt) C.cur = -1;
  auto& S = C.slot[tgt];
  if (S.everUploaded) MB200_CUDA_CHECK(ctx, cudaEventSynchronize(S.upEnd));  // an earlier copy into this slot
  const uint64_t nReads = readOff[nLoci];
  if (nReads && !readViews) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  for (uint32_t i = 0; i < nLoci; ++i) {
    if (cuts && (cuts[i].leadingCut < 0 || cuts[i].trailingCut < 0 || cuts[i].maxLeadingCut < 0 || cuts[i].maxTrailingCut < 0))
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "negative cut");
    if (ref1[i].len == 0 || ref1[i].off + ref1[i].len > poolBytes) return mb200_fail(ctx, MB200_ERR_EMPTY_REF, "Unexpected empty reference sequence");
    if (isJump[i] && (ref2[i].len == 0 || ref2[i].off + ref2[i].len > poolBytes))
      return mb200_fail(ctx, MB200_ERR_EMPTY_REF2, "Unexpected empty reference2 sequence");
  }
  // validate the read views and size the assembler workspace now (for the word length the last run used), so that a
  // run on the resident batch starts on the device at once; the views themselves are not kept on the host
  int rc;
  if ((rc = mb200_asm_scan(ctx, C.scanWordLength, nLoci, readOff, readViews, poolBytes, &S.scan)) != MB200_OK) return rc;
  S.nLoci     = nLoci;
  S.poolBytes = poolBytes;
  S.nReads    = nReads;
  S.readOff.assign(readOff, readOff + nLoci + 1);
  S.ref1.assign(ref1, ref1 + nLoci);
  S.ref2.assign(ref2, ref2 + nLoci);
  S.isJump.assign(isJump, isJump + nLoci);
  S.cuts.clear();
  if (cuts) S.cuts.assign(cuts, cuts + nLoci);
  if ((rc = mb200_reserve(ctx, S.dPool, poolBytes + MB200_POOL_FRONT + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, S.dReadOff, 8 * (size_t)(nLoci + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, S.dViews, sizeof(mb200_seqview) * (size_t)(nReads + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, S.dIsJump, (size_t)nLoci + 4)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, S.dRef1, sizeof(mb200_seqview) * (size_t)nLoci + 16)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, S.dRef2, sizeof(mb200_seqview) * (size_t)nLoci + 16)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, S.dCuts, sizeof(mb200_ref_cuts) * (size_t)nLoci + 16)) != MB200_OK) return rc;
  // asynchronous from here on (uploadStream): the copy overlaps a run on the other slot.  The caller's buffers must
  // stay untouched until the mb200_candidates_run that consumes this batch has returned.
  cudaStream_t up = ctx->uploadStream;
  uint8_t*     dp = reinterpret_cast<uint8_t*>(S.dPool.ptr);
  MB200_CUDA_CHECK(ctx, cudaEventRecord(S.upBegin, up));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(dp, 0, MB200_POOL_FRONT, up));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(dp + MB200_POOL_FRONT, pool, poolBytes, cudaMemcpyHostToDevice, up));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(dp + MB200_POOL_FRONT + poolBytes, 0, 64, up));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(S.dReadOff.ptr, readOff, 8 * (size_t)(nLoci + 1), cudaMemcpyHostToDevice, up));
  if (nReads) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(S.dViews.ptr, readViews, sizeof(mb200_seqview) * (size_t)nReads, cudaMemcpyHostToDevice, up));
  // the small per-locus arrays are copied from the slot's own host copies (they outlive the caller's arrays)
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(S.dIsJump.ptr, S.isJump.data(), (size_t)nLoci, cudaMemcpyHostToDevice, up));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(S.dRef1.ptr, S.ref1.data(), sizeof(mb200_seqview) * (size_t)nLoci, cudaMemcpyHostToDevice, up));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(S.dRef2.ptr, S.ref2.data(), sizeof(mb200_seqview) * (size_t)nLoci, cudaMemcpyHostToDevice, up));
  if (cuts) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(S.dCuts.ptr, S.cuts.data(), sizeof(mb200_ref_cuts) * (size_t)nLoci, cudaMemcpyHostToDevice, up));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(S.upEnd, up));
  S.everUploaded = true;
  S.h2dBytes     = poolBytes + 8 * (uint64_t)(nLoci + 1) + sizeof(mb200_seqview) * nReads +
               (uint64_t)nLoci * (1 + 2 * sizeof(mb200_seqview) + (cuts ? sizeof(mb200_ref_cuts) : 0));
  C.queue[C.nQueued++] = tgt;
  return MB200_OK;
}

int mb200_candidates_run(mb200_ctx* ctx, const mb200_candidate_opts* opts, mb200_asm_locus* loci, mb200_contig* contigs,
                         mb200_seqview* pseudoReads, uint8_t* contigPool, uint64_t contigPoolCap,
                         uint64_t* contigPoolUsed, uint32_t* setPool, uint64_t setPoolCap, uint64_t* setPoolUsed,
                         mb200_aln_result* alnResults, mb200_jump_result* jumpResults, uint32_t* cigarPool,
                         uint64_t cigarCap, uint64_t* cigarUsed, mb200_trim_result* trim, mb200_contig_qc* qc, int32_t* best)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!opts || !loci || !contigs || !pseudoReads || !contigPool || !setPool || !alnResults || !jumpResults ||
      !cigarPool || !contigPoolUsed || !setPoolUsed || !cigarUsed)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  auto& C = ctx->cand;
  if (C.nQueued > 0) {  // the oldest staged batch becomes the current one
    C.cur      = C.queue[0];
    C.queue[0] = C.queue[1];
    C.queue[1] = -1;
    C.nQueued--;
  }
  if (C.cur < 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "no staged batch (call mb200_candidates_upload first)");
  auto&          c     = C.slot[C.cur];
  const uint32_t nLoci = c.nLoci;
  const bool doTrim   = !c.cuts.empty();
  const bool doSelect = (qc != nullptr) || (best != nullptr);
  if (doSelect && (!qc || !best)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "qc and best must be requested together");
  if (opts->jumpScores.isAllowEdgeInsertion) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "jump aligner: isAllowEdgeInsertion unsupported");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const uint32_t A = opts->asmOpts.maxAssemblyCount;
  if (A < 1 || A > 16) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "maxAssemblyCount outside [1,16]");
  mb200_candidate_timing& T = ctx->candTiming;
  T                         = mb200_candidate_timing();
  T.h2dBytes                = c.h2dBytes;
  ctx->timing               = mb200_timing();
  ctx->candSv.valid         = false;
  MB200_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, c.upEnd, 0));  // the batch's own upload

  int rc;
  const size_t nSlots = (size_t)nLoci * A;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut0, sizeof(mb200_asm_locus) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut1, sizeof(mb200_contig) * nSlots)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut2, sizeof(mb200_seqview) * nSlots * 2)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut3, contigPoolCap + MB200_POOL_FRONT + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut4, 4 * (size_t)setPoolCap + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAln, sizeof(mb200_aln_result) * nSlots)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dJump, sizeof(mb200_jump_result) * nSlots)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dCigar, (size_t)(cigarCap ? cigarCap : 1) * sizeof(uint32_t))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dMisc, 256)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcTrim, sizeof(mb200_trim_result) * nSlots + 16)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQc, sizeof(mb200_contig_qc) * nSlots + 16)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcBest, 4 * (size_t)nLoci + 16)) != MB200_OK) return rc;
  std::vector<mb200_trim_result> hTrim;

  uint32_t*      misc        = reinterpret_cast<uint32_t*>(ctx->dMisc.ptr);
  const uint8_t* dPool       = reinterpret_cast<const uint8_t*>(c.dPool.ptr) + MB200_POOL_FRONT;
  uint8_t*       dContigPool = reinterpret_cast<uint8_t*>(ctx->dAsmOut3.ptr) + MB200_POOL_FRONT;
  // dMisc: [0] loci counter, [4] errFlag, [8] contigPoolUsed, [10] setPoolUsed, [12] cigarUsed,
  //        [32..39] indel class work counters, [48..55] jump class work counters
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[2], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dMisc.ptr, 0, 256, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dAsmOut3.ptr, 0, MB200_POOL_FRONT, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dAsmOut1.ptr, 0, sizeof(mb200_contig) * nSlots, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dAsmOut2.ptr, 0, sizeof(mb200_seqview) * nSlots * 2, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dAln.ptr, 0, sizeof(mb200_aln_result) * nSlots, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dJump.ptr, 0, sizeof(mb200_jump_result) * nSlots, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[3], ctx->stream));
  if (!c.scan.valid || c.scan.minWordLength != opts->asmOpts.minWordLength) {
    // first word length differs from the one the batch was scanned for: fetch the views back and scan again
    std::vector<mb200_seqview> hv(c.nReads ? c.nReads : 1);
    MB200_CUDA_CHECK(ctx, cudaEventSynchronize(c.upEnd));  // the views are uploaded on uploadStream, not the legacy stream
    MB200_CUDA_CHECK(ctx, cudaMemcpy(hv.data(), c.dViews.ptr, sizeof(mb200_seqview) * (size_t)c.nReads, cudaMemcpyDeviceToHost));
    if ((rc = mb200_asm_scan(ctx, opts->asmOpts.minWordLength, nLoci, c.readOff.data(), hv.data(), c.poolBytes, &c.scan)) != MB200_OK) return rc;
    C.scanWordLength = opts->asmOpts.minWordLength;
  }
  rc = mb200_asm_run(ctx, &opts->asmOpts, nLoci, &c.scan, dPool,
                     reinterpret_cast<const uint64_t*>(c.dReadOff.ptr),
                     reinterpret_cast<const mb200_seqview*>(c.dViews.ptr),
                     reinterpret_cast<mb200_asm_locus*>(ctx->dAsmOut0.ptr), reinterpret_cast<mb200_contig*>(ctx->dAsmOut1.ptr),
                     reinterpret_cast<mb200_seqview*>(ctx->dAsmOut2.ptr), dContigPool, contigPoolCap,
                     reinterpret_cast<unsigned long long*>(misc + 8), reinterpret_cast<uint32_t*>(ctx->dAsmOut4.ptr), setPoolCap,
                     reinterpret_cast<unsigned long long*>(misc + 10), misc + 0, reinterpret_cast<int*>(misc + 4),
                     &T.asmWindows, &T.asmBytes);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[4], ctx->stream));

  // ---- slot bookkeeping for the QC, and the 10-mer anchored trim of every complex contig's window (:1984-2011)
  if (doTrim || doSelect) {
    rc = mb200_qc_prep_fused(ctx, nLoci, A, reinterpret_cast<const mb200_asm_locus*>(ctx->dAsmOut0.ptr),
                             reinterpret_cast<const mb200_contig*>(ctx->dAsmOut1.ptr), reinterpret_cast<const uint32_t*>(ctx->dAsmOut4.ptr));
    if (rc != MB200_OK) return rc;
  }
  if (doTrim) {
    MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dQcTrim.ptr, 0, sizeof(mb200_trim_result) * nSlots, ctx->stream));
    rc = mb200_trim_run(ctx, (uint32_t)nSlots, A, dContigPool, nullptr, reinterpret_cast<const mb200_contig*>(ctx->dAsmOut1.ptr), dPool,
                        reinterpret_cast<const mb200_seqview*>(c.dRef1.ptr), reinterpret_cast<const mb200_ref_cuts*>(c.dCuts.ptr),
                        reinterpret_cast<const uint8_t*>(c.dIsJump.ptr), ctx->cfg.maxContigLen ? ctx->cfg.maxContigLen : 2048,
                        reinterpret_cast<mb200_trim_result*>(ctx->dQcTrim.ptr), reinterpret_cast<int*>(misc + 4));
    if (rc != MB200_OK) return rc;
    hTrim.resize(nSlots);
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hTrim.data(), ctx->dQcTrim.ptr, sizeof(mb200_trim_result) * nSlots, cudaMemcpyDeviceToHost, ctx->stream));
  }
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cevX[0], ctx->stream));

  // ---- contig records to the host: bin the DP tasks by contig length
  uint32_t hMisc[32];
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hMisc, ctx->dMisc.ptr, sizeof(hMisc), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(loci, ctx->dAsmOut0.ptr, sizeof(mb200_asm_locus) * (size_t)nLoci, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(contigs, ctx->dAsmOut1.ptr, sizeof(mb200_contig) * nSlots, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long usedBytes, usedSets;
  memcpy(&usedBytes, hMisc + 8, 8);
  memcpy(&usedSets, hMisc + 10, 8);
  *contigPoolUsed = usedBytes;
  *setPoolUsed    = usedSets;
  if (hMisc[4] != 0) return mb200_fail(ctx, (int)hMisc[4], "assembler: a locus failed (bad base / workspace overflow / too many reads)");
  if (usedBytes > contigPoolCap || usedSets > setPoolCap) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "assembler output pools too small");
  T.asmBytes += usedBytes + 4 * usedSets;

  const bool trace = std::getenv("MB200_TRACE") != nullptr;  // host-side critical-path times to stderr
  const auto tHost0 = std::chrono::steady_clock::now();
  std::vector<DpTask> indelTasks, jumpTasks;
  indelTasks.reserve((size_t)nLoci * 4);
  jumpTasks.reserve((size_t)nLoci * 4);
  for (uint32_t li = 0; li < nLoci; ++li) {
    for (uint32_t k = 0; k < loci[li].nContigs; ++k) {
      const mb200_contig& cg = contigs[(size_t)li * A + k];
      DpTask              t;
      t.q        = dContigPool + cg.seqOff;
      t.Q        = cg.seqLen;
      t.r1       = dPool + c.ref1[li].off;
      t.R1       = c.ref1[li].len;
      t.outIndex = li * A + k;
      if (c.isJump[li]) {
        t.r2 = dPool + c.ref2[li].off;
        t.R2 = c.ref2[li].len;
        jumpTasks.push_back(t);
      } else {
        if (doTrim) {
          const mb200_trim_result& tr = hTrim[(size_t)li * A + k];
          const int64_t            len = (int64_t)t.R1 - tr.adjLeadingCut - tr.adjTrailingCut;
          if (tr.adjLeadingCut < 0 || tr.adjTrailingCut < 0 || len <= 0)
            return mb200_fail(ctx, MB200_ERR_EMPTY_REF, "Unexpected empty reference sequence");  // nothing left of the window
          t.r1 += tr.adjLeadingCut;
          t.R1 = (uint32_t)len;
        }
        t.r2 = t.r1;
        t.R2 = 0;
        indelTasks.push_back(t);
      }
      T.nContigs++;
    }
  }
  const auto tHost1 = std::chrono::steady_clock::now();
  T.nIndelTasks = (uint32_t)indelTasks.size();
  T.nJumpTasks  = (uint32_t)jumpTasks.size();
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[5], ctx->stream));
  if (!indelTasks.empty()) {
    rc = mb200_dp_run(ctx, MB200_KIND_INDEL, &opts->indelScores, opts->largeIndelScore, indelTasks, ctx->dAln.ptr,
                      reinterpret_cast<uint32_t*>(ctx->dCigar.ptr), cigarCap, reinterpret_cast<unsigned long long*>(misc + 12),
                      reinterpret_cast<int*>(misc + 4), misc + 32, &T.indelCells, &T.indelBytes);
    if (rc != MB200_OK) return rc;
  }
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[6], ctx->stream));
  const auto tHost2 = std::chrono::steady_clock::now();
  if (!jumpTasks.empty()) {
    rc = mb200_dp_run(ctx, MB200_KIND_JUMP, &opts->jumpScores, opts->jumpScore, jumpTasks, ctx->dJump.ptr,
                      reinterpret_cast<uint32_t*>(ctx->dCigar.ptr), cigarCap, reinterpret_cast<unsigned long long*>(misc + 12),
                      reinterpret_cast<int*>(misc + 4), misc + 48, &T.jumpCells, &T.jumpBytes);
    if (rc != MB200_OK) return rc;
  }
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[7], ctx->stream));
  if (trace) {
    const auto tHost3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    fprintf(stderr, "[mb200 trace] host critical path: build tasks %.2f ms, plan+launch large-indel DP %.2f ms, plan+launch jump DP %.2f ms\n",
            ms(tHost0, tHost1), ms(tHost1, tHost2), ms(tHost2, tHost3));
  }

  // ---- beginPos back to window coordinates (:2038), contig QC and bestAlignmentIndex on the resident alignments
  if (doTrim) {
    rc = mb200_apply_trim(ctx, (uint32_t)nSlots, reinterpret_cast<const uint32_t*>(ctx->dQcSlotLocus.ptr),
                          reinterpret_cast<const uint8_t*>(c.dIsJump.ptr), reinterpret_cast<const mb200_trim_result*>(ctx->dQcTrim.ptr),
                          reinterpret_cast<mb200_aln_result*>(ctx->dAln.ptr));
    if (rc != MB200_OK) return rc;
  }
  if (doSelect) {
    rc = mb200_qc_run(ctx, &opts->qcOpts, nLoci, (uint32_t)nSlots, reinterpret_cast<const uint8_t*>(c.dIsJump.ptr),
                      reinterpret_cast<const uint32_t*>(ctx->dQcSlotFirst.ptr), reinterpret_cast<const uint32_t*>(ctx->dQcNContigs.ptr),
                      reinterpret_cast<const uint32_t*>(ctx->dQcSlotLocus.ptr), reinterpret_cast<const mb200_aln_result*>(ctx->dAln.ptr),
                      reinterpret_cast<const mb200_jump_result*>(ctx->dJump.ptr), reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr),
                      dContigPool, reinterpret_cast<const mb200_seqview*>(ctx->dQcViews.ptr), dPool,
                      reinterpret_cast<const mb200_seqview*>(c.dRef1.ptr), reinterpret_cast<const uint32_t*>(ctx->dQcSupport.ptr),
                      reinterpret_cast<mb200_contig_qc*>(ctx->dQc.ptr), reinterpret_cast<int32_t*>(ctx->dQcBest.ptr),
                      reinterpret_cast<const unsigned long long*>(misc + 12), cigarCap, reinterpret_cast<const int*>(misc + 4));
    if (rc != MB200_OK) return rc;
  }
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cevX[1], ctx->stream));

  // ---- the assembler's outputs are final since the contig-record sync above: they go to the host on the copy stream
  // while the DP / QC kernels queued above run
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(pseudoReads, ctx->dAsmOut2.ptr, sizeof(mb200_seqview) * nSlots * 2, cudaMemcpyDeviceToHost, ctx->copyStream));
  if (usedBytes) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(contigPool, dContigPool, usedBytes, cudaMemcpyDeviceToHost, ctx->copyStream));
  if (usedSets) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(setPool, ctx->dAsmOut4.ptr, 4 * usedSets, cudaMemcpyDeviceToHost, ctx->copyStream));

  // ---- results to the host
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hMisc, ctx->dMisc.ptr, sizeof(hMisc), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(alnResults, ctx->dAln.ptr, sizeof(mb200_aln_result) * nSlots, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(jumpResults, ctx->dJump.ptr, sizeof(mb200_jump_result) * nSlots, cudaMemcpyDeviceToHost, ctx->stream));
  if (doSelect) {
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(qc, ctx->dQc.ptr, sizeof(mb200_contig_qc) * nSlots, cudaMemcpyDeviceToHost, ctx->stream));
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(best, ctx->dQcBest.ptr, 4 * (size_t)nLoci, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (trim) {
    if (doTrim) memcpy(trim, hTrim.data(), sizeof(mb200_trim_result) * nSlots);
    else memset(trim, 0, sizeof(mb200_trim_result) * nSlots);
  }
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long usedCig;
  memcpy(&usedCig, hMisc + 12, 8);
  *cigarUsed = usedCig;
  if (hMisc[4] != 0 || usedCig > cigarCap) {
    MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->copyStream));  // copies into the caller's buffers are in flight
    if (hMisc[4] != 0) return mb200_fail(ctx, (int)hMisc[4], "device-side failure in dpAlignKernel");
    return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "cigar pool too small");
  }
  if (usedCig) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(cigarPool, ctx->dCigar.ptr, usedCig * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->cev[0], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->copyStream));
  cudaEventElapsedTime(&T.h2dMs, c.upBegin, c.upEnd);
  cudaEventElapsedTime(&T.asmMs, ctx->cev[3], ctx->cev[4]);
  cudaEventElapsedTime(&T.indelMs, ctx->cev[5], ctx->cev[6]);
  cudaEventElapsedTime(&T.jumpMs, ctx->cev[6], ctx->cev[7]);
  cudaEventElapsedTime(&T.trimMs, ctx->cev[4], ctx->cevX[0]);
  cudaEventElapsedTime(&T.qcMs, ctx->cev[7], ctx->cevX[1]);
  cudaEventElapsedTime(&T.d2hMs, ctx->cevX[1], ctx->cev[0]);
  cudaEventElapsedTime(&T.totalMs, ctx->cev[2], ctx->cev[0]);
  T.launches = ctx->timing.launches;
  if (doSelect) {
    ctx->candSv.valid   = true;
    ctx->candSv.slot    = C.cur;
    ctx->candSv.nLoci   = nLoci;
    ctx->candSv.minSize = opts->qcOpts.minCandidateVariantSize;
  }
  T.d2hBytes = sizeof(mb200_asm_locus) * (uint64_t)nLoci + (sizeof(mb200_contig) + sizeof(mb200_aln_result) + sizeof(mb200_jump_result) + 2 * sizeof(mb200_seqview)) * (uint64_t)nSlots +
               usedBytes + 4 * usedSets + 4 * usedCig + (doTrim ? sizeof(mb200_trim_result) * (uint64_t)nSlots : 0) +
               (doSelect ? sizeof(mb200_contig_qc) * (uint64_t)nSlots + 4 * (uint64_t)nLoci : 0);
  return MB200_OK;
}

int mb200_candidates_small_svs(mb200_ctx* ctx, uint32_t* svFirst, mb200_small_sv* svs, uint64_t svCap, uint64_t* svUsed)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!svFirst || !svUsed || (svCap && !svs)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  *svUsed = 0;
  if (!ctx->candSv.valid || ctx->candSv.slot != ctx->cand.cur)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "no fused run with contig selection to refine (run mb200_candidates_run with qc/best first)");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto&          c     = ctx->cand.slot[ctx->candSv.slot];
  const uint32_t nLoci = ctx->candSv.nLoci;
  int            rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvFirst, 4 * (size_t)(nLoci + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvs, sizeof(mb200_small_sv) * (size_t)(svCap ? svCap : 1))) != MB200_OK) return rc;
  const uint8_t* dPool       = reinterpret_cast<const uint8_t*>(c.dPool.ptr) + MB200_POOL_FRONT;
  const uint8_t* dContigPool = reinterpret_cast<const uint8_t*>(ctx->dAsmOut3.ptr) + MB200_POOL_FRONT;
  rc = mb200_small_sv_run(ctx, ctx->candSv.minSize, nLoci, reinterpret_cast<const uint8_t*>(c.dIsJump.ptr),
                          reinterpret_cast<const uint32_t*>(ctx->dQcSlotFirst.ptr), reinterpret_cast<const int32_t*>(ctx->dQcBest.ptr),
                          reinterpret_cast<const mb200_contig_qc*>(ctx->dQc.ptr), reinterpret_cast<const mb200_aln_result*>(ctx->dAln.ptr),
                          reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr), dContigPool,
                          reinterpret_cast<const mb200_seqview*>(ctx->dQcViews.ptr), dPool, reinterpret_cast<const mb200_seqview*>(c.dRef1.ptr),
                          reinterpret_cast<uint32_t*>(ctx->dSvFirst.ptr), reinterpret_cast<mb200_small_sv*>(ctx->dSvs.ptr), svCap);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(svFirst, ctx->dSvFirst.ptr, 4 * (size_t)(nLoci + 1), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t used = svFirst[nLoci];
  *svUsed             = used;
  if (used > svCap) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "small SV record array too small");
  if (used) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(svs, ctx->dSvs.ptr, sizeof(mb200_small_sv) * (size_t)used, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

int mb200_candidates_jump_svs(mb200_ctx* ctx, const mb200_jump_orient* orient, mb200_jump_sv* out)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!orient || !out) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (!ctx->candSv.valid || ctx->candSv.slot != ctx->cand.cur)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "no fused run with contig selection to refine (run mb200_candidates_run with qc/best first)");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto&          c     = ctx->cand.slot[ctx->candSv.slot];
  const uint32_t nLoci = ctx->candSv.nLoci;
  int            rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvFirst, sizeof(mb200_jump_orient) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvs, sizeof(mb200_jump_sv) * (size_t)nLoci)) != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dSvFirst.ptr, orient, sizeof(mb200_jump_orient) * (size_t)nLoci, cudaMemcpyHostToDevice, ctx->stream));
  rc = mb200_jump_sv_run(ctx, nLoci, reinterpret_cast<const uint8_t*>(c.dIsJump.ptr), reinterpret_cast<const uint32_t*>(ctx->dQcSlotFirst.ptr),
                         reinterpret_cast<const int32_t*>(ctx->dQcBest.ptr), reinterpret_cast<const mb200_jump_result*>(ctx->dJump.ptr),
                         reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr), reinterpret_cast<const mb200_seqview*>(c.dRef1.ptr),
                         reinterpret_cast<const mb200_seqview*>(c.dRef2.ptr), reinterpret_cast<const mb200_jump_orient*>(ctx->dSvFirst.ptr),
                         reinterpret_cast<mb200_jump_sv*>(ctx->dSvs.ptr));
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->dSvs.ptr, sizeof(mb200_jump_sv) * (size_t)nLoci, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

int mb200_candidates_pending(mb200_ctx* ctx, uint32_t* nLoci, uint64_t* nReads, uint64_t* poolBytes)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  const auto& C   = ctx->cand;
  const int   idx = C.nQueued > 0 ? C.queue[0] : C.cur;
  if (idx < 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "no staged batch (call mb200_candidates_upload first)");
  if (nLoci) *nLoci = C.slot[idx].nLoci;
  if (nReads) *nReads = C.slot[idx].nReads;
  if (poolBytes) *poolBytes = C.slot[idx].poolBytes;
  return MB200_OK;
}

int mb200_candidates_timing(const mb200_ctx* ctx, mb200_candidate_timing* out)
{
  if (!ctx || !out) return MB200_ERR_BAD_ARG;
  *out = ctx->candTiming;
  return MB200_OK;
}

}  // extern "C"
