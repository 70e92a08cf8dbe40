/*
 * manta_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference algorithms on the hot path, one function per seam of
 * include/manta_b200.h, same argument meaning (no ctx, no poolBytes).  It exists so the CUDA path can
 * be checked bit-for-bit on the GPU box, where /root/reference (and therefore oracle/_ref) may be
 * absent.  Parity of THIS file is pinned in tests/ against (a) the golden vectors transcribed from the
 * reference's own unit tests (tests/golden/) and (b) oracle/_ref/libmanta_ref.so, the unmodified
 * reference compiled by oracle/build_ref.sh, on seeded random inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * compile, link, load or call anything declared here.  The product never does.
 */
#ifndef MANTA_ORACLE_H
#define MANTA_ORACLE_H

#include "manta_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_jump_align_batch(const mb200_scores* scores, int32_t jumpScore, uint32_t n, const uint8_t* pool,
                                 const mb200_seqview* query, const mb200_seqview* ref1,
                                 const mb200_seqview* ref2, mb200_jump_result* results, uint32_t* cigarPool,
                                 uint64_t cigarCap, uint64_t* cigarUsed);

ORC_API int orc_indel_align_batch(const mb200_scores* scores, int32_t largeIndelScore, uint32_t n,
                                  const uint8_t* pool, const mb200_seqview* query, const mb200_seqview* ref,
                                  mb200_aln_result* results, uint32_t* cigarPool, uint64_t cigarCap,
                                  uint64_t* cigarUsed);

ORC_API int orc_global_align_batch(const mb200_scores* scores, uint32_t n, const uint8_t* pool,
                                   const mb200_seqview* query, const mb200_seqview* ref,
                                   mb200_aln_result* results, uint32_t* cigarPool, uint64_t cigarCap,
                                   uint64_t* cigarUsed);

/* qscore_snp(snpPrior) tables, lib/blt_util/qscore_snp.cpp:29-41 */
ORC_API void orc_qscore_tables(double snpPrior0, double snpPrior1, mb200_qscore_tables* out);

ORC_API int orc_split_align_batch(const mb200_qscore_tables* tables, uint32_t flankScoreSize, uint32_t n,
                                  const uint8_t* pool, const mb200_sr_job* jobs, mb200_sr_result* results);

ORC_API int orc_assemble_batch(const mb200_asm_opts* opts, uint32_t nLoci, const uint64_t* readOff,
                               const mb200_seqview* readViews, const uint8_t* pool, mb200_asm_locus* loci,
                               mb200_contig* contigs, uint64_t contigCap, mb200_seqview* pseudoReads,
                               uint64_t pseudoCap, uint8_t* contigPool, uint64_t contigPoolCap,
                               uint32_t* setPool, uint64_t setPoolCap);

/* libstdc++ std::unordered_map<std::string,T> iteration-order emulation (exported for its own test) */
ORC_API void orc_unordered_order(uint32_t n, const uint8_t* keys, uint32_t keyLen, uint32_t* order);

/* 10-mer anchored reference-window trim and contig QC / selection (include/manta_b200.h, rows a5 / a9) */
ORC_API int orc_ref_trim_batch(uint32_t n, const uint8_t* pool, const mb200_seqview* contig, const mb200_seqview* ref,
                               const mb200_ref_cuts* cuts, mb200_trim_result* out);
ORC_API int orc_select_contigs_batch(const mb200_qc_opts* opts, uint32_t nLoci, const uint8_t* isJump,
                                     const uint32_t* slotFirst, const uint32_t* nContigs, uint32_t nSlots,
                                     const mb200_aln_result* aln, const mb200_jump_result* jump,
                                     const uint32_t* cigarPool, const uint8_t* pool, const mb200_seqview* contigSeq,
                                     const mb200_seqview* ref1, const uint32_t* supportCount, mb200_contig_qc* qc,
                                     int32_t* best);

ORC_API int orc_small_sv_batch(uint32_t minCandidateVariantSize, uint32_t nLoci, const uint8_t* isJump,
                               const uint32_t* slotFirst, uint32_t nSlots, const int32_t* best, const mb200_contig_qc* qc,
                               const mb200_aln_result* aln, const uint32_t* cigarPool, const uint8_t* pool,
                               const mb200_seqview* contigSeq, const mb200_seqview* ref1, uint32_t* svFirst,
                               mb200_small_sv* svs, uint64_t svCap, uint64_t* svUsed);

/* getReadSplitScore + incrementSplitReadEvidence (GSC/SVScorerSplit.cpp:42-267) on top of orc_split_align_batch */
ORC_API int orc_split_score_batch(const mb200_qscore_tables* tables, uint32_t flankScoreSize, uint32_t minMapQ,
                                  uint32_t minTier2MapQ, uint32_t nCandidates, const mb200_split_candidate* candidates,
                                  const uint32_t* readFirst, const mb200_split_read* reads, const uint8_t* pool,
                                  mb200_sr_result* sr, mb200_split_read_evidence* ev, mb200_split_sample_counts* counts);

ORC_API int orc_large_insertion_batch(const mb200_scores* edgeScores, const mb200_scores* completeScores, uint32_t nLoci,
                                      const uint8_t* isJump, const uint32_t* slotFirst, const uint32_t* nContigs,
                                      uint32_t nSlots, const mb200_aln_result* aln, const uint32_t* cigarPool,
                                      const uint8_t* pool, const mb200_seqview* contigSeq, const int32_t* consRange,
                                      const mb200_seqview* ref1, const mb200_ref_cuts* cuts, const uint32_t* svFirst,
                                      const mb200_small_sv* svs, mb200_large_insert_info* info, mb200_large_insertion* out,
                                      uint32_t* outCigar, uint64_t outCigarCap, uint64_t* outCigarUsed);

ORC_API int orc_jump_sv_batch(uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst, uint32_t nSlots,
                              const int32_t* best, const mb200_jump_result* jump, const uint32_t* cigarPool,
                              const mb200_seqview* ref1, const mb200_seqview* ref2, const mb200_jump_orient* orient,
                              mb200_jump_sv* out);

/* per-candidate composition (assemble, [trim,] align every contig, [QC + select]) on `threads` pthreads.
 * cuts == NULL: no trim; doSelect == 0: no QC.  scoreSum = sum of all alignment scores; qcSum = the checksum
 * MB200_QC_CHECKSUM_* below over trims, QC records and selections (both are reproduced from the GPU outputs). */
ORC_API int orc_candidates_mt(const mb200_candidate_opts* opts, uint32_t nLoci, const uint64_t* readOff,
                              const mb200_seqview* readViews, const uint8_t* isJump, const mb200_seqview* ref1,
                              const mb200_seqview* ref2, const uint8_t* pool, const mb200_ref_cuts* cuts, int doSelect,
                              int threads, uint64_t* nContigsOut, uint64_t* cellsOut, int64_t* scoreSumOut,
                              int64_t* qcSumOut);

/* checksum terms: per contig and per locus */
#define ORC_QC_TERM_TRIM(t) ((int64_t)(t).adjLeadingCut * 16 + (int64_t)(t).adjTrailingCut * 32)
#define ORC_QC_TERM_CONTIG(q) \
  ((int64_t)(q).isCandidate + 2 * (int64_t)(q).isLowQuality + 4 * (int64_t)(q).variantSize + 8 * (int64_t)(q).nSegments + 64 * (int64_t)(q).pathFirst + 128 * (int64_t)(q).pathLast)
#define ORC_QC_TERM_LOCUS(best) (((int64_t)(best) + 1) * 1000003)

ORC_API const char* orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
