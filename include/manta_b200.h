/*
 * manta_b200.h -- C ABI of the B200-native GenerateSVCandidates hot path.
 *
 * This is the drop-in boundary ("B2 -- in-process C++ seams"):
 * each entry point below is the batched replacement of one synchronous C++ call in
 * Illumina/manta.  Citations are path:line under the reference tree
 * (src/c++/lib/ abbreviated lib/, lib/applications/GenerateSVCandidates/ abbreviated GSC/).
 *
 *   mb200_jump_align_batch   <- GlobalJumpAligner<int>::align        lib/alignment/GlobalJumpAligner.hpp:45-53
 *                                called from GSC/SVCandidateAssemblyRefiner.cpp:1668,1706
 *   mb200_indel_align_batch  <- GlobalLargeIndelAligner<int>::align  lib/alignment/GlobalLargeIndelAligner.hpp:48-54
 *                                called from GSC/SVCandidateAssemblyRefiner.cpp:2032
 *   mb200_global_align_batch <- GlobalAligner<int>::align            lib/alignment/GlobalAligner.hpp (same shape)
 *                                called from GSC/SVCandidateAssemblyRefiner.cpp:933, GSC/SVScorePairAltProcessor.cpp:211
 *   mb200_split_align_batch  <- splitReadAligner                     GSC/SplitReadAlignment.hpp:63-70
 *                                called from GSC/SVScorerSplit.cpp:190-226
 *   mb200_assemble_batch     <- runIterativeAssembler                lib/assembly/IterativeAssembler.hpp:43-47
 *                                called from lib/manta/SVCandidateAssembler.cpp:674,697
 *   mb200_candidates_run     <- the assemble -> align-every-contig composition of
 *                                GSC/SVCandidateAssemblyRefiner.cpp:1422-1743 (spanning) and :1925-2039 (complex),
 *                                kept resident on the device (BASELINE.json metric "candidates assembled+aligned").
 *
 * Conventions
 *   - plain C types only; all pointers are HOST pointers unless the name ends in _dev.
 *   - every call returns an int status (MB200_OK == 0).  On failure mb200_last_error(ctx)
 *     returns a message; the C++ shim (include/manta_b200_shim.hpp) turns it into the same
 *     exception type the reference throws.
 *   - caller owns host buffers, the library owns device memory.  Calls on one ctx must be
 *     serialised by the caller (one ctx per host driver thread), no global state.
 *   - there is NO CPU fallback: if the CUDA device or kernels are unavailable, mb200_create fails.
 *   - sequences are ASCII bytes in a caller-provided pool, addressed by (offset,length) views.
 *     Aligners compare raw bytes (any alphabet; N==N is a match in the DP, lib/alignment/GlobalJumpAlignerImpl.hpp:132).
 *
 * Alignment paths (CIGARs) are returned run-length encoded, one uint32 per segment:
 *      segment = (length << 4) | type,   type = ALIGNPATH::align_t  (lib/blt_util/align_path.hpp:34)
 */
#ifndef MANTA_B200_H
#define MANTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_API __attribute__((visibility("default")))

/* ---- status codes ---------------------------------------------------------------------- */
enum {
  MB200_OK                = 0,
  MB200_ERR_CUDA          = 1,  /* CUDA runtime / launch failure (fatal, like any exception in the reference) */
  MB200_ERR_NO_DEVICE     = 2,  /* no usable sm_100 device -- there is no CPU fallback */
  MB200_ERR_EMPTY_QUERY   = 3,  /* "Unexpected empty query sequence"      GlobalJumpAlignerImpl.hpp:50 */
  MB200_ERR_EMPTY_REF     = 4,  /* "Unexpected empty reference sequence"  GlobalAlignerImpl.hpp:47 */
  MB200_ERR_EMPTY_REF1    = 5,  /* "Unexpected empty reference1 sequence" GlobalJumpAlignerImpl.hpp:53 */
  MB200_ERR_EMPTY_REF2    = 6,  /* "Unexpected empty reference2 sequence" GlobalJumpAlignerImpl.hpp:56 */
  MB200_ERR_SPLIT_INPUT   = 7,  /* querySize >= targetSize                SplitReadAlignment.cpp:236 */
  MB200_ERR_SPLIT_SCAN    = 8,  /* scanEnd < scanStart                    SplitReadAlignment.cpp:264 */
  MB200_ERR_SPLIT_OUTCOME = 9,  /* leftSize > querySize                   SplitReadAlignment.cpp:305 */
  MB200_ERR_BAD_QUAL      = 10, /* basecall quality > 70 or < 0           lib/blt_util/qscore_cache.hpp:48-52 */
  MB200_ERR_POOL_OVERFLOW = 11, /* caller-provided output pool too small (call again with a larger pool) */
  MB200_ERR_BAD_ARG       = 12,
  MB200_ERR_BAD_BASE      = 13, /* assembler input byte outside {A,C,G,T,N} (documented restriction) */
  MB200_ERR_ASM_OVERFLOW  = 14, /* assembler workspace cap exceeded for some locus */
  MB200_ERR_UNSUPPORTED   = 15,
  MB200_ERR_SV_RANGE      = 16  /* "Attempting to convert alignment to sv candidate": cipos of a precise variant does
                                   not start at 0                            GSC/SVCandidateAssemblyRefiner.cpp:768 */
};

/* ---- alignment path segment types == ALIGNPATH::align_t -------------------------------- */
enum {
  MB200_SEG_NONE = 0,
  MB200_SEG_MATCH = 1,
  MB200_SEG_INSERT = 2,
  MB200_SEG_DELETE = 3,
  MB200_SEG_SKIP = 4,
  MB200_SEG_SOFT_CLIP = 5,
  MB200_SEG_HARD_CLIP = 6,
  MB200_SEG_PAD = 7,
  MB200_SEG_SEQ_MATCH = 8,
  MB200_SEG_SEQ_MISMATCH = 9
};
#define MB200_SEG_TYPE(s) ((uint32_t)(s) & 0xFu)
#define MB200_SEG_LEN(s) ((uint32_t)(s) >> 4)
#define MB200_SEG_MAKE(len, type) ((((uint32_t)(len)) << 4) | ((uint32_t)(type)))

/* ---- common structs -------------------------------------------------------------------- */

/* AlignmentScores<int>  lib/alignment/AlignmentScores.hpp:23-57 */
typedef struct mb200_scores {
  int32_t match;
  int32_t mismatch;
  int32_t open;
  int32_t extend;
  int32_t offEdge;
  int32_t isAllowEdgeInsertion; /* must be 0 for the jump aligner (GlobalJumpAligner.hpp:41) */
} mb200_scores;

/* a byte range of the caller's sequence pool */
typedef struct mb200_seqview {
  uint64_t off;
  uint32_t len;
  uint32_t reserved;
} mb200_seqview;

/* AlignmentResult<int>  lib/alignment/SingleRefAlignerShared.hpp:34-47 */
typedef struct mb200_aln_result {
  int32_t  score;
  int32_t  beginPos;
  uint32_t isJumped;
  uint32_t cigarLen; /* number of path segments */
  uint64_t cigarOff; /* first segment, in uint32 units, inside the caller's cigar pool */
} mb200_aln_result;

/* JumpAlignmentResult<int>  lib/alignment/JumpAlignerBase.hpp:37-56 */
typedef struct mb200_jump_result {
  int32_t  score;
  uint32_t jumpInsertSize;
  uint32_t jumpRange;
  int32_t  beginPos1;
  int32_t  beginPos2;
  uint32_t cigarLen1;
  uint32_t cigarLen2;
  uint32_t reserved;
  uint64_t cigarOff1;
  uint64_t cigarOff2;
} mb200_jump_result;

/* SRAlignmentInfo  GSC/SplitReadAlignment.hpp:35-49 */
typedef struct mb200_sr_result {
  uint32_t alignPos;
  uint32_t leftSize;
  uint32_t homSize;
  uint32_t rightSize;
  uint32_t leftMismatches;
  uint32_t homMismatches;
  uint32_t rightMismatches;
  uint32_t alignScore;
  float    alignLnLhood;
  uint32_t isEvidence;
  uint32_t isTier2Evidence;
  float    evidence;
  int32_t  status; /* MB200_OK or the MB200_ERR_SPLIT_* / MB200_ERR_BAD_QUAL the reference would have thrown for this item */
  uint32_t reserved;
} mb200_sr_result;

/* one (read, target) split-read job  GSC/SVScorerSplit.cpp:190-226 */
typedef struct mb200_sr_job {
  mb200_seqview query;   /* read bases (ASCII) in the pool */
  uint64_t      qualOff; /* query.len basecall qualities (raw phred bytes) in the pool */
  mb200_seqview target;  /* target sequence in the pool */
  int32_t       bpBegin; /* targetBpOffsetRange.begin_pos() */
  int32_t       bpEnd;   /* targetBpOffsetRange.end_pos()   */
  uint32_t      table;   /* which qscore_snp table pair: 0 or 1 (ref allele / alt allele, GSC/SVScorer.hpp:46-51) */
  uint32_t      reserved;
} mb200_sr_job;

/* qscore_snp tables  lib/blt_util/qscore_snp.hpp:43-53 (71 doubles each, index = phred) */
#define MB200_MAX_QSCORE 70
typedef struct mb200_qscore_tables {
  double lnComp[2][MB200_MAX_QSCORE + 1]; /* qphred_to_ln_comp_error_prob */
  double lnErr[2][MB200_MAX_QSCORE + 1];  /* qphred_to_ln_error_prob      */
} mb200_qscore_tables;

/* IterativeAssemblerOptions  lib/options/IterativeAssemblerOptions.hpp:26-59
 * (alphabet is fixed to "ACGT"; minContigLength/maxError are never read by the reference) */
typedef struct mb200_asm_opts {
  uint32_t minWordLength;           /* 41 */
  uint32_t maxWordLength;           /* 76 */
  uint32_t wordStepSize;            /* 5  */
  uint32_t minCoverage;             /* 1  */
  uint32_t minConservativeCoverage; /* 2  */
  uint32_t maxAssemblyCount;        /* 10 */
  uint32_t minUnusedReads;          /* 3  */
  uint32_t minSupportReads;         /* 2  */
} mb200_asm_opts;

/* AssembledContig  lib/assembly/AssembledContig.hpp:38-50 ; read sets are bitsets of
 * setWords uint32 words (bit r of word r/32 == read index r, pseudo reads included). */
typedef struct mb200_contig {
  uint64_t seqOff; /* contig bases (ASCII) inside the caller's contig byte pool */
  uint32_t seqLen;
  uint32_t seedReadCount;
  int32_t  conservativeBegin;
  int32_t  conservativeEnd;
  uint64_t supportOff; /* first uint32 word of supportReads inside the caller's set pool */
  uint64_t rejectOff;  /* first uint32 word of rejectReads  */
} mb200_contig;

/* per-locus assembly output (Assembly + the mutated AssemblyReadInput/AssemblyReadOutput) */
typedef struct mb200_asm_locus {
  uint32_t firstContig; /* index into the contig array */
  uint32_t nContigs;    /* <= maxAssemblyCount */
  uint32_t firstPseudo; /* index into the pseudo-read array (reads appended by the reference, IterativeAssembler.cpp:895-905) */
  uint32_t nPseudo;
  uint32_t setWords;    /* uint32 words per read set for this locus */
  uint32_t lastWordLength; /* k of the last buildContigs iteration */
  int32_t  status;
  uint32_t reserved;
} mb200_asm_locus;

typedef struct mb200_config {
  uint32_t ctasPerSm;        /* 0 = library default */
  uint32_t maxContigLen;     /* assembler workspace cap per contig, 0 = default */
  uint64_t scratchBytes;     /* device scratch budget for DP back-pointers, 0 = default */
  uint32_t reserved[4];
} mb200_config;

typedef struct mb200_ctx mb200_ctx;

/* kernel timing of the last call on this ctx (CUDA events on the launching stream) */
typedef struct mb200_timing {
  float    h2dMs;
  float    kernelMs;
  float    d2hMs;
  uint32_t launches; /* kernels launched by the last call */
  uint64_t cells;    /* DP cells filled (aligners) / base comparisons (split) / k-mer windows (assembler) */
  uint64_t algoBytes; /* algorithmic HBM bytes of the dominant kernel */
} mb200_timing;

/* ---- lifecycle --------------------------------------------------------------------------- */
MB200_API int         mb200_create(int device, const mb200_config* cfg, mb200_ctx** out);
MB200_API void        mb200_destroy(mb200_ctx* ctx);
MB200_API const char* mb200_last_error(const mb200_ctx* ctx);
MB200_API const char* mb200_version(void);
MB200_API int         mb200_get_timing(const mb200_ctx* ctx, mb200_timing* out);

/* ---- aligners ------------------------------------------------------------------------------
 * n independent alignments.  cigarPool receives the path segments; *cigarUsed returns the
 * number of uint32 written (MB200_ERR_POOL_OVERFLOW if cigarCap was too small; *cigarUsed then
 * holds the required size). */
MB200_API int mb200_jump_align_batch(mb200_ctx* ctx, const mb200_scores* scores, int32_t jumpScore, uint32_t n,
                                     const uint8_t* pool, uint64_t poolBytes, const mb200_seqview* query,
                                     const mb200_seqview* ref1, const mb200_seqview* ref2,
                                     mb200_jump_result* results, uint32_t* cigarPool, uint64_t cigarCap,
                                     uint64_t* cigarUsed);

MB200_API int mb200_indel_align_batch(mb200_ctx* ctx, const mb200_scores* scores, int32_t largeIndelScore,
                                      uint32_t n, const uint8_t* pool, uint64_t poolBytes,
                                      const mb200_seqview* query, const mb200_seqview* ref,
                                      mb200_aln_result* results, uint32_t* cigarPool, uint64_t cigarCap,
                                      uint64_t* cigarUsed);

MB200_API int mb200_global_align_batch(mb200_ctx* ctx, const mb200_scores* scores, uint32_t n,
                                       const uint8_t* pool, uint64_t poolBytes, const mb200_seqview* query,
                                       const mb200_seqview* ref, mb200_aln_result* results,
                                       uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed);

/* ---- split-read scan ----------------------------------------------------------------------- */
MB200_API int mb200_split_align_batch(mb200_ctx* ctx, const mb200_qscore_tables* tables,
                                      uint32_t flankScoreSize, uint32_t n, const uint8_t* pool,
                                      uint64_t poolBytes, const mb200_sr_job* jobs, mb200_sr_result* results);

/* ---- split-read scoring of a candidate's reads (row a11) --------------------------------------------------------------
 * getReadSplitScore (GSC/SVScorerSplit.cpp:131-267): every read near a breakend is realigned to the four targets of the
 * candidate's SVAlignmentInfo (lib/manta/SVScoreInfo.cpp:31-167, mb200::SVAlignmentInfo in the C++ shim) -- alt contig at
 * bp1 and bp2 with dopt.altQ, reference at bp1 and bp2 with dopt.refQ -- and incrementSplitReadEvidence (:88-128) turns the
 * four SRAlignmentInfo into the read's SVFragmentEvidence entries and the sample's allele counts.  DNA only (the RNA branch
 * :216-221 reads the BAM alignment instead).  A fragment's read is evaluated once per candidate (isSplitEvaluated, :150-153):
 * later occurrences of the same (fragment, isRead1) get status MB200_SPLIT_DUPLICATE and change nothing. */
typedef struct mb200_split_target {
  mb200_seqview seq;              /* bp?ContigSeq() / bp?ReferenceSeq() in the pool */
  int32_t       bpBegin, bpEnd;   /* bp?ContigOffset / bp?RefOffset */
} mb200_split_target;
typedef struct mb200_split_candidate {
  mb200_split_target altBp1, altBp2, refBp1, refBp2;
} mb200_split_candidate;
typedef struct mb200_split_read {
  mb200_seqview seq;      /* read bases as splitReadAligner gets them (already reverse-complemented for a reversed shadow) */
  uint64_t      qualOff;
  uint32_t      fragment; /* index of the read's fragment (qname) inside the candidate */
  uint8_t       isRead1;
  uint8_t       mapq;
  uint8_t       reserved[2];
} mb200_split_read;
#define MB200_SPLIT_DUPLICATE 1
typedef struct mb200_split_read_evidence {
  int32_t status;                      /* 0, MB200_SPLIT_DUPLICATE, or the error status of a splitReadAligner call */
  uint8_t isAnchored, isTier2Anchored; /* setReadEvidence  GSC/SVScorerShared.cpp:26-42 */
  uint8_t countedRef, countedAlt;      /* the read was added to sample.ref / sample.alt (incrementAlleleEvidence :70) */
  float   splitLnLhood[4];             /* [altBp1, altBp2, refBp1, refBp2] SVFragmentEvidenceAlleleBreakendPerRead::splitLnLhood */
  float   splitEvidence[4];
  uint8_t isSplitSupport[4];
  uint8_t isTier2SplitSupport[4];
  float   refEvidence, altEvidence;    /* what was added to the allele's splitReadEvidence */
} mb200_split_read_evidence;
typedef struct mb200_split_allele_counts { /* SVSampleAlleleInfo  lib/manta/SVScoreInfo.hpp:83-123 (split-read part) */
  uint32_t splitReadCount;
  float    splitReadEvidence; /* accumulated in read order, like the reference */
  float    splitReadMapQ;
  uint32_t reserved;
} mb200_split_allele_counts;
typedef struct mb200_split_sample_counts {
  mb200_split_allele_counts ref, alt;
} mb200_split_sample_counts;

/* reads of candidate i: reads[readFirst[i] .. readFirst[i+1]) in the order the reference meets them (bp1's region, then
 * bp2's).  sr[4 * r + k]: the four alignments of read r in call order (altBp1, altBp2, refBp1, refBp2); ev[r]; counts[i]. */
MB200_API int mb200_split_score_batch(mb200_ctx* ctx, const mb200_qscore_tables* tables, uint32_t flankScoreSize,
                                      uint32_t minMapQ, uint32_t minTier2MapQ, uint32_t nCandidates,
                                      const mb200_split_candidate* candidates, const uint32_t* readFirst,
                                      const mb200_split_read* reads, const uint8_t* pool, uint64_t poolBytes,
                                      mb200_sr_result* sr, mb200_split_read_evidence* ev, mb200_split_sample_counts* counts);

/* ---- assembler ------------------------------------------------------------------------------
 * nLoci independent read sets.  reads of locus i are readViews[readOff[i] .. readOff[i+1]).
 * Outputs: loci[nLoci]; contigs[contigCap]; pseudoReads[pseudoCap] (views into contigPool);
 * contigPool (bytes) and setPool (uint32 words) are bump-allocated output pools. */
MB200_API int mb200_assemble_batch(mb200_ctx* ctx, const mb200_asm_opts* opts, uint32_t nLoci,
                                   const uint64_t* readOff, const mb200_seqview* readViews,
                                   const uint8_t* pool, uint64_t poolBytes, mb200_asm_locus* loci,
                                   mb200_contig* contigs, uint64_t contigCap, mb200_seqview* pseudoReads,
                                   uint64_t pseudoCap, uint8_t* contigPool, uint64_t contigPoolCap,
                                   uint32_t* setPool, uint64_t setPoolCap);

/* ---- reference-window trim + contig QC / selection ------------------------------
 * The per-contig steps that surround the large-indel / jump DP in the refiner. */

/* edge cuts of a complex candidate's reference window  GSC/SVCandidateAssemblyRefiner.cpp:1915-1918 */
typedef struct mb200_ref_cuts {
  int32_t leadingCut;     /* 100 in the normal case (extraRefSplitSize) */
  int32_t trailingCut;
  int32_t maxLeadingCut;  /* 800 in the normal case (extraRefSize)      */
  int32_t maxTrailingCut;
} mb200_ref_cuts;

/* adjustedLeadingCut / adjustedTrailingCut of one contig  :1984-2011 */
typedef struct mb200_trim_result {
  int32_t adjLeadingCut;
  int32_t adjTrailingCut;
} mb200_trim_result;

typedef struct mb200_qc_opts {
  mb200_scores filterScores;            /* SVRefinerOptions::contigFilterScores (2,-8,-18,0,-1)  lib/options/SVRefinerOptions.hpp:50 */
  uint32_t     minCandidateVariantSize; /* ReadScannerOptions::minCandidateVariantSize = 10      lib/options/ReadScannerOptions.hpp:68 */
  uint32_t     reserved;
} mb200_qc_opts;

/* per-contig QC outcome.
 * complex candidates (:2046-2066): isCandidate = isSmallSVCandidate; the kept candidateSegments are always a
 *   contiguous run of getLargeIndelSegments(apath, minCandidateVariantSize) (:173): nSegments of them starting at
 *   ordinal segOrdinal, covering apath[pathFirst .. pathLast]; variantSize = getLargestIndelSize (:210).
 * spanning candidates: isCandidate = !isJumpAlignmentQCFail (:1266); isLowQuality = isLowQualityJumpAlignment (:1287,
 *   evaluated for every contig; the reference only looks at it for the top-scoring one). */
typedef struct mb200_contig_qc {
  uint32_t isCandidate;
  uint32_t isLowQuality;
  uint32_t segOrdinal;
  uint32_t nSegments;
  uint32_t pathFirst;
  uint32_t pathLast;
  uint32_t variantSize;
  uint32_t reserved;
} mb200_contig_qc;

/* 10-mer anchored trim of the reference window of n (contig, window) items  :1984-2011.
 * MB200_ERR_BAD_ARG if a contig is shorter than 10 bases (the reference's unsigned loop bound underflows there). */
MB200_API int mb200_ref_trim_batch(mb200_ctx* ctx, uint32_t n, const uint8_t* pool, uint64_t poolBytes,
                                   const mb200_seqview* contig, const mb200_seqview* ref, const mb200_ref_cuts* cuts,
                                   mb200_trim_result* out);

/* contig QC and bestAlignmentIndex selection for nLoci candidates.  Contig c of locus i owns slot slotFirst[i] + c
 * (c < nContigs[i]) of aln (complex, beginPos relative to ref1[i]) / jump (spanning), contigSeq, supportCount
 * (= AssembledContig::supportReads.size()) and qc.  best[i] = SVCandidateAssemblyData::bestAlignmentIndex or -1 when
 * the reference selects nothing (complex: rank1/rank2 rule :2119-2227; spanning: selectJumpContigDNA :1364-1398). */
MB200_API int mb200_select_contigs_batch(mb200_ctx* ctx, const mb200_qc_opts* opts, uint32_t nLoci,
                                         const uint8_t* isJump, const uint32_t* slotFirst, const uint32_t* nContigs,
                                         uint32_t nSlots, const mb200_aln_result* aln, const mb200_jump_result* jump,
                                         const uint32_t* cigarPool, uint64_t cigarLen, const uint8_t* pool,
                                         uint64_t poolBytes, const mb200_seqview* contigSeq, const mb200_seqview* ref1,
                                         const uint32_t* supportCount, mb200_contig_qc* qc, int32_t* best);

/* One refined small SV per kept candidate segment of the selected contig of a complex candidate: the arithmetic of
 * setSmallCandSV (GSC/SVCandidateAssemblyRefiner.cpp:720-800, called per segment at :2243-2251) with getVariantRange
 * (:677-717).  Coordinates are offsets into the candidate's ref1 window (the reference adds ref.get_offset()).
 *   sv.bp1.interval.range = [bp1Begin, bp1End), state RIGHT_OPEN ; sv.bp2.interval.range = [bp2Begin, bp2End), LEFT_OPEN
 *   sv.insertSeq          = contig[readBegin, readEnd)  ; sv.insertAlignment = apath[pathFirst .. pathLast]
 *   sv.assemblyAlignIndex = contig ; sv.assemblySegmentIndex = segment
 * status == MB200_ERR_SV_RANGE where the reference throws (cipos.begin_pos() != 0, :768); the record then still holds
 * the ranges and cipos, the breakend fields are zero. */
typedef struct mb200_small_sv {
  uint32_t locus;
  uint32_t contig;
  uint32_t segment;
  int32_t  status;
  uint32_t pathFirst, pathLast;
  int32_t  refBegin, refEnd;
  int32_t  readBegin, readEnd;
  int32_t  ciposBegin, ciposEnd;
  int32_t  bp1Begin, bp1End, bp2Begin, bp2End;
} mb200_small_sv;

/* Inputs as mb200_select_contigs_batch plus its outputs qc / best.  svFirst[nLoci + 1]: records of candidate i are
 * svs[svFirst[i] .. svFirst[i+1]) in segment order (none for spanning candidates or best[i] < 0).
 * MB200_ERR_POOL_OVERFLOW if svCap is too small (*svUsed = records needed). */
MB200_API int mb200_small_sv_batch(mb200_ctx* ctx, uint32_t minCandidateVariantSize, uint32_t nLoci, const uint8_t* isJump,
                                   const uint32_t* slotFirst, uint32_t nSlots, const int32_t* best,
                                   const mb200_contig_qc* qc, const mb200_aln_result* aln, const uint32_t* cigarPool,
                                   uint64_t cigarLen, const uint8_t* pool, uint64_t poolBytes,
                                   const mb200_seqview* contigSeq, const mb200_seqview* ref1, uint32_t* svFirst,
                                   mb200_small_sv* svs, uint64_t svCap, uint64_t* svUsed);

/* ---- large insertions with an incompletely assembled insert (rest of row a9) ------------------------------------------
 * getSmallSVAssembly tests every contig alignment of a complex candidate as the LEFT or RIGHT side of a large insertion
 * (GSC/SVCandidateAssemblyRefiner.cpp:2079-2112: isLargeInsertAlignment :611 on the alignment limited to the contig's
 * conservative range and on the whole alignment, scored with largeInsertEdgeAlignScores), and processLargeInsertion
 * (:833-1007) pairs a left with a right candidate (all against all, nearest breakpoints first, then best score), joins
 * the two contigs around 100 'N's, aligns the joined contig with GlobalAligner<int>(largeInsertCompleteAlignScores)
 * against the window inside the ORIGINAL leading/trailing cuts, and accepts it if the alignment shows one insertion
 * covering the N block with two well-scoring flanks (getLargestInsertSegment :230, isFinishedLargeInsertAlignment :642,
 * getInsertTrim :801); the accepted pair becomes a refined SV by setSmallCandSV (:720) unless a fully assembled
 * insertion already starts at the same position. */
typedef struct mb200_large_insert_info { /* LargeInsertionInfo  lib/manta/SVCandidateAssemblyData.hpp:60-78 (zeros: no candidate) */
  uint8_t  isLeftCandidate, isRightCandidate, reserved[2];
  uint32_t contigOffset;
  uint32_t refOffset;
  int32_t  score;
} mb200_large_insert_info;

enum { /* mb200_large_insertion.status: where processLargeInsertion left */
  MB200_LI_NONE        = 0, /* no left/right pair within maxBreakDist (:916) */
  MB200_LI_SEGMENT_QC  = 1, /* joined contig aligned, but not exactly one insert segment >= 100 (:950) */
  MB200_LI_FLANK_QC    = 2, /* isFinishedLargeInsertAlignment failed (:955) */
  MB200_LI_TRIM_QC     = 3, /* partial insert sequences shorter than 40 (:966-973) */
  MB200_LI_RANGE_ERROR = 4, /* setSmallCandSV throws here: cipos does not start at 0 (:768) */
  MB200_LI_DUPLICATE   = 5, /* a fully assembled insertion of the selected contig starts at the same position (:991) */
  MB200_LI_ACCEPTED    = 6
};
typedef struct mb200_large_insertion {
  int32_t  status;
  uint32_t leftContig, rightContig; /* contig indices inside the candidate (status >= 1) */
  int32_t  breakDist, breakScore;
  int32_t  score, beginPos;         /* alignment of left + 100 N + right; beginPos in the UNcut window (:941) */
  uint32_t cigarLen;
  uint64_t cigarOff;                /* into the call's output cigar pool */
  uint32_t segFirst, segLast;       /* the insert segment (status >= 2) */
  int32_t  insertTrimBegin, insertTrimEnd; /* getInsertTrim (status >= 3) */
  int32_t  refBegin, refEnd, readBegin, readEnd, ciposBegin, ciposEnd; /* as mb200_small_sv (status >= 4) */
  int32_t  bp1Begin, bp1End, bp2Begin, bp2End;                         /* status >= 5 */
  uint32_t leftSeqBegin;  /* unknownSizeInsertionLeftSeq  = leftContig[leftSeqBegin, end)   (status == ACCEPTED) */
  uint32_t rightSeqLen;   /* unknownSizeInsertionRightSeq = rightContig[0, rightSeqLen) */
} mb200_large_insertion;

/* Layout as mb200_select_contigs_batch.  aln[] are the large-indel alignments of the contigs (beginPos in the uncut window),
 * consRange[2*slot .. ] = AssembledContig::conservativeRange (begin, end), cuts may be NULL (no cut), svFirst/svs = the
 * refined small SVs of the selected contigs (mb200_small_sv_batch; NULL: nothing is excluded).  info[nSlots] and
 * out[nLoci] are written for complex candidates (zeros for spanning ones).  MB200_ERR_POOL_OVERFLOW when outCigarCap is
 * too small (*outCigarUsed = segments needed). */
MB200_API int mb200_large_insertion_batch(mb200_ctx* ctx, const mb200_scores* edgeScores, const mb200_scores* completeScores,
                                          uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst,
                                          const uint32_t* nContigs, uint32_t nSlots, const mb200_aln_result* aln,
                                          const uint32_t* cigarPool, uint64_t cigarLen, const uint8_t* pool, uint64_t poolBytes,
                                          const mb200_seqview* contigSeq, const int32_t* consRange, const mb200_seqview* ref1,
                                          const mb200_ref_cuts* cuts, const uint32_t* svFirst, const mb200_small_sv* svs,
                                          mb200_large_insert_info* info, mb200_large_insertion* out, uint32_t* outCigar,
                                          uint64_t outCigarCap, uint64_t* outCigarUsed);

/* Refined breakends of a spanning candidate from its selected jump alignment: generateRefinedSVCandidateFromJumpAlignment
 * (GSC/SVCandidateAssemblyRefiner.cpp:1190-1223) = adjustAssembledBreakend (:59-86) on both breakends.  What the
 * refiner knows about the candidate besides the alignment: */
typedef struct mb200_jump_orient {
  uint8_t isBp2AlignedFirst; /* SVCandidateAssemblyData::bporient  lib/manta/SVCandidateAssemblyData.hpp:39-58 */
  uint8_t isBp1Reversed;
  uint8_t isBp2Reversed;
  uint8_t bp1RightOpen;      /* sv.bp1.state == SVBreakendState::RIGHT_OPEN (else the breakend sits at the alignment begin) */
  uint8_t bp2RightOpen;
  uint8_t reserved[3];
} mb200_jump_orient;

typedef struct mb200_jump_sv {
  int32_t  valid;       /* 1 for a spanning candidate with a selected contig, else 0 (rest zero) */
  uint32_t contig;      /* sv.assemblyAlignIndex */
  int32_t  bp1Begin, bp1End; /* sv.bp1.interval.range as offsets into bp1's reference window (add bp1ref.get_offset()) */
  int32_t  bp2Begin, bp2End;
  int32_t  insertBegin, insertEnd; /* sv.insertSeq = contig[insertBegin, insertEnd), reverse-complemented if isBp1Reversed
                                      (getFwdStrandInsertSegment, lib/alignment/AlignmentUtil.cpp) */
} mb200_jump_sv;

/* ref1 / ref2 are the windows the contigs were aligned to as align1 / align2 (bp1's window is ref2 when
 * isBp2AlignedFirst).  out[nLoci]. */
MB200_API int mb200_jump_sv_batch(mb200_ctx* ctx, uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst,
                                  uint32_t nSlots, const int32_t* best, const mb200_jump_result* jump,
                                  const uint32_t* cigarPool, uint64_t cigarLen, const mb200_seqview* ref1,
                                  const mb200_seqview* ref2, const mb200_jump_orient* orient, mb200_jump_sv* out);

/* ---- fused candidate path -------------------------------------------------------------------------
 * "one candidate" = one SVCandidate junction entering getCandidateAssemblyData (GSC/SVCandidateProcessor.cpp:310):
 * its read set is assembled (all word lengths) and EVERY resulting contig is aligned
 *   - complex candidates (isJump == 0): GlobalLargeIndelAligner against ref1      (SVCandidateAssemblyRefiner.cpp:2032)
 *   - spanning candidates (isJump == 1): GlobalJumpAligner against (ref1, ref2)   (SVCandidateAssemblyRefiner.cpp:1668)
 * with contigs staying resident on the device between the two stages.
 * Result slot of contig c of locus i is i*maxAssemblyCount + c in alnResults (complex) / jumpResults (spanning).
 * Optional stages, each to the reference's semantics:
 *   - cuts != NULL at upload: every complex contig's ref1 window is first trimmed to its outermost shared 10-mers
 *     (:1984-2011) and beginPos is reported relative to the untrimmed ref1 (:2038); trim[slot] returns the cuts used
 *   - qc / best != NULL at run: contig QC and bestAlignmentIndex selection (mb200_select_contigs_batch) run on the
 *     device right after the DP, on the still-resident alignments. */
typedef struct mb200_candidate_opts {
  mb200_asm_opts asmOpts;
  mb200_scores   indelScores;     /* SVRefinerOptions::smallSVAlignScores (2,-8,-24,-1,-1)  lib/options/SVRefinerOptions.hpp:40 */
  int32_t        largeIndelScore; /* largeGapOpenScore -100                                 :44 */
  mb200_scores   jumpScores;      /* spanningAlignScores (2,-8,-12,-1,-1)                   :43 */
  int32_t        jumpScore;       /* -100                                                   :45 */
  mb200_qc_opts  qcOpts;          /* contig QC / selection (used when qc/best outputs are requested) */
} mb200_candidate_opts;

typedef struct mb200_candidate_timing {
  float    h2dMs, asmMs, indelMs, jumpMs, d2hMs, totalMs; /* CUDA events on the ctx stream */
  uint32_t launches, nContigs, nIndelTasks, nJumpTasks;
  uint64_t asmWindows, indelCells, jumpCells;
  uint64_t asmBytes, indelBytes, jumpBytes; /* algorithmic HBM bytes per stage */
  uint64_t h2dBytes, d2hBytes;
  float    trimMs, qcMs;                   /* reference-window trim / contig QC + selection kernels */
} mb200_candidate_timing;

/* Stage one batch of candidate loci in device memory (inputs resident in HBM afterwards).  The copy is asynchronous
 * (its own stream): the caller's buffers must stay untouched until the mb200_candidates_run that consumes the batch
 * has returned.  Up to two batches can be staged; runs consume them first-in first-out, so
 *     upload(b0); for (i...) { upload(b[i+1]); run(); }      copies batch i+1 while batch i is processed.
 * With nothing staged, mb200_candidates_run repeats on the batch it consumed last (resident inputs). */
MB200_API int mb200_candidates_upload(mb200_ctx* ctx, uint32_t nLoci, const uint64_t* readOff,
                                      const mb200_seqview* readViews, const uint8_t* isJump,
                                      const mb200_seqview* ref1, const mb200_seqview* ref2, const uint8_t* pool,
                                      uint64_t poolBytes, const mb200_ref_cuts* cuts);
/* assemble + align every contig of the staged batch; all outputs are host buffers */
MB200_API int mb200_candidates_run(mb200_ctx* ctx, const mb200_candidate_opts* opts, mb200_asm_locus* loci,
                                   mb200_contig* contigs, mb200_seqview* pseudoReads, uint8_t* contigPool,
                                   uint64_t contigPoolCap, uint64_t* contigPoolUsed, uint32_t* setPool,
                                   uint64_t setPoolCap, uint64_t* setPoolUsed, mb200_aln_result* alnResults,
                                   mb200_jump_result* jumpResults, uint32_t* cigarPool, uint64_t cigarCap,
                                   uint64_t* cigarUsed, mb200_trim_result* trim, mb200_contig_qc* qc, int32_t* best);
MB200_API int mb200_candidates_timing(const mb200_ctx* ctx, mb200_candidate_timing* out);
/* shape of the batch the NEXT mb200_candidates_run will work on (the oldest staged one, else the one consumed last):
 * the record arrays of that run must hold nLoci (loci, best), nLoci*maxAssemblyCount (contigs, alnResults,
 * jumpResults, trim, qc) and 2*nLoci*maxAssemblyCount (pseudoReads) entries.  With uploads pipelined ahead of runs
 * the newest upload is NOT the batch the next run consumes.  MB200_ERR_BAD_ARG when nothing is staged. */
MB200_API int mb200_candidates_pending(mb200_ctx* ctx, uint32_t* nLoci, uint64_t* nReads, uint64_t* poolBytes);
/* refined small SVs (mb200_small_sv_batch) of the batch the last mb200_candidates_run with qc/best worked on, computed
 * on the still-resident alignments */
MB200_API int mb200_candidates_small_svs(mb200_ctx* ctx, uint32_t* svFirst, mb200_small_sv* svs, uint64_t svCap,
                                         uint64_t* svUsed);
/* refined spanning breakends (mb200_jump_sv_batch) of the same run; orient[nLoci] (entries of complex candidates are
 * ignored), out[nLoci] */
MB200_API int mb200_candidates_jump_svs(mb200_ctx* ctx, const mb200_jump_orient* orient, mb200_jump_sv* out);

/* ---- row (e): the one collective of the path -- merge the per-rank result buffers on `root` over NCCL ------------------
 * Loci shard across the GPUs of a box with no data-path collective; this gather of the per-rank candidate /
 * result buffers is the only exchange.  `ncclComm` is an ncclComm_t (the host's own, or one made with the two helpers
 * below when the host has no NCCL bootstrap of its own: rank 0 calls mb200_nccl_unique_id and hands the 128 bytes to the
 * other ranks by any means, then every rank calls mb200_nccl_comm_init).  localDev / rootDev are DEVICE pointers;
 * rankBytes[nRanks] receives every rank's byte count on every rank; on `root` the merged buffer is the concatenation in
 * rank order.  Collective: every rank of the communicator must call it; it returns after the transfer has completed on
 * the ctx stream.  MB200_ERR_POOL_OVERFLOW (on every rank alike) when the merged size exceeds rootCap.  NCCL is loaded on
 * first use (dlopen libnccl.so.2): MB200_ERR_UNSUPPORTED when it is not installed. */
MB200_API int mb200_gather_results(mb200_ctx* ctx, void* ncclComm, int root, const void* localDev, uint64_t localBytes,
                                   void* rootDev, uint64_t rootCap, uint64_t* rankBytes);
MB200_API int mb200_nccl_unique_id(uint8_t id[128]);
MB200_API int mb200_nccl_comm_init(mb200_ctx* ctx, int nRanks, const uint8_t id[128], int rank, void** ncclComm);
MB200_API int mb200_nccl_comm_destroy(void* ncclComm);

/* ---- row a1: the per-bin edge partition (host arithmetic, no device, no ctx) ---------------------------------------
 * GSC/EdgeRetrieverBin.cpp:43-235 on a flattened SVLocusSet (include/manta_b200_graph.hpp: mb200::FlatLocusGraph, filled on
 * the reference side by mb200::flattenLocusSet): locusNodeOff[nLoci+1], nodeEdgeOff[nNodes+1], and per out-edge its target
 * node index inside the locus (ascending per node, the std::map order) and its directed observation count.  Writes the
 * edges of bin `binIndex` of `binCount` in the retriever's order; *nEdges is the number the bin holds (MB200_ERR_POOL_OVERFLOW
 * when that exceeds edgeCap: nothing beyond edgeCap is written).  graphNodeMaxEdgeCount 0 = no noisy-node filter. */
typedef struct mb200_edge_info {
  uint32_t locusIndex, nodeIndex1, nodeIndex2; /* lib/svgraph/EdgeInfo.hpp:30-39 */
} mb200_edge_info;
MB200_API int mb200_edge_bin_enumerate(uint32_t nLoci, const uint32_t* locusNodeOff, const uint64_t* nodeEdgeOff,
                                       const uint32_t* edgeTo, const uint32_t* edgeCount, uint32_t graphNodeMaxEdgeCount,
                                       uint32_t binCount, uint32_t binIndex, mb200_edge_info* edges, uint64_t edgeCap,
                                       uint64_t* nEdges);

#ifdef __cplusplus
}
#endif
#endif /* MANTA_B200_H */
