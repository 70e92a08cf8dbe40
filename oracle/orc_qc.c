/*
 * orc_qc.c -- TEST INFRASTRUCTURE ONLY (see manta_oracle.h).
 *
 * CPU restatement of the per-contig steps that surround the refiner's DP:
 *   - the 10-mer anchored trim of a complex candidate's reference window
 *     (GSC/SVCandidateAssemblyRefiner.cpp:1984-2011)
 *   - contig QC (isLowQualitySmallSVAlignment :318, findCandidateVariantsFromComplexSVContigAlignment :430,
 *     isJumpAlignmentQCFail :1266, isLowQualityJumpAlignment :1287 -> isLowQualitySpanningSVAlignment :93)
 *   - bestAlignmentIndex selection (rank1/rank2 rule :2119-2227, selectJumpContigDNA :1364-1398)
 * on the C ABI's flat structures.  Pinned against the reference's own static functions through
 * oracle/_ref (ref_select_contigs_batch / ref_ref_trim_batch, tests/test_oracle_vs_ref.py).
 */
#include <stdlib.h>
#include <string.h>

#include "manta_oracle.h"

/* ---------------------------------------------------------------------------------------------------------------
 * alignment-path helpers on packed segments (lib/blt_util/align_path.{hpp,cpp})
 * ------------------------------------------------------------------------------------------------------------- */

static int seg_is_read(uint32_t t) /* is_segment_type_read_length  align_path.hpp:89 */
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH ||
         t == MB200_SEG_SEQ_MISMATCH;
}

static int seg_is_ref(uint32_t t) /* is_segment_type_ref_length  align_path.hpp:103 */
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH ||
         t == MB200_SEG_SEQ_MISMATCH;
}

static uint32_t path_read_length(const uint32_t* s, uint32_t n) /* apath_read_length  align_path.cpp:97 */
{
  uint32_t v = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (seg_is_read(MB200_SEG_TYPE(s[i]))) v += MB200_SEG_LEN(s[i]);
  return v;
}

static uint32_t path_ref_length(const uint32_t* s, uint32_t n) /* apath_ref_length  align_path.cpp:126 */
{
  uint32_t v = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (seg_is_ref(MB200_SEG_TYPE(s[i]))) v += MB200_SEG_LEN(s[i]);
  return v;
}

/* One flank ("arm") of a breakpoint: the sub-path s[0..n), read away from the breakpoint (reversed when it leads
 * the breakpoint), cut after the first maxSpan reference bases (apath_limit_ref_length, align_path.cpp:273), then
 * measured: reference length, read length, trailing soft clip (apath_soft_clip_right_size :184) and
 * getPathScore(scores, apath, isScoreOffEdge=false) (lib/alignment/AlignmentScoringUtilImpl.hpp:35-92; the open
 * score is charged for every insert/delete segment, adjacent ones included). */
typedef struct {
  uint32_t refSize, readSize, clipRight;
  int32_t  score;
} arm_t;

static arm_t arm_measure(const uint32_t* s, uint32_t n, int reversed, uint32_t maxSpan, const mb200_scores* sc)
{
  arm_t a;
  memset(&a, 0, sizeof(a));
  uint32_t refLen = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t seg = reversed ? s[n - 1 - i] : s[i];
    const uint32_t t   = MB200_SEG_TYPE(seg);
    uint32_t       len = MB200_SEG_LEN(seg);
    int            stop = 0;
    if (seg_is_ref(t)) {
      refLen += len;
      if (refLen >= maxSpan) {
        len -= (refLen - maxSpan);
        stop = 1;
      }
      a.refSize += len;
    }
    if (seg_is_read(t)) a.readSize += len;
    if (t == MB200_SEG_SOFT_CLIP) a.clipRight += len;
    else if (t != MB200_SEG_HARD_CLIP) a.clipRight = 0;
    switch (t) {
    case MB200_SEG_SEQ_MATCH: a.score += sc->match * (int32_t)len; break;
    case MB200_SEG_SEQ_MISMATCH: a.score += sc->mismatch * (int32_t)len; break;
    case MB200_SEG_INSERT:
    case MB200_SEG_DELETE: a.score += sc->open + sc->extend * (int32_t)len; break;
    default: break;
    }
    if (stop) break;
  }
  return a;
}

static int low_score_fraction(const arm_t* a, uint32_t clipped, const mb200_scores* sc)
{
  const int32_t nonClip = a->score > 0 ? a->score : 0;
  const int32_t optimal = (int32_t)clipped * sc->match;
  const float   frac    = (float)nonClip / (float)optimal;
  return frac < 0.75f;
}

/* isLowQualitySmallSVAlignment  :318-388 */
static int low_quality_small(const uint32_t* s, uint32_t n, int leading, int isComplex, uint32_t maxSpan,
                             const mb200_scores* sc)
{
  const uint32_t minRef  = isComplex ? 35 : 30;
  const uint32_t minRead = isComplex ? 35 : 30;
  const arm_t    a       = arm_measure(s, n, leading, maxSpan, sc);
  if (a.refSize < minRef) return 1;
  const uint32_t clipped = a.readSize - a.clipRight;
  if (clipped < minRead) return 1;
  return low_score_fraction(&a, clipped, sc);
}

/* isLowQualitySpanningSVAlignment  :93-163 (isRNA == false) */
static int low_quality_spanning(const uint32_t* s, uint32_t n, int leading, uint32_t maxSpan, const mb200_scores* sc)
{
  const arm_t    a       = arm_measure(s, n, leading, maxSpan, sc);
  const uint32_t clipped = a.readSize - a.clipRight;
  if (clipped < 30) return 1;
  return low_score_fraction(&a, clipped, sc);
}

/* getQuerySeqMatchCount  :393-418 */
static uint32_t match_count(const uint8_t* target, uint32_t targetSize, const uint8_t* query, uint32_t querySize)
{
  uint32_t count = 0;
  if (querySize > targetSize) return 0;
  const uint32_t scanEnd = targetSize - querySize;
  for (uint32_t i = 0; i <= scanEnd; ++i) {
    uint32_t mm = 0;
    for (uint32_t j = 0; j < querySize; ++j)
      if (query[j] != target[i + j] || query[j] == 'N') mm++;
    if ((float)mm / (float)querySize <= 0.05f) count++;
  }
  return count;
}

typedef struct {
  uint32_t first, last;
} seg_t;

/* getLargeIndelSegments  :173-208 ; returns the segment count, writes at most cap of them */
static uint32_t large_indel_segments(const uint32_t* s, uint32_t n, uint32_t minSize, seg_t* out, uint32_t cap)
{
  uint32_t cnt = 0, start = 0;
  int      inSeg = 0, cand = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t t = MB200_SEG_TYPE(s[i]);
    if (t == MB200_SEG_DELETE || t == MB200_SEG_INSERT) {
      if (MB200_SEG_LEN(s[i]) >= minSize) cand = 1;
      if (!inSeg) start = i;
      inSeg = 1;
    } else {
      if (cand) {
        if (cnt < cap) { out[cnt].first = start; out[cnt].last = i - 1; }
        cnt++;
      }
      inSeg = 0;
      cand  = 0;
    }
  }
  if (cand) {
    if (cnt < cap) { out[cnt].first = start; out[cnt].last = n - 1; }
    cnt++;
  }
  return cnt;
}

/* findCandidateVariantsFromComplexSVContigAlignment  :430-553 for one maxQCRefSpan.
 * The surviving candidateSegments are segs[*a .. *b); the final ">= minSize" filter (:538-550) keeps every segment
 * (each one was opened by an indel of that size). */
static int find_candidates(const uint32_t* path, uint32_t n, int32_t beginPos, const uint8_t* contig, uint32_t contigLen,
                           const uint8_t* ref, uint32_t refLen, const seg_t* segs, uint32_t nSegs, uint32_t maxSpan,
                           const mb200_scores* sc, uint32_t* a, uint32_t* b)
{
  if (nSegs == 0) return 0;
  const int isComplex = (nSegs > 1) || (segs[0].first != segs[0].last);
  uint32_t  lo = 0, hi = nSegs;
  for (;;) {
    if (!low_quality_small(path, segs[lo].first, 1, isComplex, maxSpan, sc)) break;
    if (hi - lo == 1) return 0;
    lo++;
  }
  for (;;) {
    const uint32_t from = segs[hi - 1].last + 1;
    if (!low_quality_small(path + from, n - from, 0, isComplex, maxSpan, sc)) break;
    if (hi - lo == 1) return 0;
    hi--;
  }
  {
    const int32_t leftSize  = (int32_t)path_read_length(path, segs[lo].first);
    const int32_t endPos    = (int32_t)path_read_length(path, segs[hi - 1].last + 1);
    int32_t       rightSize = (int32_t)contigLen - endPos;
    /* std::string::substr clamps the count to the end of the string */
    uint32_t leftQ = (uint32_t)leftSize > contigLen ? contigLen : (uint32_t)leftSize;
    if (rightSize < 0) rightSize = 0;
    const int32_t refAlignStart = beginPos;
    const int32_t refAlignEnd   = beginPos + (int32_t)path_ref_length(path, n);
    const int32_t leftStart     = (refAlignEnd - 500) > 0 ? (refAlignEnd - 500) : 0;
    uint32_t      tl            = (uint32_t)(refAlignEnd - leftStart);
    if ((uint32_t)leftStart + tl > refLen) tl = refLen - (uint32_t)leftStart;
    if (match_count(ref + leftStart, tl, contig, leftQ) > 1) return 0;
    int32_t rs = (int32_t)refLen - refAlignStart;
    if (rs > 500) rs = 500;
    if (rs < 0) rs = 0;
    if (match_count(ref + refAlignStart, (uint32_t)rs, contig + endPos, (uint32_t)rightSize) > 1) return 0;
  }
  *a = lo;
  *b = hi;
  return 1;
}

#define ORC_MAX_SEGS 4096

static void qc_complex_contig(const mb200_qc_opts* o, const mb200_aln_result* r, const uint32_t* cigarPool,
                              const uint8_t* contig, uint32_t contigLen, const uint8_t* ref, uint32_t refLen,
                              mb200_contig_qc* q)
{
  static const uint32_t spanSet[2] = {100, 200};
  memset(q, 0, sizeof(*q));
  const uint32_t* path = cigarPool + r->cigarOff;
  const uint32_t  n    = r->cigarLen;
  seg_t*          segs = (seg_t*)malloc(sizeof(seg_t) * (n ? n : 1));
  const uint32_t  nSegs = large_indel_segments(path, n, o->minCandidateVariantSize, segs, n ? n : 1);
  uint32_t        bestA = 0, bestB = 0;
  int             isCand = 0;
  for (int k = 0; k < 2; ++k) {
    uint32_t a = 0, b = 0;
    if (find_candidates(path, n, r->beginPos, contig, contigLen, ref, refLen, segs, nSegs, spanSet[k], &o->filterScores, &a, &b)) {
      if (b - a > bestB - bestA) { bestA = a; bestB = b; }
      isCand = 1;
    }
  }
  if (isCand) {
    q->isCandidate = 1;
    q->segOrdinal  = bestA;
    q->nSegments   = bestB - bestA;
    q->pathFirst   = segs[bestA].first;
    q->pathLast    = segs[bestB - 1].last;
    uint32_t big   = 0; /* getLargestIndelSize  :210-226 (indels between two kept segments belong to no segment) */
    for (uint32_t sidx = bestA; sidx < bestB; ++sidx)
      for (uint32_t i = segs[sidx].first; i <= segs[sidx].last; ++i) {
        const uint32_t t = MB200_SEG_TYPE(path[i]);
        if ((t == MB200_SEG_DELETE || t == MB200_SEG_INSERT) && MB200_SEG_LEN(path[i]) > big) big = MB200_SEG_LEN(path[i]);
      }
    q->variantSize = big;
  }
  free(segs);
}

static void qc_spanning_contig(const mb200_qc_opts* o, const mb200_jump_result* r, const uint32_t* cigarPool,
                               mb200_contig_qc* q)
{
  static const uint32_t spanSet[3] = {75, 100, 200};
  memset(q, 0, sizeof(*q));
  const uint32_t* p1 = cigarPool + r->cigarOff1;
  const uint32_t* p2 = cigarPool + r->cigarOff2;
  /* isJumpAlignmentSegmentQCFail :1254 (isAligned() == !apath.empty(), lib/alignment/Alignment.hpp) */
  const int fail = (r->cigarLen1 == 0) || (path_ref_length(p1, r->cigarLen1) < 20) || (r->cigarLen2 == 0) ||
                   (path_ref_length(p2, r->cigarLen2) < 20);
  q->isCandidate = !fail;
  int low1 = 1, low2 = 1;
  for (int k = 0; k < 3; ++k) {
    if (!low_quality_spanning(p1, r->cigarLen1, 1, spanSet[k], &o->filterScores)) low1 = 0;
    if (!low_quality_spanning(p2, r->cigarLen2, 0, spanSet[k], &o->filterScores)) low2 = 0;
  }
  q->isLowQuality = (low1 || low2);
}

/* the rank1/rank2 rule of getSmallSVAssembly  :1955-1957, 2119-2227 */
static int32_t select_complex(uint32_t nContigs, const mb200_aln_result* aln, const mb200_contig_qc* qc,
                              const uint32_t* support)
{
  int      def1 = 0, def2 = 0;
  uint32_t idx1 = 0, idx2 = 0, var1 = 0, var2 = 0;
  int32_t  sc1 = 0, sc2 = 0;
  int      jmp1 = 0, jmp2 = 0;
  for (uint32_t c = 0; c < nContigs; ++c) {
    if (!qc[c].isCandidate) continue;
    const int     jumped = aln[c].isJumped != 0;
    const int32_t score  = aln[c].score;
    const int     bothJ  = jumped && jmp1;
    const int     bothN  = (!jumped) && (!jmp1);
    if ((!def1) || (jumped && !jmp1) || ((bothJ || bothN) && (score > sc1))) {
      if (def1) { def2 = 1; idx2 = idx1; sc2 = sc1; var2 = var1; jmp2 = jmp1; }
      def1 = 1; idx1 = c; sc1 = score; var1 = qc[c].variantSize; jmp1 = jumped;
    } else if ((!def2) || (score > sc2)) {
      def2 = 1; idx2 = c; sc2 = score; var2 = qc[c].variantSize; jmp2 = jumped;
    }
  }
  if (def2) {
    const uint32_t n1 = support[idx1], n2 = support[idx2];
    const int      rank1Sel = jmp1 && !jmp2;
    if (!rank1Sel) {
      const int rank2Best = ((float)sc2 > ((float)sc1 * 0.9f)) &&
                            (((float)n2 > ((float)n1 * 1.2f)) || ((float)var2 > ((float)var1 * 1.1f)));
      if (rank2Best) idx1 = idx2;
    }
  }
  return def1 ? (int32_t)idx1 : -1;
}

/* selectJumpContigDNA  :1364-1398 */
static int32_t select_spanning(uint32_t nContigs, const mb200_jump_result* jump, const mb200_contig_qc* qc)
{
  int32_t best = -1;
  for (uint32_t c = 0; c < nContigs; ++c) {
    if (!qc[c].isCandidate) continue;
    if (best == -1 || jump[c].score > jump[best].score) best = (int32_t)c;
  }
  if (best == -1 || qc[best].isLowQuality) return -1;
  return best;
}

int orc_select_contigs_batch(const mb200_qc_opts* opts, uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst,
                             const uint32_t* nContigs, uint32_t nSlots, const mb200_aln_result* aln,
                             const mb200_jump_result* jump, const uint32_t* cigarPool, const uint8_t* pool,
                             const mb200_seqview* contigSeq, const mb200_seqview* ref1, const uint32_t* supportCount,
                             mb200_contig_qc* qc, int32_t* best)
{
  (void)nSlots;
  for (uint32_t li = 0; li < nLoci; ++li) {
    const uint32_t s0 = slotFirst[li];
    for (uint32_t c = 0; c < nContigs[li]; ++c) {
      const uint32_t s = s0 + c;
      if (isJump[li]) qc_spanning_contig(opts, &jump[s], cigarPool, &qc[s]);
      else
        qc_complex_contig(opts, &aln[s], cigarPool, pool + contigSeq[s].off, contigSeq[s].len, pool + ref1[li].off,
                          ref1[li].len, &qc[s]);
    }
    best[li] = isJump[li] ? select_spanning(nContigs[li], jump + s0, qc + s0)
                          : select_complex(nContigs[li], aln + s0, qc + s0, supportCount + s0);
  }
  return MB200_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * refined small SVs: setSmallCandSV :720-800 over the kept segments of the selected contig (:2243-2251)
 * ------------------------------------------------------------------------------------------------------------- */

/* getVariantRange  :677-717 */
static void variant_range(const uint8_t* ref, int64_t refLen, int32_t refB, int32_t refE, const uint8_t* read, int64_t readLen,
                          int32_t readB, int32_t readE, int32_t* left, int32_t* right)
{
  int64_t maxRight = refLen - refE;
  if (readLen - readE < maxRight) maxRight = readLen - readE;
  int32_t r = 0;
  for (; r < maxRight; ++r)
    if (ref[refB + r] != read[readB + r]) break;
  const int32_t minLeft = (-refB > -readB) ? -refB : -readB;
  int32_t       l       = 0;
  for (; l >= minLeft; --l) {
    /* (the reference indexes position -1 here when an empty range slides to the very start: undefined there,
     *  a stop here) */
    if (refE + l - 1 < 0 || readE + l - 1 < 0) break;
    if (ref[refE + l - 1] != read[readE + l - 1]) break;
  }
  *left  = l;
  *right = r;
}

int orc_small_sv_batch(uint32_t minCandidateVariantSize, uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst,
                       uint32_t nSlots, const int32_t* best, const mb200_contig_qc* qc, const mb200_aln_result* aln,
                       const uint32_t* cigarPool, const uint8_t* pool, const mb200_seqview* contigSeq,
                       const mb200_seqview* ref1, uint32_t* svFirst, mb200_small_sv* svs, uint64_t svCap, uint64_t* svUsed)
{
  (void)nSlots;
  uint64_t used = 0;
  for (uint32_t li = 0; li < nLoci; ++li) {
    svFirst[li] = (uint32_t)used;
    if (isJump[li] || best[li] < 0) continue;
    const uint32_t          s    = slotFirst[li] + (uint32_t)best[li];
    const mb200_contig_qc*  q    = &qc[s];
    const mb200_aln_result* r    = &aln[s];
    const uint32_t*         path = cigarPool + r->cigarOff;
    const uint32_t          n    = r->cigarLen;
    seg_t*                  segs = (seg_t*)malloc(sizeof(seg_t) * (n ? n : 1));
    large_indel_segments(path, n, minCandidateVariantSize, segs, n ? n : 1);
    for (uint32_t k = 0; k < q->nSegments; ++k, ++used) {
      if (used >= svCap) continue;
      const seg_t    sg = segs[q->segOrdinal + k];
      mb200_small_sv v;
      memset(&v, 0, sizeof(v));
      v.locus = li; v.contig = (uint32_t)best[li]; v.segment = k; v.pathFirst = sg.first; v.pathLast = sg.last;
      int32_t readPos = 0, refPos = r->beginPos;
      for (uint32_t i = 0; i < n; ++i) {
        if (i == sg.first) { v.refBegin = refPos; v.readBegin = readPos; }
        if (seg_is_ref(MB200_SEG_TYPE(path[i]))) refPos += (int32_t)MB200_SEG_LEN(path[i]);
        if (seg_is_read(MB200_SEG_TYPE(path[i]))) readPos += (int32_t)MB200_SEG_LEN(path[i]);
        if (i == sg.last) { v.refEnd = refPos; v.readEnd = readPos; }
      }
      variant_range(pool + ref1[li].off, ref1[li].len, v.refBegin, v.refEnd, pool + contigSeq[s].off, contigSeq[s].len,
                    v.readBegin, v.readEnd, &v.ciposBegin, &v.ciposEnd);
      if (v.ciposBegin != 0) v.status = MB200_ERR_SV_RANGE;
      else {
        v.bp1Begin = v.refBegin - 1;
        v.bp1End   = v.bp1Begin + v.ciposEnd + 1;
        v.bp2Begin = v.refEnd;
        v.bp2End   = v.bp2Begin + v.ciposEnd + 1;
      }
      svs[used] = v;
    }
    free(segs);
  }
  svFirst[nLoci] = (uint32_t)used;
  *svUsed        = used;
  return used > svCap ? MB200_ERR_POOL_OVERFLOW : MB200_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * refined spanning breakends: generateRefinedSVCandidateFromJumpAlignment :1190-1223, adjustAssembledBreakend :59-86
 * ------------------------------------------------------------------------------------------------------------- */
static void adjust_breakend(int32_t beginPos, const uint32_t* path, uint32_t n, int isAlign1, uint32_t jumpRange,
                            uint32_t refSize, int isReversed, int rightOpen, int32_t* rb, int32_t* re)
{
  const int32_t alignEnd = beginPos + (int32_t)path_ref_length(path, n); /* alignEnd()  lib/alignment/Alignment.hpp */
  /* getAlignBeginOffset / getAlignEndOffset  lib/alignment/AlignmentUtil.hpp:38-56 */
  const int32_t beginOff = isReversed ? ((int32_t)refSize - alignEnd) : beginPos;
  const int32_t endOff   = isReversed ? ((int32_t)refSize - beginPos) : alignEnd;
  const int32_t bpOff    = rightOpen ? (endOff - 1) : beginOff;
  const int     leftAligned = (isAlign1 != 0) == (rightOpen != 0);
  if (leftAligned) { *rb = bpOff; *re = bpOff + (int32_t)jumpRange + 1; }
  else { *rb = bpOff - (int32_t)jumpRange; *re = bpOff + 1; }
}

int orc_jump_sv_batch(uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst, uint32_t nSlots, const int32_t* best,
                      const mb200_jump_result* jump, const uint32_t* cigarPool, const mb200_seqview* ref1,
                      const mb200_seqview* ref2, const mb200_jump_orient* orient, mb200_jump_sv* out)
{
  (void)nSlots;
  for (uint32_t li = 0; li < nLoci; ++li) {
    mb200_jump_sv v;
    memset(&v, 0, sizeof(v));
    if (isJump[li] && best[li] >= 0) {
      const mb200_jump_result* r = &jump[slotFirst[li] + (uint32_t)best[li]];
      const mb200_jump_orient* o = &orient[li];
      const uint32_t*          p1 = cigarPool + r->cigarOff1;
      const uint32_t*          p2 = cigarPool + r->cigarOff2;
      v.valid  = 1;
      v.contig = (uint32_t)best[li];
      /* bp1 takes align2 (and ref2's window) when bp2 was aligned first */
      if (!o->isBp2AlignedFirst) {
        adjust_breakend(r->beginPos1, p1, r->cigarLen1, 1, r->jumpRange, ref1[li].len, o->isBp1Reversed, o->bp1RightOpen, &v.bp1Begin, &v.bp1End);
        adjust_breakend(r->beginPos2, p2, r->cigarLen2, 0, r->jumpRange, ref2[li].len, o->isBp2Reversed, o->bp2RightOpen, &v.bp2Begin, &v.bp2End);
      } else {
        adjust_breakend(r->beginPos2, p2, r->cigarLen2, 0, r->jumpRange, ref2[li].len, o->isBp1Reversed, o->bp1RightOpen, &v.bp1Begin, &v.bp1End);
        adjust_breakend(r->beginPos1, p1, r->cigarLen1, 1, r->jumpRange, ref1[li].len, o->isBp2Reversed, o->bp2RightOpen, &v.bp2Begin, &v.bp2End);
      }
      v.insertBegin = (int32_t)path_read_length(p1, r->cigarLen1);
      v.insertEnd   = v.insertBegin + (int32_t)r->jumpInsertSize;
    }
    out[li] = v;
  }
  return MB200_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * 10-mer anchored reference trim  :1984-2011
 * ------------------------------------------------------------------------------------------------------------- */
#define MER 10

static uint32_t mer_hash(const uint8_t* p)
{
  uint32_t h = 2166136261u;
  for (int i = 0; i < MER; ++i) h = (h ^ p[i]) * 16777619u;
  return h;
}

static int has_mer(const int32_t* tab, uint32_t cap, const uint8_t* cs, const uint8_t* q)
{
  uint32_t h = mer_hash(q) & (cap - 1);
  while (tab[h] >= 0 && memcmp(cs + tab[h], q, MER) != 0) h = (h + 1) & (cap - 1);
  return tab[h] >= 0;
}

int orc_ref_trim_batch(uint32_t n, const uint8_t* pool, const mb200_seqview* contig, const mb200_seqview* ref,
                       const mb200_ref_cuts* cuts, mb200_trim_result* out)
{
  for (uint32_t it = 0; it < n; ++it) {
    const uint8_t* cs = pool + contig[it].off;
    const uint8_t* rs = pool + ref[it].off;
    const uint32_t cl = contig[it].len;
    if (cl < MER) return MB200_ERR_BAD_ARG;
    /* set of the contig's 10-mers: open addressing over contig positions */
    uint32_t cap = 16;
    while (cap < 2 * cl) cap <<= 1;
    int32_t* tab = (int32_t*)malloc(sizeof(int32_t) * cap);
    for (uint32_t i = 0; i < cap; ++i) tab[i] = -1;
    for (uint32_t p = 0; p + MER <= cl; ++p) {
      uint32_t h = mer_hash(cs + p) & (cap - 1);
      while (tab[h] >= 0 && memcmp(cs + tab[h], cs + p, MER) != 0) h = (h + 1) & (cap - 1);
      if (tab[h] < 0) tab[h] = (int32_t)p;
    }
#define HAS_MER(q) has_mer(tab, cap, cs, (q))
    const int32_t refSize     = (int32_t)ref[it].len;
    const int32_t minRefIndex = cuts[it].leadingCut;
    const int32_t maxRefIndex = refSize - (cuts[it].trailingCut + MER);
    const int32_t maxFwd      = cuts[it].maxLeadingCut < maxRefIndex ? cuts[it].maxLeadingCut : maxRefIndex;
    int32_t       ri;
    for (ri = minRefIndex; ri <= maxFwd; ++ri)
      if (HAS_MER(rs + ri)) break;
    out[it].adjLeadingCut = ri;
    const int32_t minRev = minRefIndex > (refSize - cuts[it].maxTrailingCut) ? minRefIndex : (refSize - cuts[it].maxTrailingCut);
    for (ri = maxRefIndex; ri >= minRev; --ri)
      if (HAS_MER(rs + ri)) break;
    out[it].adjTrailingCut = refSize - (ri + MER);
#undef HAS_MER
    free(tab);
  }
  return MB200_OK;
}
