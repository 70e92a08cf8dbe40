/*
 * orc_large_insert.c -- TEST INFRASTRUCTURE ONLY (see manta_oracle.h).
 *
 * C restatement of the large-insertion search of the refiner (rest of SURVEY.md row a9):
 *   isLargeInsertSegment / isLargeInsertAlignment     GSC/SVCandidateAssemblyRefiner.cpp:563-640
 *   the per-contig candidate test in getSmallSVAssembly  :2079-2112
 *   processLargeInsertion                               :833-1007  (with getLargestInsertSegment :230-279,
 *       isFinishedLargeInsertAlignment :642-668, getInsertTrim :801-830, setSmallCandSV :720-799, getVariantRange :677-717)
 *   getMaxPathScore                                     lib/alignment/AlignmentScoringUtilImpl.hpp:90-155
 *   apath_limit_read_length                             lib/blt_util/align_path.cpp:295-329
 * Pinned against the reference's own statics (oracle/_ref, ref_large_insertion_batch) by tests/test_large_insert.py.
 */
#include <stdlib.h>
#include <string.h>

#include "manta_oracle.h"

#define MIDDLE 100u                 /* the 'N' block between the two contigs (:927-929) */
#define MIN_SEMI_LARGE_INSERT 40    /* minSemiLargeInsertionLength (:556) */

static int seg_read(uint32_t t) { return t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH; }
static int seg_ref(uint32_t t) { return t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH; }

static uint32_t read_len(const uint32_t* s, uint32_t n)
{
  uint32_t v = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (seg_read(MB200_SEG_TYPE(s[i]))) v += MB200_SEG_LEN(s[i]);
  return v;
}
static uint32_t ref_len(const uint32_t* s, uint32_t n)
{
  uint32_t v = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (seg_ref(MB200_SEG_TYPE(s[i]))) v += MB200_SEG_LEN(s[i]);
  return v;
}

/* getMaxPathScore with isScoreOffEdge = true.  The reference declares `isIndel` INSIDE the loop, so every indel
 * segment pays the open score (the "unintended bug with positive results" of the source comment). */
static int max_path_score(const mb200_scores* sc, const uint32_t* s, uint32_t n, int reversed, uint32_t* maxRead, uint32_t* maxRef)
{
  int      val = 0, maxVal = 0;
  uint32_t readOff = 0, refOff = 0;
  *maxRead = *maxRef = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t seg = s[reversed ? n - 1 - k : k];
    const uint32_t t = MB200_SEG_TYPE(seg), len = MB200_SEG_LEN(seg);
    switch (t) {
    case MB200_SEG_SEQ_MATCH: val += sc->match * (int)len; readOff += len; refOff += len; break;
    case MB200_SEG_SEQ_MISMATCH: val += sc->mismatch * (int)len; readOff += len; refOff += len; break;
    case MB200_SEG_INSERT: val += sc->open + sc->extend * (int)len; readOff += len; break;
    case MB200_SEG_DELETE: val += sc->open + sc->extend * (int)len; refOff += len; break;
    case MB200_SEG_SOFT_CLIP: val += sc->offEdge * (int)len; readOff += len; break;
    default: break;
    }
    if (val > maxVal) { maxVal = val; *maxRead = readOff; *maxRef = refOff; }
  }
  return maxVal;
}

/* isLargeInsertSegment :563-608 */
static int large_insert_segment(const mb200_scores* sc, const uint32_t* s, uint32_t n, int reversed, uint32_t* contigOffset,
                                uint32_t* refOffset, int* score, uint32_t trimInsertLength)
{
  const uint32_t pathSize = read_len(s, n);
  const int      m = max_path_score(sc, s, n, reversed, contigOffset, refOffset);
  *score = m > 0 ? m : 0;
  if (*refOffset < 40) return 0;
  if (*contigOffset < 40) return 0;
  if ((pathSize - *contigOffset) < (MIN_SEMI_LARGE_INSERT + trimInsertLength)) return 0;
  const int   optimal = (int)(*contigOffset) * sc->match;
  const float frac = (float)(*score) / (float)optimal;
  if (frac < 0.75f) return 0;
  return 1;
}

/* isLargeInsertAlignment :612-640 */
static int large_insert_alignment(const mb200_scores* sc, const uint32_t* s, uint32_t n, mb200_large_insert_info* o)
{
  memset(o, 0, sizeof(*o));
  o->isLeftCandidate = (uint8_t)large_insert_segment(sc, s, n, 0, &o->contigOffset, &o->refOffset, &o->score, 0);
  if (o->isLeftCandidate) return 1;
  o->isRightCandidate = (uint8_t)large_insert_segment(sc, s, n, 1, &o->contigOffset, &o->refOffset, &o->score, 0);
  if (o->isRightCandidate) {
    o->contigOffset = read_len(s, n) - o->contigOffset;
    o->refOffset    = ref_len(s, n) - o->refOffset;
    return 1;
  }
  return 0;
}

/* apath_limit_read_length(begin, end, apath) into out[] (<= n segments); returns the new size */
static uint32_t limit_read_length(const uint32_t* s, uint32_t n, uint32_t tBegin, uint32_t tEnd, uint32_t* out)
{
  int      startSet = 0;
  uint32_t readLength = 0, startSeg = 0, endSeg = n;
  memcpy(out, s, sizeof(uint32_t) * n);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t t = MB200_SEG_TYPE(out[i]);
    uint32_t       len = MB200_SEG_LEN(out[i]);
    if (!seg_read(t)) continue;
    readLength += len;
    if (!startSet && readLength > tBegin) {
      const uint32_t extra = len - (readLength - tBegin);
      len -= extra;
      out[i]   = MB200_SEG_MAKE(len, t);
      startSeg = i;
      startSet = 1;
    }
    if (readLength >= tEnd) {
      if (readLength > tEnd) {
        const uint32_t extra = readLength - tEnd;
        len -= extra;
        out[i] = MB200_SEG_MAKE(len, t);
      }
      endSeg = i + 1;
      break;
    }
  }
  if (endSeg < startSeg) return 0;
  memmove(out, out + startSeg, sizeof(uint32_t) * (endSeg - startSeg));
  return endSeg - startSeg;
}

/* getVariantRange :677-717 (the left scan stops at index 0 where the reference would read index -1: same rule as
 * orc_small_sv_batch) */
static void variant_range(const uint8_t* ref, int64_t refLen, int32_t refB, int32_t refE, const uint8_t* read, int64_t readLen,
                          int32_t readB, int32_t readE, int32_t* left, int32_t* right)
{
  const int64_t maxRight = (refLen - refE) < (readLen - readE) ? (refLen - refE) : (readLen - readE);
  int64_t       r = 0;
  for (; r < maxRight; ++r)
    if (ref[refB + r] != read[readB + r]) break;
  const int32_t minLeft = (-refB) > (-readB) ? (-refB) : (-readB);
  int32_t       l = 0;
  for (; l >= minLeft; --l) {
    const int64_t ri = (int64_t)refE + l - 1, qi = (int64_t)readE + l - 1;
    if (ri < 0 || qi < 0) break;
    if (ref[ri] != read[qi]) break;
  }
  *left  = l;
  *right = (int32_t)r;
}

int orc_large_insertion_batch(const mb200_scores* edgeScores, const mb200_scores* completeScores, uint32_t nLoci,
                              const uint8_t* isJump, const uint32_t* slotFirst, const uint32_t* nContigs, uint32_t nSlots,
                              const mb200_aln_result* aln, const uint32_t* cigarPool, const uint8_t* pool,
                              const mb200_seqview* contigSeq, const int32_t* consRange, const mb200_seqview* ref1,
                              const mb200_ref_cuts* cuts, const uint32_t* svFirst, const mb200_small_sv* svs,
                              mb200_large_insert_info* info, mb200_large_insertion* out, uint32_t* outCigar, uint64_t outCigarCap,
                              uint64_t* outCigarUsed)
{
  uint64_t used = 0;
  int      overflow = 0;
  memset(info, 0, sizeof(*info) * nSlots);
  memset(out, 0, sizeof(*out) * nLoci);
  for (uint32_t li = 0; li < nLoci; ++li) {
    if (isJump[li]) continue;
    const uint32_t s0 = slotFirst[li], nc = nContigs[li];
    /* ---- per-contig candidate test (:2079-2112) */
    for (uint32_t k = 0; k < nc; ++k) {
      const mb200_aln_result* r = &aln[s0 + k];
      const uint32_t*         path = cigarPool + r->cigarOff;
      const uint32_t          n = r->cigarLen;
      uint32_t*               lim = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
      const uint32_t          nl = limit_read_length(path, n, (uint32_t)consRange[2 * (s0 + k)], (uint32_t)consRange[2 * (s0 + k) + 1], lim);
      mb200_large_insert_info a, b;
      int                     isCand = large_insert_alignment(edgeScores, lim, nl, &a);
      if (isCand) {
        isCand = large_insert_alignment(edgeScores, path, n, &b);
        if (a.isLeftCandidate != b.isLeftCandidate || a.isRightCandidate != b.isRightCandidate) isCand = 0;
        a.contigOffset = b.contigOffset;
        a.refOffset    = b.refOffset;
      }
      if (isCand) info[s0 + k] = a;
      free(lim);
    }
    /* ---- processLargeInsertion: pair search (:870-915) */
    mb200_large_insertion* o = &out[li];
    int      isPair = 0, bestDist = 0, bestScore = 0;
    uint32_t left = 0, right = 0;
    for (uint32_t c1 = 0; c1 + 1 < nc; ++c1) {
      const mb200_large_insert_info* i1 = &info[s0 + c1];
      if (!i1->isLeftCandidate && !i1->isRightCandidate) continue;
      for (uint32_t c2 = c1 + 1; c2 < nc; ++c2) {
        const mb200_large_insert_info* i2 = &info[s0 + c2];
        if (!i2->isLeftCandidate && !i2->isRightCandidate) continue;
        if (!((i1->isLeftCandidate && i2->isRightCandidate) || (i2->isLeftCandidate && i1->isRightCandidate))) continue;
        const long d = (long)(aln[s0 + c1].beginPos + (long)i1->refOffset) - (long)(aln[s0 + c2].beginPos + (long)i2->refOffset);
        const int  dist = (int)(d < 0 ? -d : d);
        if (dist > 35) continue;
        const int sc = i1->score + i2->score;
        const int better = (dist < bestDist) || (dist == bestDist && sc > bestScore);
        if (!isPair || better) {
          isPair = 1;
          left   = c1;
          right  = c2;
          if (i1->isRightCandidate) { left = c2; right = c1; }
          bestDist  = dist;
          bestScore = sc;
        }
      }
    }
    if (!isPair) continue; /* MB200_LI_NONE */
    o->leftContig = left; o->rightContig = right; o->breakDist = bestDist; o->breakScore = bestScore;
    /* ---- joined contig, aligned inside the original cuts (:919-941) */
    const mb200_seqview lv = contigSeq[s0 + left], rv = contigSeq[s0 + right];
    const uint32_t      fakeLen = lv.len + MIDDLE + rv.len;
    uint8_t*            fake = (uint8_t*)malloc(fakeLen);
    memcpy(fake, pool + lv.off, lv.len);
    memset(fake + lv.len, 'N', MIDDLE);
    memcpy(fake + lv.len + MIDDLE, pool + rv.off, rv.len);
    const int32_t lead = cuts ? cuts[li].leadingCut : 0, trail = cuts ? cuts[li].trailingCut : 0;
    mb200_seqview q, rr;
    q.off = (uint64_t)(fake - pool); q.len = fakeLen; q.reserved = 0;
    rr = ref1[li];
    rr.off += (uint64_t)lead;
    rr.len -= (uint32_t)(lead + trail);
    mb200_aln_result fa;
    uint32_t*        fcig = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(fakeLen + rr.len + 8));
    uint64_t         fused = 0;
    orc_global_align_batch(completeScores, 1, pool, &q, &rr, &fa, fcig, (uint64_t)fakeLen + rr.len + 8, &fused);
    fa.beginPos += lead;
    const uint32_t* fp = fcig + fa.cigarOff;
    const uint32_t  fn = fa.cigarLen;
    o->score = fa.score; o->beginPos = fa.beginPos; o->cigarLen = fn; o->cigarOff = used;
    if (used + fn <= outCigarCap) memcpy(outCigar + used, fp, sizeof(uint32_t) * fn);
    else overflow = 1;
    used += fn;
    o->status = MB200_LI_SEGMENT_QC;
    do {
      /* getLargestInsertSegment(apath, 100) :230-279 */
      int      inSeg = 0, isCand = 0, isMax = 0;
      uint32_t segStart = 0, maxSize = MIDDLE, mFirst = 0, mLast = 0;
      for (uint32_t i = 0; i < fn; ++i) {
        const uint32_t t = MB200_SEG_TYPE(fp[i]), len = MB200_SEG_LEN(fp[i]);
        if (t == MB200_SEG_DELETE || t == MB200_SEG_INSERT) {
          if (t == MB200_SEG_INSERT && len >= maxSize) { isMax = 1; maxSize = len; isCand = 1; }
          if (!inSeg) segStart = i;
          inSeg = 1;
        } else {
          if (isCand) { mFirst = segStart; mLast = i - 1; }
          inSeg = 0;
          isCand = 0;
        }
      }
      if (isCand) { mFirst = segStart; mLast = fn - 1; }
      if (!isMax || mLast < mFirst) break;
      o->segFirst = mFirst; o->segLast = mLast;
      o->status = MB200_LI_FLANK_QC;
      /* isFinishedLargeInsertAlignment :642-668 */
      uint32_t co, ro;
      int      sc;
      if (!large_insert_segment(completeScores, fp, mLast + 1, 0, &co, &ro, &sc, MIDDLE)) break;
      if (!large_insert_segment(completeScores, fp + mFirst, fn - mFirst, 1, &co, &ro, &sc, MIDDLE)) break;
      o->status = MB200_LI_TRIM_QC;
      /* getInsertTrim :801-830 */
      int32_t readPos = 0, tb = 0, te = 0;
      for (uint32_t i = 0; i < fn; ++i) {
        if (i == mFirst) tb = readPos;
        if (seg_read(MB200_SEG_TYPE(fp[i]))) readPos += (int32_t)MB200_SEG_LEN(fp[i]);
        if (i == mLast) { te = readPos; break; }
      }
      o->insertTrimBegin = tb; o->insertTrimEnd = te;
      if (tb + MIN_SEMI_LARGE_INSERT > (int32_t)lv.len) break;
      const int32_t rightOffset = (int32_t)(lv.len + MIDDLE);
      if (rightOffset + MIN_SEMI_LARGE_INSERT > te) break;
      o->status = MB200_LI_RANGE_ERROR;
      /* setSmallCandSV on the joined contig (:720-799) */
      int32_t rp = 0, fpos = fa.beginPos;
      for (uint32_t i = 0; i < fn; ++i) {
        const uint32_t t = MB200_SEG_TYPE(fp[i]), len = MB200_SEG_LEN(fp[i]);
        if (i == mFirst) { o->refBegin = fpos; o->readBegin = rp; }
        if (seg_ref(t)) fpos += (int32_t)len;
        if (seg_read(t)) rp += (int32_t)len;
        if (i == mLast) { o->refEnd = fpos; o->readEnd = rp; }
      }
      variant_range(pool + ref1[li].off, ref1[li].len, o->refBegin, o->refEnd, fake, fakeLen, o->readBegin, o->readEnd,
                    &o->ciposBegin, &o->ciposEnd);
      if (o->ciposBegin != 0) break;
      o->status   = MB200_LI_DUPLICATE;
      o->bp1Begin = o->refBegin - 1;
      o->bp1End   = o->bp1Begin + o->ciposEnd + 1;
      o->bp2Begin = o->refEnd;
      o->bp2End   = o->bp2Begin + o->ciposEnd + 1;
      /* excludedPos: start positions of the fully assembled insertions of the selected contig (:2253-2256,
       * classifyIndel lib/manta/SVCandidateUtil.cpp:99-112: insert when deleteSize < insertSize) */
      int dup = 0;
      if (svFirst && svs) {
        for (uint32_t v = svFirst[li]; v < svFirst[li + 1]; ++v) {
          if (svs[v].status != 0) continue;
          const uint32_t deleteSize = (uint32_t)(svs[v].bp2Begin - svs[v].bp1Begin);
          const uint32_t insertSize = (uint32_t)(svs[v].readEnd - svs[v].readBegin);
          if (deleteSize < insertSize && svs[v].bp1Begin == o->bp1Begin) dup = 1;
        }
      }
      if (dup) break;
      o->status       = MB200_LI_ACCEPTED;
      o->leftSeqBegin = (uint32_t)tb;
      o->rightSeqLen  = (uint32_t)(te - rightOffset);
    } while (0);
    free(fcig);
    free(fake);
  }
  (void)nSlots;
  *outCigarUsed = used;
  return overflow ? MB200_ERR_POOL_OVERFLOW : MB200_OK;
}
