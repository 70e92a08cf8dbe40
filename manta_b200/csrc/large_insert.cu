// large_insert.cu -- large insertions whose insert sequence is only partly assembled (rest of SURVEY.md row a9).
//
// Replaces, with identical results, the refiner's search of GSC/SVCandidateAssemblyRefiner.cpp:
//   * the per-contig test in getSmallSVAssembly (:2079-2112): isLargeInsertAlignment (:612-640) on the alignment limited to
//     the contig's conservative range (apath_limit_read_length, lib/blt_util/align_path.cpp:295-329) and on the whole
//     alignment, with isLargeInsertSegment (:563-608) / getMaxPathScore (lib/alignment/AlignmentScoringUtilImpl.hpp:90-155);
//   * processLargeInsertion (:833-1007): the all-against-all pairing of left and right candidates, the joined contig
//     left + 100 'N' + right aligned by GlobalAligner<int> inside the original window cuts, getLargestInsertSegment (:230),
//     isFinishedLargeInsertAlignment (:642), getInsertTrim (:801), setSmallCandSV (:720) with getVariantRange (:677) and the
//     excluded-position rule (:2253-2256, :991).
//
// Three small kernels around the existing GlobalAligner DP (dp_kernels.cuh, KIND_GLOBAL): (1) one thread per contig scores
// its alignment path both ways; (2) one thread per candidate picks the pair and the warp writes the joined contig into a
// device pool; the DP runs on those; (3) one warp per candidate runs the acceptance tests on the joined alignment (lane 0
// walks the CIGAR, all lanes compare bases for the variant range).  CIGARs are tens of segments: this stage is latency,
// not bandwidth (a few ms per 50 000 candidates).
#include <cstdint>
#include <vector>

#include "mb200_internal.h"

namespace {

constexpr uint32_t MIDDLE                = 100;  // the 'N' block (:927-929)
constexpr int      MIN_SEMI_LARGE_INSERT = 40;   // minSemiLargeInsertionLength (:556)
constexpr int      MAX_BREAK_DIST        = 35;   // :866

__device__ __forceinline__ bool segRead(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH;
}
__device__ __forceinline__ bool segRef(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH;
}

// A path, or a sub-range of one with the first / last segment shortened (what apath_limit_read_length leaves), walked
// forwards or backwards without being copied.
struct PathView {
  const uint32_t* s;
  uint32_t        n;                  // segments
  uint32_t        cutFirst, cutLast;  // bases removed from segment 0 / segment n-1
  __device__ uint32_t seg(uint32_t i) const
  {
    uint32_t v = s[i], len = MB200_SEG_LEN(v);
    if (i == 0) len -= cutFirst;
    if (i == n - 1) len -= cutLast;
    return MB200_SEG_MAKE(len, MB200_SEG_TYPE(v));
  }
};

__device__ void pathLengths(const PathView& p, uint32_t& readLen, uint32_t& refLen)
{
  readLen = refLen = 0;
  for (uint32_t i = 0; i < p.n; ++i) {
    const uint32_t v = p.seg(i), t = MB200_SEG_TYPE(v), len = MB200_SEG_LEN(v);
    if (segRead(t)) readLen += len;
    if (segRef(t)) refLen += len;
  }
}

// getMaxPathScore (isScoreOffEdge = true).  `isIndel` is declared inside the reference's loop, so every indel segment pays
// the open score.
__device__ int maxPathScore(const mb200_scores& sc, const PathView& p, bool reversed, uint32_t& maxRead, uint32_t& maxRef)
{
  int      val = 0, maxVal = 0;
  uint32_t readOff = 0, refOff = 0;
  maxRead = maxRef = 0;
  for (uint32_t k = 0; k < p.n; ++k) {
    const uint32_t v = p.seg(reversed ? p.n - 1 - k : k), t = MB200_SEG_TYPE(v), len = MB200_SEG_LEN(v);
    if (t == MB200_SEG_SEQ_MATCH) { val += sc.match * (int)len; readOff += len; refOff += len; }
    else if (t == MB200_SEG_SEQ_MISMATCH) { val += sc.mismatch * (int)len; readOff += len; refOff += len; }
    else if (t == MB200_SEG_INSERT) { val += sc.open + sc.extend * (int)len; readOff += len; }
    else if (t == MB200_SEG_DELETE) { val += sc.open + sc.extend * (int)len; refOff += len; }
    else if (t == MB200_SEG_SOFT_CLIP) { val += sc.offEdge * (int)len; readOff += len; }
    if (val > maxVal) { maxVal = val; maxRead = readOff; maxRef = refOff; }
  }
  return maxVal;
}

// isLargeInsertSegment :563-608
__device__ bool largeInsertSegment(const mb200_scores& sc, const PathView& p, bool reversed, uint32_t& contigOffset, uint32_t& refOffset,
                                   int& score, uint32_t trimInsertLength)
{
  uint32_t pathSize, refSize;
  pathLengths(p, pathSize, refSize);
  const int m = maxPathScore(sc, p, reversed, contigOffset, refOffset);
  score = m > 0 ? m : 0;
  if (refOffset < 40) return false;
  if (contigOffset < 40) return false;
  if ((pathSize - contigOffset) < ((uint32_t)MIN_SEMI_LARGE_INSERT + trimInsertLength)) return false;
  const int   optimal = (int)contigOffset * sc.match;
  const float frac    = __fdiv_rn((float)score, (float)optimal);
  return !(frac < 0.75f);
}

// isLargeInsertAlignment :612-640
__device__ bool largeInsertAlignment(const mb200_scores& sc, const PathView& p, mb200_large_insert_info& o)
{
  o = mb200_large_insert_info();
  o.isLeftCandidate = largeInsertSegment(sc, p, false, o.contigOffset, o.refOffset, o.score, 0) ? 1 : 0;
  if (o.isLeftCandidate) return true;
  o.isRightCandidate = largeInsertSegment(sc, p, true, o.contigOffset, o.refOffset, o.score, 0) ? 1 : 0;
  if (o.isRightCandidate) {
    uint32_t rl, fl;
    pathLengths(p, rl, fl);
    o.contigOffset = rl - o.contigOffset;
    o.refOffset    = fl - o.refOffset;
    return true;
  }
  return false;
}

// apath_limit_read_length as a view
__device__ PathView limitReadLength(const uint32_t* s, uint32_t n, uint32_t tBegin, uint32_t tEnd)
{
  bool     startSet = false;
  uint32_t readLength = 0, startSeg = 0, endSeg = n, cutFirst = 0, cutLast = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t t = MB200_SEG_TYPE(s[i]), len = MB200_SEG_LEN(s[i]);
    if (!segRead(t)) continue;
    readLength += len;
    if (!startSet && readLength > tBegin) {
      cutFirst = len - (readLength - tBegin);
      startSeg = i;
      startSet = true;
    }
    if (readLength >= tEnd) {
      if (readLength > tEnd) cutLast = readLength - tEnd;
      endSeg = i + 1;
      break;
    }
  }
  PathView p;
  if (endSeg < startSeg) { p.s = s; p.n = 0; p.cutFirst = p.cutLast = 0; return p; }
  p.s = s + startSeg; p.n = endSeg - startSeg; p.cutFirst = cutFirst; p.cutLast = cutLast;
  return p;
}

struct LiParams {
  mb200_scores             edge, complete;
  uint32_t                 nLoci, nSlots;
  const uint8_t*           isJump;
  const uint32_t*          slotFirst;
  const uint32_t*          nContigs;
  const uint32_t*          slotLocus;   // [nSlots]
  const mb200_aln_result*  aln;
  const uint32_t*          cigar;
  const uint8_t*           pool;
  const mb200_seqview*     contigSeq;
  const int32_t*           consRange;
  const mb200_seqview*     ref1;
  const mb200_ref_cuts*    cuts;       // may be null
  const uint32_t*          svFirst;    // may be null
  const mb200_small_sv*    svs;
  mb200_large_insert_info* info;
  mb200_large_insertion*   out;
  // joined contigs and their DP
  uint8_t*                 fakePool;   // MB200_POOL_FRONT bytes in front
  unsigned long long       fakeCap;
  unsigned long long*      fakeUsed;
  mb200_seqview*           fakeView;   // [nLoci]: joined contig of the candidate (len 0: none)
  const mb200_aln_result*  fakeAln;    // [nLoci]
  const uint32_t*          fakeCigar;
};

__global__ void liInfoKernel(const LiParams P)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.nSlots) return;
  mb200_large_insert_info o = mb200_large_insert_info();
  const uint32_t li = P.slotLocus[s];
  if (!P.isJump[li]) {
    const mb200_aln_result r = P.aln[s];
    const uint32_t*        path = P.cigar + r.cigarOff;
    PathView               full;
    full.s = path; full.n = r.cigarLen; full.cutFirst = full.cutLast = 0;
    const PathView          lim = limitReadLength(path, r.cigarLen, (uint32_t)P.consRange[2 * s], (uint32_t)P.consRange[2 * s + 1]);
    mb200_large_insert_info a, b;
    bool                    isCand = largeInsertAlignment(P.edge, lim, a);
    if (isCand) {
      isCand = largeInsertAlignment(P.edge, full, b);
      if (a.isLeftCandidate != b.isLeftCandidate || a.isRightCandidate != b.isRightCandidate) isCand = false;
      a.contigOffset = b.contigOffset;
      a.refOffset    = b.refOffset;
    }
    if (isCand) o = a;
  }
  P.info[s] = o;
}

// one warp per candidate: lane 0 searches the pair (<= 16 contigs), the warp copies the joined contig
__global__ void __launch_bounds__(128) liPairKernel(const LiParams P)
{
  const int      lane = threadIdx.x & 31;
  const uint32_t li   = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (li >= P.nLoci) return;
  mb200_large_insertion o = mb200_large_insertion();
  mb200_seqview         fv;
  fv.off = 0; fv.len = 0; fv.reserved = 0;
  uint32_t left = 0, right = 0;
  int      isPair = 0;
  if (!P.isJump[li]) {
    const uint32_t s0 = P.slotFirst[li], nc = P.nContigs[li];
    int bestDist = 0, bestScore = 0;
    for (uint32_t c1 = 0; c1 + 1 < nc; ++c1) {
      const mb200_large_insert_info i1 = P.info[s0 + c1];
      if (!i1.isLeftCandidate && !i1.isRightCandidate) continue;
      for (uint32_t c2 = c1 + 1; c2 < nc; ++c2) {
        const mb200_large_insert_info i2 = P.info[s0 + c2];
        if (!((i1.isLeftCandidate && i2.isRightCandidate) || (i2.isLeftCandidate && i1.isRightCandidate))) continue;
        const long long d = ((long long)P.aln[s0 + c1].beginPos + (long long)i1.refOffset) - ((long long)P.aln[s0 + c2].beginPos + (long long)i2.refOffset);
        const long long ad = d < 0 ? -d : d;
        if (ad > MAX_BREAK_DIST) continue;
        const int dist = (int)ad, sc = i1.score + i2.score;
        const bool better = (dist < bestDist) || (dist == bestDist && sc > bestScore);
        if (!isPair || better) {
          isPair = 1;
          left   = c1;
          right  = c2;
          if (i1.isRightCandidate) { left = c2; right = c1; }
          bestDist  = dist;
          bestScore = sc;
        }
      }
    }
    if (isPair) {
      o.leftContig = left; o.rightContig = right; o.breakDist = bestDist; o.breakScore = bestScore;
      o.status = MB200_LI_SEGMENT_QC;  // refined by liFinishKernel
      const mb200_seqview lv = P.contigSeq[s0 + left], rv = P.contigSeq[s0 + right];
      const uint32_t      fakeLen = lv.len + MIDDLE + rv.len;
      unsigned long long  at = 0;
      if (lane == 0) at = atomicAdd(P.fakeUsed, (unsigned long long)((fakeLen + 15u) & ~15u));
      at = __shfl_sync(0xffffffffu, at, 0);
      if (at + fakeLen <= P.fakeCap) {
        uint8_t*       dst = P.fakePool + at;
        const uint8_t* l   = P.pool + lv.off;
        const uint8_t* r   = P.pool + rv.off;
        for (uint32_t i = lane; i < fakeLen; i += 32)
          dst[i] = i < lv.len ? l[i] : (i < lv.len + MIDDLE ? (uint8_t)'N' : r[i - lv.len - MIDDLE]);
        fv.off = at;
        fv.len = fakeLen;
      } else {
        o.status = MB200_ERR_POOL_OVERFLOW;  // host sizes the pool from the contig lengths: cannot happen
      }
    }
  }
  if (lane == 0) {
    P.out[li]      = o;
    P.fakeView[li] = fv;
  }
}

// acceptance tests on the joined alignment: one warp per candidate
__global__ void __launch_bounds__(128) liFinishKernel(const LiParams P)
{
  const int      lane = threadIdx.x & 31;
  const uint32_t li   = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (li >= P.nLoci) return;
  const mb200_seqview fv = P.fakeView[li];
  if (fv.len == 0) return;
  mb200_large_insertion  o  = P.out[li];
  const mb200_aln_result fa = P.fakeAln[li];
  const int32_t          lead = P.cuts ? P.cuts[li].leadingCut : 0;
  const uint32_t*        fp = P.fakeCigar + fa.cigarOff;
  const uint32_t         fn = fa.cigarLen;
  const uint32_t         s0 = P.slotFirst[li];
  const uint32_t         leftLen = P.contigSeq[s0 + o.leftContig].len;
  o.score = fa.score; o.beginPos = fa.beginPos + lead; o.cigarLen = fn; o.cigarOff = fa.cigarOff;
  o.status = MB200_LI_SEGMENT_QC;
  bool goOn = true;
  uint32_t mFirst = 0, mLast = 0;
  {
    // getLargestInsertSegment(apath, 100) :230-279 (every lane walks the short path: uniform control flow)
    bool     inSeg = false, isCand = false, isMax = false;
    uint32_t segStart = 0, maxSize = MIDDLE;
    for (uint32_t i = 0; i < fn; ++i) {
      const uint32_t t = MB200_SEG_TYPE(fp[i]), len = MB200_SEG_LEN(fp[i]);
      if (t == MB200_SEG_DELETE || t == MB200_SEG_INSERT) {
        if (t == MB200_SEG_INSERT && len >= maxSize) { isMax = true; maxSize = len; isCand = true; }
        if (!inSeg) segStart = i;
        inSeg = true;
      } else {
        if (isCand) { mFirst = segStart; mLast = i - 1; }
        inSeg  = false;
        isCand = false;
      }
    }
    if (isCand) { mFirst = segStart; mLast = fn - 1; }
    if (!isMax || mLast < mFirst) goOn = false;
  }
  if (goOn) {
    o.segFirst = mFirst; o.segLast = mLast;
    o.status = MB200_LI_FLANK_QC;
    // isFinishedLargeInsertAlignment :642-668
    uint32_t co, ro;
    int      sc;
    PathView pl, pr;
    pl.s = fp; pl.n = mLast + 1; pl.cutFirst = pl.cutLast = 0;
    pr.s = fp + mFirst; pr.n = fn - mFirst; pr.cutFirst = pr.cutLast = 0;
    if (!largeInsertSegment(P.complete, pl, false, co, ro, sc, MIDDLE)) goOn = false;
    else if (!largeInsertSegment(P.complete, pr, true, co, ro, sc, MIDDLE)) goOn = false;
  }
  int32_t tb = 0, te = 0;
  const int32_t rightOffset = (int32_t)(leftLen + MIDDLE);
  if (goOn) {
    o.status = MB200_LI_TRIM_QC;
    int32_t readPos = 0;  // getInsertTrim :801-830
    for (uint32_t i = 0; i < fn; ++i) {
      if (i == mFirst) tb = readPos;
      if (segRead(MB200_SEG_TYPE(fp[i]))) readPos += (int32_t)MB200_SEG_LEN(fp[i]);
      if (i == mLast) { te = readPos; break; }
    }
    o.insertTrimBegin = tb; o.insertTrimEnd = te;
    if (tb + MIN_SEMI_LARGE_INSERT > (int32_t)leftLen) goOn = false;
    else if (rightOffset + MIN_SEMI_LARGE_INSERT > te) goOn = false;
  }
  if (goOn) {
    o.status = MB200_LI_RANGE_ERROR;
    // setSmallCandSV on the joined contig (:720-799)
    int32_t rp = 0, fpos = o.beginPos;
    for (uint32_t i = 0; i < fn; ++i) {
      const uint32_t t = MB200_SEG_TYPE(fp[i]), len = MB200_SEG_LEN(fp[i]);
      if (i == mFirst) { o.refBegin = fpos; o.readBegin = rp; }
      if (segRef(t)) fpos += (int32_t)len;
      if (segRead(t)) rp += (int32_t)len;
      if (i == mLast) { o.refEnd = fpos; o.readEnd = rp; }
    }
    // getVariantRange (:677-717): lanes compare 32 positions per round (as smallSvKernel, contig_qc.cu)
    const uint8_t* ref     = P.pool + P.ref1[li].off;
    const int64_t  refLen  = P.ref1[li].len;
    const uint8_t* read    = P.fakePool + fv.off;
    const int64_t  readLen = fv.len;
    const int64_t  maxRight = min(refLen - o.refEnd, readLen - o.readEnd);
    int32_t        right = (int32_t)max((int64_t)0, maxRight);
    for (int64_t base = 0; base < maxRight; base += 32) {
      const int64_t  k    = base + lane;
      const bool     stop = (k >= maxRight) || (ref[o.refBegin + k] != read[o.readBegin + k]);
      const uint32_t b    = __ballot_sync(0xffffffffu, stop);
      if (b) { right = (int32_t)(base + (__ffs(b) - 1)); break; }
    }
    const int32_t minLeft = max(-o.refBegin, -o.readBegin);
    int32_t       left    = 0;
    for (int32_t base = 0;; base += 32) {
      const int32_t  j  = base + lane;
      const int32_t  ri = o.refEnd - j - 1, qi = o.readEnd - j - 1;
      const bool     stop = (-j < minLeft) || ri < 0 || qi < 0 || (ref[ri] != read[qi]);
      const uint32_t b    = __ballot_sync(0xffffffffu, stop);
      if (b) { left = -(base + (__ffs(b) - 1)); break; }
    }
    o.ciposBegin = left; o.ciposEnd = right;
    if (left != 0) goOn = false;
  }
  if (goOn) {
    o.status   = MB200_LI_DUPLICATE;
    o.bp1Begin = o.refBegin - 1;
    o.bp1End   = o.bp1Begin + o.ciposEnd + 1;
    o.bp2Begin = o.refEnd;
    o.bp2End   = o.bp2Begin + o.ciposEnd + 1;
    // a fully assembled insertion of the selected contig at the same start (:2253-2256, :991)
    bool dup = false;
    if (P.svFirst && P.svs) {
      for (uint32_t v = P.svFirst[li]; v < P.svFirst[li + 1]; ++v) {
        const mb200_small_sv& w = P.svs[v];
        if (w.status != 0) continue;
        const uint32_t deleteSize = (uint32_t)(w.bp2Begin - w.bp1Begin), insertSize = (uint32_t)(w.readEnd - w.readBegin);
        if (deleteSize < insertSize && w.bp1Begin == o.bp1Begin) dup = true;
      }
    }
    if (!dup) {
      o.status       = MB200_LI_ACCEPTED;
      o.leftSeqBegin = (uint32_t)tb;
      o.rightSeqLen  = (uint32_t)(te - rightOffset);
    }
  }
  if (lane == 0) P.out[li] = o;
}

}  // namespace

extern "C" int mb200_large_insertion_batch(mb200_ctx* ctx, const mb200_scores* edgeScores, const mb200_scores* completeScores,
                                           uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst,
                                           const uint32_t* nContigs, uint32_t nSlots, const mb200_aln_result* aln,
                                           const uint32_t* cigarPool, uint64_t cigarLen, const uint8_t* pool, uint64_t poolBytes,
                                           const mb200_seqview* contigSeq, const int32_t* consRange, const mb200_seqview* ref1,
                                           const mb200_ref_cuts* cuts, const uint32_t* svFirst, const mb200_small_sv* svs,
                                           mb200_large_insert_info* info, mb200_large_insertion* out, uint32_t* outCigar,
                                           uint64_t outCigarCap, uint64_t* outCigarUsed)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!outCigarUsed) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  *outCigarUsed = 0;
  if (nLoci == 0) return MB200_OK;
  if (!edgeScores || !completeScores || !isJump || !slotFirst || !nContigs || !ref1 || !pool || !info || !out || (outCigarCap && !outCigar))
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (nSlots && (!aln || !cigarPool || !contigSeq || !consRange)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if ((svFirst == nullptr) != (svs == nullptr) && svFirst && svFirst[nLoci] != 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "svFirst without svs");
  // host-side validation + sizes of the joined contigs
  std::vector<uint32_t> slotLocus(nSlots ? nSlots : 1);
  uint64_t              fakeCap = 0, maxFake = 0;
  for (uint32_t li = 0; li < nLoci; ++li) {
    if ((uint64_t)slotFirst[li] + nContigs[li] > nSlots) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "contig slots outside the result arrays");
    if (nContigs[li] > 16) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "more than 16 contigs in a candidate");
    if (ref1[li].off + ref1[li].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "reference window outside the pool");
    if (cuts && (cuts[li].leadingCut < 0 || cuts[li].trailingCut < 0 || (uint64_t)cuts[li].leadingCut + cuts[li].trailingCut >= ref1[li].len))
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "window cuts leave no reference");
    uint32_t m1 = 0, m2 = 0;  // the two longest contigs bound the joined length
    for (uint32_t k = 0; k < nContigs[li]; ++k) {
      const uint32_t s = slotFirst[li] + k;
      slotLocus[s]     = li;
      if (isJump[li]) continue;
      if (contigSeq[s].off + contigSeq[s].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "contig outside the pool");
      if (aln[s].cigarOff + aln[s].cigarLen > cigarLen) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "path outside the cigar pool");
      if (consRange[2 * s] < 0 || consRange[2 * s + 1] < consRange[2 * s]) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "bad conservative range");
      const uint32_t L = contigSeq[s].len;
      if (L > m1) { m2 = m1; m1 = L; } else if (L > m2) m2 = L;
    }
    if (!isJump[li] && nContigs[li] >= 2) {
      const uint64_t f = ((uint64_t)m1 + m2 + MIDDLE + 15) & ~15ull;
      fakeCap += f;
      maxFake = std::max<uint64_t>(maxFake, f);
    }
  }
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;
  int          rc;
  const size_t ns = nSlots ? nSlots : 1;
  const size_t nSv = (svFirst && svs) ? svFirst[nLoci] : 0;
  enum { B_POOL, B_CIGAR, B_ALN, B_VIEWS, B_CONS, B_REF, B_CUTS, B_SLOTFIRST, B_NCONTIGS, B_SLOTLOCUS, B_ISJUMP, B_SV, B_OUT, B_FAKE };
  DeviceBuffer* B = ctx->dLi;
  const size_t  outBytes = sizeof(mb200_large_insert_info) * ns + sizeof(mb200_large_insertion) * nLoci + sizeof(mb200_seqview) * nLoci +
                          sizeof(mb200_aln_result) * nLoci + 256;
  if ((rc = mb200_reserve(ctx, B[B_POOL], poolBytes + MB200_POOL_FRONT + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_CIGAR], 4 * (size_t)(cigarLen ? cigarLen : 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_ALN], sizeof(mb200_aln_result) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_VIEWS], sizeof(mb200_seqview) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_CONS], 8 * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_REF], sizeof(mb200_seqview) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_CUTS], sizeof(mb200_ref_cuts) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_SLOTFIRST], 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_NCONTIGS], 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_SLOTLOCUS], 4 * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_ISJUMP], (size_t)nLoci + 4)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_SV], 4 * (size_t)(nLoci + 1) + sizeof(mb200_small_sv) * (nSv ? nSv : 1) + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_OUT], outBytes)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, B[B_FAKE], fakeCap + MB200_POOL_FRONT + 128)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dMisc, 256)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dCigar, 4 * (size_t)(outCigarCap ? outCigarCap : 1))) != MB200_OK) return rc;
  cudaStream_t st = ctx->stream;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], st));
#define MB200_UP(dst, src, bytes) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, st))
  uint8_t* dPoolBase = reinterpret_cast<uint8_t*>(B[B_POOL].ptr);
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(dPoolBase, 0, MB200_POOL_FRONT, st));
  MB200_UP(dPoolBase + MB200_POOL_FRONT, pool, poolBytes);
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(dPoolBase + MB200_POOL_FRONT + poolBytes, 0, 64, st));
  if (cigarLen) MB200_UP(B[B_CIGAR].ptr, cigarPool, 4 * (size_t)cigarLen);
  if (nSlots) {
    MB200_UP(B[B_ALN].ptr, aln, sizeof(mb200_aln_result) * (size_t)nSlots);
    MB200_UP(B[B_VIEWS].ptr, contigSeq, sizeof(mb200_seqview) * (size_t)nSlots);
    MB200_UP(B[B_CONS].ptr, consRange, 8 * (size_t)nSlots);
    MB200_UP(B[B_SLOTLOCUS].ptr, slotLocus.data(), 4 * (size_t)nSlots);
  }
  MB200_UP(B[B_REF].ptr, ref1, sizeof(mb200_seqview) * (size_t)nLoci);
  if (cuts) MB200_UP(B[B_CUTS].ptr, cuts, sizeof(mb200_ref_cuts) * (size_t)nLoci);
  MB200_UP(B[B_SLOTFIRST].ptr, slotFirst, 4 * (size_t)nLoci);
  MB200_UP(B[B_NCONTIGS].ptr, nContigs, 4 * (size_t)nLoci);
  MB200_UP(B[B_ISJUMP].ptr, isJump, (size_t)nLoci);
  uint32_t*       dSvFirst = reinterpret_cast<uint32_t*>(B[B_SV].ptr);
  mb200_small_sv* dSvs     = reinterpret_cast<mb200_small_sv*>(reinterpret_cast<uint8_t*>(B[B_SV].ptr) + ((4 * (size_t)(nLoci + 1) + 63) & ~(size_t)63));
  if (svFirst && svs) {
    MB200_UP(dSvFirst, svFirst, 4 * (size_t)(nLoci + 1));
    if (nSv) MB200_UP(dSvs, svs, sizeof(mb200_small_sv) * nSv);
  }
#undef MB200_UP
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dMisc.ptr, 0, 256, st));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(B[B_FAKE].ptr, 0, MB200_POOL_FRONT, st));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], st));

  uint8_t* ob = reinterpret_cast<uint8_t*>(B[B_OUT].ptr);
  LiParams P;
  memset(&P, 0, sizeof(P));
  P.edge = *edgeScores; P.complete = *completeScores;
  P.nLoci = nLoci; P.nSlots = nSlots;
  P.isJump    = reinterpret_cast<const uint8_t*>(B[B_ISJUMP].ptr);
  P.slotFirst = reinterpret_cast<const uint32_t*>(B[B_SLOTFIRST].ptr);
  P.nContigs  = reinterpret_cast<const uint32_t*>(B[B_NCONTIGS].ptr);
  P.slotLocus = reinterpret_cast<const uint32_t*>(B[B_SLOTLOCUS].ptr);
  P.aln       = reinterpret_cast<const mb200_aln_result*>(B[B_ALN].ptr);
  P.cigar     = reinterpret_cast<const uint32_t*>(B[B_CIGAR].ptr);
  P.pool      = dPoolBase + MB200_POOL_FRONT;
  P.contigSeq = reinterpret_cast<const mb200_seqview*>(B[B_VIEWS].ptr);
  P.consRange = reinterpret_cast<const int32_t*>(B[B_CONS].ptr);
  P.ref1      = reinterpret_cast<const mb200_seqview*>(B[B_REF].ptr);
  P.cuts      = cuts ? reinterpret_cast<const mb200_ref_cuts*>(B[B_CUTS].ptr) : nullptr;
  P.svFirst   = (svFirst && svs) ? dSvFirst : nullptr;
  P.svs       = dSvs;
  P.info      = reinterpret_cast<mb200_large_insert_info*>(ob);
  P.out       = reinterpret_cast<mb200_large_insertion*>(ob + sizeof(mb200_large_insert_info) * ns);
  P.fakeView  = reinterpret_cast<mb200_seqview*>(ob + sizeof(mb200_large_insert_info) * ns + sizeof(mb200_large_insertion) * nLoci);
  mb200_aln_result* dFakeAln = reinterpret_cast<mb200_aln_result*>(reinterpret_cast<uint8_t*>(P.fakeView) + sizeof(mb200_seqview) * nLoci);
  P.fakeAln   = dFakeAln;
  P.fakePool  = reinterpret_cast<uint8_t*>(B[B_FAKE].ptr) + MB200_POOL_FRONT;
  P.fakeCap   = fakeCap;
  uint32_t* misc = reinterpret_cast<uint32_t*>(ctx->dMisc.ptr);
  P.fakeUsed  = reinterpret_cast<unsigned long long*>(misc + 16);
  P.fakeCigar = reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr);

  if (nSlots) liInfoKernel<<<(nSlots + 127) / 128, 128, 0, st>>>(P);
  liPairKernel<<<(nLoci + 3) / 4, 128, 0, st>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches += 2;
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(dFakeAln, 0, sizeof(mb200_aln_result) * (size_t)nLoci, st));
  // the joined contigs' lengths come back to plan the DP (the strips are chosen from the query lengths)
  std::vector<mb200_seqview> hFake(nLoci);
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hFake.data(), P.fakeView, sizeof(mb200_seqview) * (size_t)nLoci, cudaMemcpyDeviceToHost, st));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  std::vector<DpTask> tasks;
  for (uint32_t li = 0; li < nLoci; ++li) {
    if (hFake[li].len == 0) continue;
    const int32_t lead = cuts ? cuts[li].leadingCut : 0, trail = cuts ? cuts[li].trailingCut : 0;
    DpTask        t;
    t.q  = P.fakePool + hFake[li].off;
    t.r1 = P.pool + ref1[li].off + lead;
    t.r2 = t.r1;
    t.Q  = hFake[li].len;
    t.R1 = ref1[li].len - (uint32_t)(lead + trail);
    t.R2 = 0;
    t.outIndex = li;
    tasks.push_back(t);
  }
  uint64_t used = 0;
  if (!tasks.empty()) {
    rc = mb200_dp_run(ctx, MB200_KIND_GLOBAL, completeScores, 0, tasks, dFakeAln, reinterpret_cast<uint32_t*>(ctx->dCigar.ptr), outCigarCap,
                      reinterpret_cast<unsigned long long*>(misc + 8), reinterpret_cast<int*>(misc + 4), misc + 32, &ctx->timing.cells,
                      &ctx->timing.algoBytes);
    if (rc != MB200_OK) return rc;
    uint32_t hMisc[16];
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hMisc, ctx->dMisc.ptr, sizeof(hMisc), cudaMemcpyDeviceToHost, st));
    MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    unsigned long long u;
    memcpy(&u, hMisc + 8, sizeof(u));
    used = u;
    if (hMisc[4] != 0) return mb200_fail(ctx, (int)hMisc[4], "device-side failure in the joined-contig alignment");
    *outCigarUsed = used;
    if (used > outCigarCap) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "output cigar pool too small");
    liFinishKernel<<<(nLoci + 3) / 4, 128, 0, st>>>(P);
    MB200_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->timing.launches++;
  }
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], st));
  if (nSlots) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(info, P.info, sizeof(mb200_large_insert_info) * (size_t)nSlots, cudaMemcpyDeviceToHost, st));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(out, P.out, sizeof(mb200_large_insertion) * (size_t)nLoci, cudaMemcpyDeviceToHost, st));
  if (used) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(outCigar, ctx->dCigar.ptr, 4 * (size_t)used, cudaMemcpyDeviceToHost, st));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], st));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  (void)maxFake;
  return MB200_OK;
}
