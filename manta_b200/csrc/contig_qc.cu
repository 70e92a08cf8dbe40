// contig_qc.cu -- the per-contig steps around the refiner's DP:
//
//   refTrimKernel     10-mer anchored trim of a complex candidate's reference window
//                     (GSC/SVCandidateAssemblyRefiner.cpp:1984-2011)
//   contigQcKernel    isLowQualitySmallSVAlignment (:318), findCandidateVariantsFromComplexSVContigAlignment (:430),
//                     isJumpAlignmentQCFail (:1266), isLowQualityJumpAlignment (:1287)
//   locusSelectKernel the rank1/rank2 rule of getSmallSVAssembly (:2119-2227) and selectJumpContigDNA (:1364-1398)
//
// All three are small next to the DP (a few hundred path segments and < 1e5 byte compares per contig); they exist on
// the device so that the fused candidate path hands back bestAlignmentIndex without a host pass over every CIGAR.
// One warp owns one contig.  The sequential path logic runs redundantly in every lane (identical control flow on
// shared-memory broadcasts, no divergence); the byte scans (10-mer probes, the ambiguity search) are split across
// lanes.
#include <cstdint>
#include <vector>

#include "mb200_internal.h"

namespace {

constexpr int      QC_WARPS   = 4;            // warps per CTA
constexpr int      TRIM_TAB   = 4096;         // shared-memory hash slots per warp (contigs up to 2048 bases)
constexpr int      QC_PATH    = 256;          // path segments staged in shared memory per warp
constexpr uint32_t EMPTY16    = 0xFFFFu;
constexpr int      MER        = 10;

__device__ __forceinline__ uint32_t merHash(const uint8_t* p)
{
  uint32_t h = 2166136261u;
#pragma unroll
  for (int i = 0; i < MER; ++i) h = (h ^ p[i]) * 16777619u;
  return h;
}

__device__ __forceinline__ bool merEqual(const uint8_t* a, const uint8_t* b)
{
  bool eq = true;
#pragma unroll
  for (int i = 0; i < MER; ++i) eq &= (a[i] == b[i]);
  return eq;
}

__device__ __forceinline__ bool merPresent(const unsigned short* tab, uint32_t mask, const uint8_t* contig, const uint8_t* q)
{
  uint32_t h = merHash(q) & mask;
  for (;;) {
    const uint32_t p = tab[h];
    if (p == EMPTY16) return false;
    if (merEqual(contig + p, q)) return true;
    h = (h + 1) & mask;
  }
}

struct TrimParams {
  uint32_t              n;         // items (slots)
  uint32_t              perLocus;  // >0: item i belongs to locus i / perLocus and ref/cuts/isJump are per locus
  const uint8_t*        contigPool;
  const mb200_seqview*  contigView;  // one of contigView / contigRec
  const mb200_contig*   contigRec;
  const uint8_t*        refPool;
  const mb200_seqview*  refView;
  const mb200_ref_cuts* cuts;
  const uint8_t*        isJump;     // nullable; jump loci are skipped
  unsigned short*       bigTab;     // nullable: per-warp global tables for contigs > TRIM_TAB/2
  uint32_t              bigTabSlots;
  mb200_trim_result*    out;
  int*                  errFlag;
};

__global__ void __launch_bounds__(QC_WARPS * 32) refTrimKernel(const TrimParams P)
{
  __shared__ unsigned short sTab[QC_WARPS][TRIM_TAB];
  const int      lane  = threadIdx.x & 31;
  const int      warp  = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * QC_WARPS + warp;
  const uint32_t nwarp = gridDim.x * QC_WARPS;
  for (uint32_t it = gwarp; it < P.n; it += nwarp) {
    const uint32_t li = P.perLocus ? it / P.perLocus : it;
    if (P.isJump && P.isJump[li]) continue;
    uint64_t coff;
    uint32_t cl;
    if (P.contigRec) { coff = P.contigRec[it].seqOff; cl = P.contigRec[it].seqLen; }
    else { coff = P.contigView[it].off; cl = P.contigView[it].len; }
    if (cl == 0 && P.perLocus) continue;  // unused slot of the fused layout
    if (cl < MER || cl >= EMPTY16) {
      if (lane == 0) atomicCAS(P.errFlag, 0, cl < MER ? MB200_ERR_BAD_ARG : MB200_ERR_UNSUPPORTED);
      continue;
    }
    const uint8_t* cs = P.contigPool + coff;
    const uint8_t* rs = P.refPool + P.refView[li].off;
    unsigned short* tab;
    uint32_t        slots;
    if (2 * cl <= TRIM_TAB) { tab = sTab[warp]; slots = TRIM_TAB; }
    else {
      if (!P.bigTab || 2 * cl > P.bigTabSlots) {
        if (lane == 0) atomicCAS(P.errFlag, 0, MB200_ERR_UNSUPPORTED);
        continue;
      }
      tab   = P.bigTab + (size_t)gwarp * P.bigTabSlots;
      slots = P.bigTabSlots;
    }
    // smaller tables for short contigs keep the clear cheap
    while (slots > 64 && slots >= 8 * cl) slots >>= 1;
    const uint32_t mask = slots - 1;
    for (uint32_t i = lane; i < slots; i += 32) tab[i] = (unsigned short)EMPTY16;
    __syncwarp();
    // set of the contig's 10-mers
    for (uint32_t p = lane; p + MER <= cl; p += 32) {
      uint32_t h = merHash(cs + p) & mask;
      for (;;) {
        const unsigned short old = atomicCAS(&tab[h], (unsigned short)EMPTY16, (unsigned short)p);
        if (old == EMPTY16) break;
        if (merEqual(cs + old, cs + p)) break;
        h = (h + 1) & mask;
      }
    }
    __syncwarp();
    const mb200_ref_cuts c          = P.cuts[li];
    const int32_t        refSize    = (int32_t)P.refView[li].len;
    const int32_t        minRef     = c.leadingCut;
    const int32_t        maxRef     = refSize - (c.trailingCut + MER);
    const int32_t        maxFwd     = min(c.maxLeadingCut, maxRef);
    int32_t              lead       = max(minRef, maxFwd + 1);  // value of the loop variable when nothing matches
    for (int32_t base = minRef; base <= maxFwd; base += 32) {
      const int32_t  ri  = base + lane;
      const bool     hit = (ri <= maxFwd) && merPresent(tab, mask, cs, rs + ri);
      const uint32_t b   = __ballot_sync(0xffffffffu, hit);
      if (b) { lead = base + (__ffs(b) - 1); break; }
    }
    const int32_t minRev = max(minRef, refSize - c.maxTrailingCut);
    int32_t       last   = min(maxRef, minRev - 1);
    for (int32_t base = maxRef; base >= minRev; base -= 32) {
      const int32_t  ri  = base - lane;
      const bool     hit = (ri >= minRev) && merPresent(tab, mask, cs, rs + ri);
      const uint32_t b   = __ballot_sync(0xffffffffu, hit);
      if (b) { last = base - (__ffs(b) - 1); break; }
    }
    if (lane == 0) {
      P.out[it].adjLeadingCut  = lead;
      P.out[it].adjTrailingCut = refSize - (last + MER);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// alignment-path helpers (lib/blt_util/align_path.hpp:89-115)
__device__ __forceinline__ bool segIsRead(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH;
}
__device__ __forceinline__ bool segIsRef(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH;
}
__device__ __forceinline__ bool segIsIndel(uint32_t t) { return t == MB200_SEG_INSERT || t == MB200_SEG_DELETE; }

struct Arm {
  uint32_t refSize, readSize, clipRight;
  int32_t  score;
};

// the sub-path s[0..n) oriented away from the breakpoint, cut after maxSpan reference bases
// (apath_limit_ref_length, align_path.cpp:273), measured like apath_ref_length / apath_read_length /
// apath_soft_clip_right_size / getPathScore(..., isScoreOffEdge=false) (AlignmentScoringUtilImpl.hpp:35-92)
__device__ Arm armMeasure(const uint32_t* s, uint32_t n, bool reversed, uint32_t maxSpan, const mb200_scores& sc)
{
  Arm      a      = {0, 0, 0, 0};
  uint32_t refLen = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t seg  = reversed ? s[n - 1 - i] : s[i];
    const uint32_t t    = MB200_SEG_TYPE(seg);
    uint32_t       len  = MB200_SEG_LEN(seg);
    bool           stop = false;
    if (segIsRef(t)) {
      refLen += len;
      if (refLen >= maxSpan) {
        len -= (refLen - maxSpan);
        stop = true;
      }
      a.refSize += len;
    }
    if (segIsRead(t)) a.readSize += len;
    if (t == MB200_SEG_SOFT_CLIP) a.clipRight += len;
    else if (t != MB200_SEG_HARD_CLIP) a.clipRight = 0;
    if (t == MB200_SEG_SEQ_MATCH) a.score += sc.match * (int32_t)len;
    else if (t == MB200_SEG_SEQ_MISMATCH) a.score += sc.mismatch * (int32_t)len;
    else if (segIsIndel(t)) a.score += sc.open + sc.extend * (int32_t)len;
    if (stop) break;
  }
  return a;
}

__device__ __forceinline__ bool lowScoreFraction(const Arm& a, uint32_t clipped, const mb200_scores& sc)
{
  const int32_t nonClip = max(0, a.score);
  const int32_t optimal = (int32_t)clipped * sc.match;
  const float   frac    = __fdiv_rn((float)nonClip, (float)optimal);
  return frac < 0.75f;
}

__device__ bool lowQualitySmall(const uint32_t* s, uint32_t n, bool leading, bool isComplex, uint32_t maxSpan, const mb200_scores& sc)
{
  const uint32_t minLen = isComplex ? 35u : 30u;  // minAlignRefSpan* == minAlignReadLength*  (:326-335)
  const Arm      a      = armMeasure(s, n, leading, maxSpan, sc);
  if (a.refSize < minLen) return true;
  const uint32_t clipped = a.readSize - a.clipRight;
  if (clipped < minLen) return true;
  return lowScoreFraction(a, clipped, sc);
}

__device__ bool lowQualitySpanning(const uint32_t* s, uint32_t n, bool leading, uint32_t maxSpan, const mb200_scores& sc)
{
  const Arm      a       = armMeasure(s, n, leading, maxSpan, sc);
  const uint32_t clipped = a.readSize - a.clipRight;
  if (clipped < 30u) return true;
  return lowScoreFraction(a, clipped, sc);
}

// next / previous candidate indel run (getLargeIndelSegments :173): a maximal run of insert/delete segments holding
// at least one of length >= minSize.  Forward: first run starting at or after `from`; backward: last run ending
// before `before`.
__device__ bool nextRun(const uint32_t* s, uint32_t n, uint32_t from, uint32_t minSize, uint32_t& first, uint32_t& last)
{
  uint32_t i = from;
  while (i < n) {
    if (!segIsIndel(MB200_SEG_TYPE(s[i]))) { ++i; continue; }
    const uint32_t st = i;
    bool           cand = false;
    while (i < n && segIsIndel(MB200_SEG_TYPE(s[i]))) { cand |= (MB200_SEG_LEN(s[i]) >= minSize); ++i; }
    if (cand) { first = st; last = i - 1; return true; }
  }
  return false;
}

__device__ bool prevRun(const uint32_t* s, uint32_t before, uint32_t minSize, uint32_t& first, uint32_t& last)
{
  int32_t i = (int32_t)before - 1;
  while (i >= 0) {
    if (!segIsIndel(MB200_SEG_TYPE(s[i]))) { --i; continue; }
    const int32_t en = i;
    bool          cand = false;
    while (i >= 0 && segIsIndel(MB200_SEG_TYPE(s[i]))) { cand |= (MB200_SEG_LEN(s[i]) >= minSize); --i; }
    if (cand) { first = (uint32_t)(i + 1); last = (uint32_t)en; return true; }
  }
  return false;
}

// getQuerySeqMatchCount (:393-418) > 1 ?   Lanes take scan offsets; an offset is abandoned as soon as its mismatch
// count passes the largest count that still satisfies float(mm)/float(querySize) <= 0.05f.
__device__ bool ambiguousFlank(const uint8_t* target, uint32_t targetSize, const uint8_t* query, uint32_t querySize, int lane)
{
  if (querySize == 0 || querySize > targetSize) return false;  // 0/0 is NaN: no offset ever matches
  uint32_t maxMm = querySize / 20 + 2;
  while (maxMm > 0 && !(__fdiv_rn((float)maxMm, (float)querySize) <= 0.05f)) --maxMm;
  const uint32_t scanEnd = targetSize - querySize;
  uint32_t       count   = 0;
  for (uint32_t base = 0; base <= scanEnd; base += 32) {
    const uint32_t i  = base + lane;
    bool           ok = (i <= scanEnd);
    if (ok) {
      uint32_t mm = 0;
      for (uint32_t j = 0; j < querySize; ++j) {
        const uint8_t q = query[j];
        mm += (q != target[i + j] || q == 'N') ? 1u : 0u;
        if (mm > maxMm) break;
      }
      ok = (mm <= maxMm);
    }
    count += __popc(__ballot_sync(0xffffffffu, ok));
    if (count > 1) return true;
  }
  return false;
}

struct QcParams {
  mb200_qc_opts            opts;
  uint32_t                 nLoci;
  const uint8_t*           isJump;
  const uint32_t*          slotFirst;
  const uint32_t*          nContigs;
  const mb200_aln_result*  aln;
  const mb200_jump_result* jump;
  const uint32_t*          cigar;
  const uint8_t*           contigPool;
  const mb200_seqview*     contigView;
  const uint8_t*           refPool;
  const mb200_seqview*     ref1;
  const uint32_t*          support;
  const uint32_t*          slotLocus;  // locus of each slot
  uint32_t                 nSlots;
  mb200_contig_qc*         qc;
  int32_t*                 best;
  // fused path: skip everything when the DP stage overflowed its cigar pool or failed (its results are unusable
  // and the caller is about to get that error)
  const unsigned long long* cigarUsed;
  unsigned long long        cigarCap;
  const int*                errFlag;
};

__device__ __forceinline__ bool upstreamFailed(const QcParams& P)
{
  return (P.cigarUsed && *P.cigarUsed > P.cigarCap) || (P.errFlag && *P.errFlag != 0);
}

// findCandidateVariantsFromComplexSVContigAlignment (:430-553) for one maxQCRefSpan; kept segments = ordinals
// [ordLo, ordHi] covering path[keptFirst .. keptLast]
__device__ bool findCandidates(const uint32_t* path, uint32_t n, int32_t beginPos, const uint8_t* contig, uint32_t contigLen,
                               const uint8_t* ref, uint32_t refLen, uint32_t nRuns, uint32_t f0, uint32_t l0, uint32_t fL,
                               uint32_t lL, bool isComplex, uint32_t maxSpan, const mb200_qc_opts& o, int lane,
                               uint32_t& ordLo, uint32_t& ordHi, uint32_t& keptFirst, uint32_t& keptLast)
{
  const uint32_t minSize = o.minCandidateVariantSize;
  uint32_t       lo = 0, hi = nRuns - 1;
  uint32_t       loFirst = f0, loLast = l0, hiFirst = fL, hiLast = lL;
  for (;;) {
    if (!lowQualitySmall(path, loFirst, true, isComplex, maxSpan, o.filterScores)) break;
    if (hi == lo) return false;
    nextRun(path, n, loLast + 1, minSize, loFirst, loLast);
    ++lo;
  }
  for (;;) {
    if (!lowQualitySmall(path + hiLast + 1, n - (hiLast + 1), false, isComplex, maxSpan, o.filterScores)) break;
    if (hi == lo) return false;
    prevRun(path, hiFirst, minSize, hiFirst, hiLast);
    --hi;
  }
  uint32_t leftSize = 0, endPos = 0, refSpan = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t t = MB200_SEG_TYPE(path[i]), len = MB200_SEG_LEN(path[i]);
    if (segIsRead(t)) {
      if (i < loFirst) leftSize += len;
      if (i <= hiLast) endPos += len;
    }
    if (segIsRef(t)) refSpan += len;
  }
  // std::string::substr clamps the count to the end of the string
  const uint32_t leftQ         = min(leftSize, contigLen);
  const int32_t  rightSize     = max(0, (int32_t)contigLen - (int32_t)endPos);
  const int32_t  refAlignStart = beginPos;
  const int32_t  refAlignEnd   = beginPos + (int32_t)refSpan;
  const int32_t  leftStart     = max(0, refAlignEnd - 500);
  uint32_t       tl            = (uint32_t)(refAlignEnd - leftStart);
  if ((uint32_t)leftStart + tl > refLen) tl = refLen - (uint32_t)leftStart;
  if (ambiguousFlank(ref + leftStart, tl, contig, leftQ, lane)) return false;
  const int32_t rs = max(0, min(500, (int32_t)refLen - refAlignStart));
  if (ambiguousFlank(ref + refAlignStart, (uint32_t)rs, contig + endPos, (uint32_t)rightSize, lane)) return false;
  ordLo = lo; ordHi = hi; keptFirst = loFirst; keptLast = hiLast;
  return true;
}

__global__ void __launch_bounds__(QC_WARPS * 32) contigQcKernel(const QcParams P)
{
  __shared__ uint32_t sPath[QC_WARPS][2 * QC_PATH];
  const int      lane  = threadIdx.x & 31;
  const int      warp  = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * QC_WARPS + warp;
  const uint32_t nwarp = gridDim.x * QC_WARPS;
  if (upstreamFailed(P)) return;
  for (uint32_t slot = gwarp; slot < P.nSlots; slot += nwarp) {
    const uint32_t  li = P.slotLocus[slot];
    mb200_contig_qc q  = {0, 0, 0, 0, 0, 0, 0, 0};
    __syncwarp();
    if (li == 0xFFFFFFFFu) {  // unused slot
      if (lane == 0) P.qc[slot] = q;
      continue;
    }
    if (P.isJump[li]) {
      const mb200_jump_result r = P.jump[slot];
      const uint32_t*         p1 = P.cigar + r.cigarOff1;
      const uint32_t*         p2 = P.cigar + r.cigarOff2;
      if (r.cigarLen1 <= QC_PATH && r.cigarLen2 <= QC_PATH) {
        for (uint32_t i = lane; i < r.cigarLen1; i += 32) sPath[warp][i] = p1[i];
        for (uint32_t i = lane; i < r.cigarLen2; i += 32) sPath[warp][QC_PATH + i] = p2[i];
        __syncwarp();
        p1 = sPath[warp];
        p2 = sPath[warp] + QC_PATH;
      }
      uint32_t ref1 = 0, ref2 = 0;
      for (uint32_t i = 0; i < r.cigarLen1; ++i)
        if (segIsRef(MB200_SEG_TYPE(p1[i]))) ref1 += MB200_SEG_LEN(p1[i]);
      for (uint32_t i = 0; i < r.cigarLen2; ++i)
        if (segIsRef(MB200_SEG_TYPE(p2[i]))) ref2 += MB200_SEG_LEN(p2[i]);
      // isJumpAlignmentSegmentQCFail (:1254): not aligned, or fewer than 20 reference bases
      const bool fail = (r.cigarLen1 == 0) || (ref1 < 20) || (r.cigarLen2 == 0) || (ref2 < 20);
      bool       low1 = true, low2 = true;
      const uint32_t spans[3] = {75, 100, 200};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (!lowQualitySpanning(p1, r.cigarLen1, true, spans[k], P.opts.filterScores)) low1 = false;
        if (!lowQualitySpanning(p2, r.cigarLen2, false, spans[k], P.opts.filterScores)) low2 = false;
      }
      q.isCandidate  = fail ? 0u : 1u;
      q.isLowQuality = (low1 || low2) ? 1u : 0u;
    } else {
      const mb200_aln_result r    = P.aln[slot];
      const uint32_t*        path = P.cigar + r.cigarOff;
      const uint32_t         n    = r.cigarLen;
      if (n <= 2 * QC_PATH) {
        for (uint32_t i = lane; i < n; i += 32) sPath[warp][i] = path[i];
        __syncwarp();
        path = sPath[warp];
      }
      const uint32_t minSize = P.opts.minCandidateVariantSize;
      uint32_t       nRuns = 0, f0 = 0, l0 = 0, fL = 0, lL = 0, f = 0, l = 0, from = 0;
      while (nextRun(path, n, from, minSize, f, l)) {
        if (nRuns == 0) { f0 = f; l0 = l; }
        fL = f; lL = l;
        ++nRuns;
        from = l + 1;
      }
      if (nRuns) {
        const bool     isComplex = (nRuns > 1) || (f0 != l0);
        const uint8_t* contig    = P.contigPool + P.contigView[slot].off;
        const uint32_t contigLen = P.contigView[slot].len;
        const uint8_t* ref       = P.refPool + P.ref1[li].off;
        const uint32_t refLen    = P.ref1[li].len;
        uint32_t       bestCnt   = 0;
        const uint32_t spans[2]  = {100, 200};
        for (int k = 0; k < 2; ++k) {
          uint32_t a, b, kf, kl;
          if (findCandidates(path, n, r.beginPos, contig, contigLen, ref, refLen, nRuns, f0, l0, fL, lL, isComplex, spans[k],
                             P.opts, lane, a, b, kf, kl)) {
            if (b - a + 1 > bestCnt) {
              bestCnt = b - a + 1;
              q.segOrdinal = a; q.nSegments = bestCnt; q.pathFirst = kf; q.pathLast = kl;
            }
            q.isCandidate = 1;
          }
        }
        if (q.isCandidate) {
          // getLargestIndelSize (:210): runs without a qualifying indel inside [pathFirst, pathLast] are all shorter
          // than minSize, so the maximum over every indel of the range is the maximum over the kept segments
          uint32_t big = 0;
          for (uint32_t i = q.pathFirst; i <= q.pathLast; ++i)
            if (segIsIndel(MB200_SEG_TYPE(path[i]))) big = max(big, MB200_SEG_LEN(path[i]));
          q.variantSize = big;
        }
      }
    }
    if (lane == 0) P.qc[slot] = q;
  }
}

__global__ void locusSelectKernel(const QcParams P)
{
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= P.nLoci || upstreamFailed(P)) return;
  const uint32_t         s0 = P.slotFirst[li];
  const uint32_t         nc = P.nContigs[li];
  const mb200_contig_qc* qc = P.qc + s0;
  int32_t                best = -1;
  if (P.isJump[li]) {  // selectJumpContigDNA (:1364-1398)
    const mb200_jump_result* jr = P.jump + s0;
    for (uint32_t c = 0; c < nc; ++c) {
      if (!qc[c].isCandidate) continue;
      if (best == -1 || jr[c].score > jr[best].score) best = (int32_t)c;
    }
    if (best != -1 && qc[best].isLowQuality) best = -1;
  } else {  // rank1/rank2 rule (:2119-2227)
    const mb200_aln_result* ar = P.aln + s0;
    bool                    def1 = false, def2 = false, jmp1 = false, jmp2 = false;
    uint32_t                idx1 = 0, idx2 = 0, var1 = 0, var2 = 0;
    int32_t                 sc1 = 0, sc2 = 0;
    for (uint32_t c = 0; c < nc; ++c) {
      if (!qc[c].isCandidate) continue;
      const bool    jumped = ar[c].isJumped != 0;
      const int32_t score  = ar[c].score;
      const bool    same   = (jumped == jmp1);
      if ((!def1) || (jumped && !jmp1) || (same && score > sc1)) {
        if (def1) { def2 = true; idx2 = idx1; sc2 = sc1; var2 = var1; jmp2 = jmp1; }
        def1 = true; idx1 = c; sc1 = score; var1 = qc[c].variantSize; jmp1 = jumped;
      } else if ((!def2) || score > sc2) {
        def2 = true; idx2 = c; sc2 = score; var2 = qc[c].variantSize; jmp2 = jumped;
      }
    }
    if (def2 && !(jmp1 && !jmp2)) {
      const uint32_t n1 = P.support[s0 + idx1], n2 = P.support[s0 + idx2];
      const bool     rank2Best = ((float)sc2 > __fmul_rn((float)sc1, 0.9f)) &&
                             (((float)n2 > __fmul_rn((float)n1, 1.2f)) || ((float)var2 > __fmul_rn((float)var1, 1.1f)));
      if (rank2Best) idx1 = idx2;
    }
    best = def1 ? (int32_t)idx1 : -1;
  }
  P.best[li] = best;
}

// supportReads.size() of every slot of the fused layout, and the slot -> locus map
__global__ void slotPrepKernel(uint32_t nLoci, uint32_t A, const mb200_asm_locus* loci, const mb200_contig* contigs,
                               const uint32_t* setPool, uint32_t* support, uint32_t* slotLocus, uint32_t* slotFirst,
                               uint32_t* nContigs, mb200_seqview* contigView)
{
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nLoci * A) return;
  const uint32_t li = slot / A, c = slot % A;
  const bool     live = c < loci[li].nContigs;
  uint32_t       cnt  = 0;
  if (live) {
    const uint32_t* w = setPool + contigs[slot].supportOff;
    for (uint32_t i = 0; i < loci[li].setWords; ++i) cnt += __popc(w[i]);
  }
  support[slot]        = cnt;
  slotLocus[slot]      = live ? li : 0xFFFFFFFFu;
  contigView[slot].off = live ? contigs[slot].seqOff : 0;
  contigView[slot].len = live ? contigs[slot].seqLen : 0;
  contigView[slot].reserved = 0;
  if (c == 0) {
    slotFirst[li] = li * A;
    nContigs[li]  = loci[li].nContigs;
  }
}

// AlignmentResult::align.beginPos += adjustedLeadingCut (:2038) for every live complex slot of the fused layout
__global__ void applyTrimKernel(uint32_t nSlots, const uint32_t* slotLocus, const uint8_t* isJump, const mb200_trim_result* trim,
                                mb200_aln_result* aln)
{
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nSlots) return;
  const uint32_t li = slotLocus[slot];
  if (li == 0xFFFFFFFFu || isJump[li]) return;
  aln[slot].beginPos += trim[slot].adjLeadingCut;
}


// ---------------------------------------------------------------------------------------------------------------------
// refined small SVs: setSmallCandSV (:720-800) over the kept segments of the selected contig (:2243-2251)
struct SvParams {
  uint32_t                nLoci, minSize;
  const uint8_t*          isJump;
  const uint32_t*         slotFirst;
  const int32_t*          best;
  const mb200_contig_qc*  qc;
  const mb200_aln_result* aln;
  const uint32_t*         cigar;
  const uint8_t*          contigPool;
  const mb200_seqview*    contigView;
  const uint8_t*          refPool;
  const mb200_seqview*    ref1;
  uint32_t*               svFirst;  // [nLoci + 1]
  mb200_small_sv*         svs;
  unsigned long long      svCap;
};

// svFirst = exclusive prefix sum of the per-candidate record counts (one CTA; the counts are tiny)
__global__ void __launch_bounds__(1024) svOffsetsKernel(const SvParams P)
{
  __shared__ uint32_t sSum[1024];
  const uint32_t      tid = threadIdx.x, nt = blockDim.x;
  const uint32_t      per = (P.nLoci + nt - 1) / nt;
  const uint32_t      lo = min(P.nLoci, tid * per), hi = min(P.nLoci, lo + per);
  auto count = [&](uint32_t li) -> uint32_t {
    if (P.isJump[li] || P.best[li] < 0) return 0u;
    return P.qc[P.slotFirst[li] + (uint32_t)P.best[li]].nSegments;
  };
  uint32_t mine = 0;
  for (uint32_t li = lo; li < hi; ++li) mine += count(li);
  sSum[tid] = mine;
  __syncthreads();
  for (uint32_t d = 1; d < nt; d <<= 1) {
    const uint32_t v = (tid >= d) ? sSum[tid - d] : 0u;
    __syncthreads();
    sSum[tid] += v;
    __syncthreads();
  }
  uint32_t run = sSum[tid] - mine;
  for (uint32_t li = lo; li < hi; ++li) {
    P.svFirst[li] = run;
    run += count(li);
  }
  if (tid == nt - 1) P.svFirst[P.nLoci] = sSum[tid];
}

__global__ void __launch_bounds__(QC_WARPS * 32) smallSvKernel(const SvParams P)
{
  const int      lane  = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * QC_WARPS + (threadIdx.x >> 5);
  const uint32_t nwarp = gridDim.x * QC_WARPS;
  for (uint32_t li = gwarp; li < P.nLoci; li += nwarp) {
    if (P.isJump[li] || P.best[li] < 0) continue;
    const uint32_t         s    = P.slotFirst[li] + (uint32_t)P.best[li];
    const mb200_contig_qc  q    = P.qc[s];
    const mb200_aln_result r    = P.aln[s];
    const uint32_t*        path = P.cigar + r.cigarOff;
    const uint32_t         n    = r.cigarLen;
    const uint8_t*         read = P.contigPool + P.contigView[s].off;
    const int64_t          readLen = P.contigView[s].len;
    const uint8_t*         ref  = P.refPool + P.ref1[li].off;
    const int64_t          refLen = P.ref1[li].len;
    // skip to the first kept candidate run
    uint32_t f = 0, l = 0, from = 0;
    bool     ok = true;
    for (uint32_t k = 0; k <= q.segOrdinal && ok; ++k) {
      ok   = nextRun(path, n, from, P.minSize, f, l);
      from = l + 1;
    }
    for (uint32_t k = 0; k < q.nSegments && ok; ++k) {
      mb200_small_sv v;
      memset(&v, 0, sizeof(v));
      v.locus = li; v.contig = (uint32_t)P.best[li]; v.segment = k; v.pathFirst = f; v.pathLast = l;
      int32_t readPos = 0, refPos = r.beginPos;
      for (uint32_t i = 0; i < n; ++i) {
        const uint32_t t = MB200_SEG_TYPE(path[i]), len = MB200_SEG_LEN(path[i]);
        if (i == f) { v.refBegin = refPos; v.readBegin = readPos; }
        if (segIsRef(t)) refPos += (int32_t)len;
        if (segIsRead(t)) readPos += (int32_t)len;
        if (i == l) { v.refEnd = refPos; v.readEnd = readPos; }
      }
      // getVariantRange (:677-717): lanes compare 32 positions per round
      const int64_t maxRight = min(refLen - v.refEnd, readLen - v.readEnd);
      int32_t       right = (int32_t)max((int64_t)0, maxRight);
      for (int64_t base = 0; base < maxRight; base += 32) {
        const int64_t  o    = base + lane;
        const bool     stop = (o >= maxRight) || (ref[v.refBegin + o] != read[v.readBegin + o]);
        const uint32_t b    = __ballot_sync(0xffffffffu, stop);
        if (b) { right = (int32_t)(base + (__ffs(b) - 1)); break; }
      }
      const int32_t minLeft = max(-v.refBegin, -v.readBegin);
      int32_t       left    = 0;
      for (int32_t base = 0;; base += 32) {
        const int32_t j  = base + lane;  // candidate offset -j
        const int32_t ri = v.refEnd - j - 1, qi = v.readEnd - j - 1;
        const bool    stop = (-j < minLeft) || ri < 0 || qi < 0 || (ref[ri] != read[qi]);
        const uint32_t b   = __ballot_sync(0xffffffffu, stop);
        if (b) { left = -(base + (__ffs(b) - 1)); break; }
      }
      v.ciposBegin = left;
      v.ciposEnd   = right;
      if (left != 0) v.status = MB200_ERR_SV_RANGE;
      else {
        v.bp1Begin = v.refBegin - 1;
        v.bp1End   = v.bp1Begin + right + 1;
        v.bp2Begin = v.refEnd;
        v.bp2End   = v.bp2Begin + right + 1;
      }
      const unsigned long long at = (unsigned long long)P.svFirst[li] + k;
      if (lane == 0 && at < P.svCap) P.svs[at] = v;
      if (k + 1 < q.nSegments) {
        ok   = nextRun(path, n, from, P.minSize, f, l);
        from = l + 1;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// refined spanning breakends: generateRefinedSVCandidateFromJumpAlignment (:1190-1223), adjustAssembledBreakend (:59-86)
struct JumpSvParams {
  uint32_t                 nLoci;
  const uint8_t*           isJump;
  const uint32_t*          slotFirst;
  const int32_t*           best;
  const mb200_jump_result* jump;
  const uint32_t*          cigar;
  const mb200_seqview*     ref1;
  const mb200_seqview*     ref2;
  const mb200_jump_orient* orient;
  mb200_jump_sv*           out;
};

__device__ __forceinline__ void adjustBreakend(int32_t beginPos, uint32_t refSpan, bool isAlign1, uint32_t jumpRange, uint32_t refSize,
                                               bool isReversed, bool rightOpen, int32_t& rb, int32_t& re)
{
  const int32_t alignEnd = beginPos + (int32_t)refSpan;
  const int32_t beginOff = isReversed ? ((int32_t)refSize - alignEnd) : beginPos;  // getAlignBeginOffset, AlignmentUtil.hpp:38
  const int32_t endOff   = isReversed ? ((int32_t)refSize - beginPos) : alignEnd;  // getAlignEndOffset :49
  const int32_t bpOff    = rightOpen ? (endOff - 1) : beginOff;
  if (isAlign1 == rightOpen) { rb = bpOff; re = bpOff + (int32_t)jumpRange + 1; }
  else { rb = bpOff - (int32_t)jumpRange; re = bpOff + 1; }
}

__global__ void jumpSvKernel(const JumpSvParams P)
{
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= P.nLoci) return;
  mb200_jump_sv v;
  memset(&v, 0, sizeof(v));
  if (P.isJump[li] && P.best[li] >= 0) {
    const mb200_jump_result r  = P.jump[P.slotFirst[li] + (uint32_t)P.best[li]];
    const mb200_jump_orient o  = P.orient[li];
    const uint32_t*         p1 = P.cigar + r.cigarOff1;
    const uint32_t*         p2 = P.cigar + r.cigarOff2;
    uint32_t                span1 = 0, span2 = 0, read1 = 0;
    for (uint32_t i = 0; i < r.cigarLen1; ++i) {
      const uint32_t t = MB200_SEG_TYPE(p1[i]), len = MB200_SEG_LEN(p1[i]);
      if (segIsRef(t)) span1 += len;
      if (segIsRead(t)) read1 += len;
    }
    for (uint32_t i = 0; i < r.cigarLen2; ++i)
      if (segIsRef(MB200_SEG_TYPE(p2[i]))) span2 += MB200_SEG_LEN(p2[i]);
    v.valid  = 1;
    v.contig = (uint32_t)P.best[li];
    // bp1 takes align2 (and its window) when bp2 was aligned first (:1196-1222, :1544-1547)
    if (!o.isBp2AlignedFirst) {
      adjustBreakend(r.beginPos1, span1, true, r.jumpRange, P.ref1[li].len, o.isBp1Reversed != 0, o.bp1RightOpen != 0, v.bp1Begin, v.bp1End);
      adjustBreakend(r.beginPos2, span2, false, r.jumpRange, P.ref2[li].len, o.isBp2Reversed != 0, o.bp2RightOpen != 0, v.bp2Begin, v.bp2End);
    } else {
      adjustBreakend(r.beginPos2, span2, false, r.jumpRange, P.ref2[li].len, o.isBp1Reversed != 0, o.bp1RightOpen != 0, v.bp1Begin, v.bp1End);
      adjustBreakend(r.beginPos1, span1, true, r.jumpRange, P.ref1[li].len, o.isBp2Reversed != 0, o.bp2RightOpen != 0, v.bp2Begin, v.bp2End);
    }
    v.insertBegin = (int32_t)read1;
    v.insertEnd   = (int32_t)(read1 + r.jumpInsertSize);
  }
  P.out[li] = v;
}

int gridFor(const mb200_ctx* ctx, uint32_t nItems)
{
  const uint32_t want = (nItems + QC_WARPS - 1) / QC_WARPS;
  const uint32_t cap  = (uint32_t)ctx->smCount * 8;
  return (int)std::max(1u, std::min(want, cap));
}

}  // namespace

// ---- device-level entry points (used by candidates.cu) --------------------------------------------------------------
int mb200_trim_run(mb200_ctx* ctx, uint32_t n, uint32_t perLocus, const uint8_t* dContigPool, const mb200_seqview* dContigView,
                   const mb200_contig* dContigRec, const uint8_t* dRefPool, const mb200_seqview* dRefView,
                   const mb200_ref_cuts* dCuts, const uint8_t* dIsJump, uint32_t maxContigLen, mb200_trim_result* dOut,
                   int* dErrFlag)
{
  if (n == 0) return MB200_OK;
  TrimParams P;
  memset(&P, 0, sizeof(P));
  P.n = n; P.perLocus = perLocus; P.contigPool = dContigPool; P.contigView = dContigView; P.contigRec = dContigRec;
  P.refPool = dRefPool; P.refView = dRefView; P.cuts = dCuts; P.isJump = dIsJump; P.out = dOut; P.errFlag = dErrFlag;
  const int grid = gridFor(ctx, n);
  if (2 * (size_t)maxContigLen > (size_t)TRIM_TAB) {
    uint32_t slots = TRIM_TAB;
    while (slots < 2 * maxContigLen) slots <<= 1;
    int rc;
    if ((rc = mb200_reserve(ctx, ctx->dQcScratch, (size_t)grid * QC_WARPS * slots * sizeof(unsigned short))) != MB200_OK) return rc;
    P.bigTab      = reinterpret_cast<unsigned short*>(ctx->dQcScratch.ptr);
    P.bigTabSlots = slots;
  }
  refTrimKernel<<<grid, QC_WARPS * 32, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

int mb200_qc_prep_fused(mb200_ctx* ctx, uint32_t nLoci, uint32_t A, const mb200_asm_locus* dLoci, const mb200_contig* dContigs,
                        const uint32_t* dSetPool)
{
  const size_t nSlots = (size_t)nLoci * A;
  int          rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSupport, 4 * nSlots + 4)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSlotLocus, 4 * nSlots + 4)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSlotFirst, 4 * (size_t)nLoci + 4)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcNContigs, 4 * (size_t)nLoci + 4)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcViews, sizeof(mb200_seqview) * nSlots + 16)) != MB200_OK) return rc;
  slotPrepKernel<<<(unsigned)((nSlots + 255) / 256), 256, 0, ctx->stream>>>(
      nLoci, A, dLoci, dContigs, dSetPool, reinterpret_cast<uint32_t*>(ctx->dQcSupport.ptr),
      reinterpret_cast<uint32_t*>(ctx->dQcSlotLocus.ptr), reinterpret_cast<uint32_t*>(ctx->dQcSlotFirst.ptr),
      reinterpret_cast<uint32_t*>(ctx->dQcNContigs.ptr), reinterpret_cast<mb200_seqview*>(ctx->dQcViews.ptr));
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

int mb200_apply_trim(mb200_ctx* ctx, uint32_t nSlots, const uint32_t* dSlotLocus, const uint8_t* dIsJump,
                     const mb200_trim_result* dTrim, mb200_aln_result* dAln)
{
  if (nSlots == 0) return MB200_OK;
  applyTrimKernel<<<(nSlots + 255) / 256, 256, 0, ctx->stream>>>(nSlots, dSlotLocus, dIsJump, dTrim, dAln);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

int mb200_qc_run(mb200_ctx* ctx, const mb200_qc_opts* opts, uint32_t nLoci, uint32_t nSlots, const uint8_t* dIsJump,
                 const uint32_t* dSlotFirst, const uint32_t* dNContigs, const uint32_t* dSlotLocus, const mb200_aln_result* dAln,
                 const mb200_jump_result* dJump, const uint32_t* dCigar, const uint8_t* dContigPool,
                 const mb200_seqview* dContigView, const uint8_t* dRefPool, const mb200_seqview* dRef1, const uint32_t* dSupport,
                 mb200_contig_qc* dQc, int32_t* dBest, const unsigned long long* dCigarUsed, uint64_t cigarCap, const int* dErrFlag)
{
  if (nLoci == 0) return MB200_OK;
  QcParams P;
  memset(&P, 0, sizeof(P));
  P.opts = *opts; P.nLoci = nLoci; P.isJump = dIsJump; P.slotFirst = dSlotFirst; P.nContigs = dNContigs; P.aln = dAln;
  P.jump = dJump; P.cigar = dCigar; P.contigPool = dContigPool; P.contigView = dContigView; P.refPool = dRefPool; P.ref1 = dRef1;
  P.support = dSupport; P.slotLocus = dSlotLocus; P.nSlots = nSlots; P.qc = dQc; P.best = dBest;
  P.cigarUsed = dCigarUsed; P.cigarCap = cigarCap; P.errFlag = dErrFlag;
  if (nSlots) {
    contigQcKernel<<<gridFor(ctx, nSlots), QC_WARPS * 32, 0, ctx->stream>>>(P);
    MB200_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->timing.launches++;
  }
  locusSelectKernel<<<(nLoci + 127) / 128, 128, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

int mb200_small_sv_run(mb200_ctx* ctx, uint32_t minSize, uint32_t nLoci, const uint8_t* dIsJump, const uint32_t* dSlotFirst,
                       const int32_t* dBest, const mb200_contig_qc* dQc, const mb200_aln_result* dAln, const uint32_t* dCigar,
                       const uint8_t* dContigPool, const mb200_seqview* dContigView, const uint8_t* dRefPool,
                       const mb200_seqview* dRef1, uint32_t* dSvFirst, mb200_small_sv* dSvs, uint64_t svCap)
{
  if (nLoci == 0) return MB200_OK;
  SvParams P;
  memset(&P, 0, sizeof(P));
  P.nLoci = nLoci; P.minSize = minSize; P.isJump = dIsJump; P.slotFirst = dSlotFirst; P.best = dBest; P.qc = dQc; P.aln = dAln;
  P.cigar = dCigar; P.contigPool = dContigPool; P.contigView = dContigView; P.refPool = dRefPool; P.ref1 = dRef1;
  P.svFirst = dSvFirst; P.svs = dSvs; P.svCap = svCap;
  svOffsetsKernel<<<1, 1024, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  smallSvKernel<<<gridFor(ctx, nLoci), QC_WARPS * 32, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches += 2;
  return MB200_OK;
}

int mb200_jump_sv_run(mb200_ctx* ctx, uint32_t nLoci, const uint8_t* dIsJump, const uint32_t* dSlotFirst, const int32_t* dBest,
                      const mb200_jump_result* dJump, const uint32_t* dCigar, const mb200_seqview* dRef1, const mb200_seqview* dRef2,
                      const mb200_jump_orient* dOrient, mb200_jump_sv* dOut)
{
  if (nLoci == 0) return MB200_OK;
  JumpSvParams P;
  memset(&P, 0, sizeof(P));
  P.nLoci = nLoci; P.isJump = dIsJump; P.slotFirst = dSlotFirst; P.best = dBest; P.jump = dJump; P.cigar = dCigar;
  P.ref1 = dRef1; P.ref2 = dRef2; P.orient = dOrient; P.out = dOut;
  jumpSvKernel<<<(nLoci + 127) / 128, 128, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

// ---- host-pointer C ABI ------------------------------------------------------------------------------------------------
extern "C" {

int mb200_ref_trim_batch(mb200_ctx* ctx, uint32_t n, const uint8_t* pool, uint64_t poolBytes, const mb200_seqview* contig,
                         const mb200_seqview* ref, const mb200_ref_cuts* cuts, mb200_trim_result* out)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (n == 0) return MB200_OK;
  if (!pool || !contig || !ref || !cuts || !out) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  uint32_t maxLen = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (contig[i].off + contig[i].len > poolBytes || ref[i].off + ref[i].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "view outside the pool");
    if (contig[i].len < MER) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "contig shorter than the 10-mer used by the window trim");
    if (cuts[i].leadingCut < 0 || cuts[i].trailingCut < 0 || cuts[i].maxLeadingCut < 0 || cuts[i].maxTrailingCut < 0)
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "negative cut");
    maxLen = std::max(maxLen, contig[i].len);
  }
  if (maxLen >= EMPTY16) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "contig longer than 65534 bases");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;  // shares device buffers with the fused path's resident results
  int rc;
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcViews, sizeof(mb200_seqview) * (size_t)n)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcRef, sizeof(mb200_seqview) * (size_t)n)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcCuts, sizeof(mb200_ref_cuts) * (size_t)n)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcTrim, sizeof(mb200_trim_result) * (size_t)n)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dMisc, 256)) != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dMisc.ptr, 0, 256, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dPool.ptr, pool, poolBytes, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dQcViews.ptr, contig, sizeof(mb200_seqview) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dQcRef.ptr, ref, sizeof(mb200_seqview) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dQcCuts.ptr, cuts, sizeof(mb200_ref_cuts) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  const uint8_t* dp = reinterpret_cast<const uint8_t*>(ctx->dPool.ptr);
  rc = mb200_trim_run(ctx, n, 0, dp, reinterpret_cast<const mb200_seqview*>(ctx->dQcViews.ptr), nullptr, dp,
                      reinterpret_cast<const mb200_seqview*>(ctx->dQcRef.ptr), reinterpret_cast<const mb200_ref_cuts*>(ctx->dQcCuts.ptr),
                      nullptr, maxLen, reinterpret_cast<mb200_trim_result*>(ctx->dQcTrim.ptr), reinterpret_cast<int*>(ctx->dMisc.ptr) + 4);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  uint32_t hMisc[8];
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hMisc, ctx->dMisc.ptr, sizeof(hMisc), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->dQcTrim.ptr, sizeof(mb200_trim_result) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  if (hMisc[4] != 0) return mb200_fail(ctx, (int)hMisc[4], "refTrimKernel: unsupported contig");
  return MB200_OK;
}

int mb200_select_contigs_batch(mb200_ctx* ctx, const mb200_qc_opts* opts, uint32_t nLoci, const uint8_t* isJump,
                               const uint32_t* slotFirst, const uint32_t* nContigs, uint32_t nSlots, const mb200_aln_result* aln,
                               const mb200_jump_result* jump, const uint32_t* cigarPool, uint64_t cigarLen, const uint8_t* pool,
                               uint64_t poolBytes, const mb200_seqview* contigSeq, const mb200_seqview* ref1,
                               const uint32_t* supportCount, mb200_contig_qc* qc, int32_t* best)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (nLoci == 0) return MB200_OK;
  if (!opts || !isJump || !slotFirst || !nContigs || !ref1 || !best || !pool) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (nSlots && (!aln || !jump || !cigarPool || !contigSeq || !supportCount || !qc)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  std::vector<uint32_t> slotLocus(nSlots ? nSlots : 1, 0xFFFFFFFFu);
  for (uint32_t li = 0; li < nLoci; ++li) {
    if ((uint64_t)slotFirst[li] + nContigs[li] > nSlots) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "slot range outside the result arrays");
    if (ref1[li].off + ref1[li].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "view outside the pool");
    for (uint32_t c = 0; c < nContigs[li]; ++c) {
      const uint32_t s = slotFirst[li] + c;
      slotLocus[s]     = li;
      if (isJump[li]) {
        if (jump[s].cigarOff1 + jump[s].cigarLen1 > cigarLen || jump[s].cigarOff2 + jump[s].cigarLen2 > cigarLen)
          return mb200_fail(ctx, MB200_ERR_BAD_ARG, "path outside the cigar pool");
      } else {
        if (aln[s].cigarOff + aln[s].cigarLen > cigarLen) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "path outside the cigar pool");
        if (contigSeq[s].off + contigSeq[s].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "view outside the pool");
        if (aln[s].beginPos < 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "negative beginPos");
        uint64_t span = 0;
        for (uint32_t i = 0; i < aln[s].cigarLen; ++i) {
          const uint32_t t = MB200_SEG_TYPE(cigarPool[aln[s].cigarOff + i]);
          if (t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH)
            span += MB200_SEG_LEN(cigarPool[aln[s].cigarOff + i]);
        }
        if ((uint64_t)aln[s].beginPos + span > ref1[li].len) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "alignment runs past the end of ref1");
      }
    }
  }
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;  // shares device buffers with the fused path's resident results
  int          rc;
  const size_t ns = nSlots ? nSlots : 1;
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dCigar, 4 * (size_t)(cigarLen ? cigarLen : 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAln, sizeof(mb200_aln_result) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dJump, sizeof(mb200_jump_result) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcViews, sizeof(mb200_seqview) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcRef, sizeof(mb200_seqview) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSupport, 4 * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSlotLocus, 4 * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSlotFirst, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcNContigs, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcIsJump, (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQc, sizeof(mb200_contig_qc) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcBest, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
#define MB200_UP(buf, src, bytes) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync((buf).ptr, (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
  MB200_UP(ctx->dPool, pool, poolBytes);
  if (cigarLen) MB200_UP(ctx->dCigar, cigarPool, 4 * (size_t)cigarLen);
  if (nSlots) {
    MB200_UP(ctx->dAln, aln, sizeof(mb200_aln_result) * (size_t)nSlots);
    MB200_UP(ctx->dJump, jump, sizeof(mb200_jump_result) * (size_t)nSlots);
    MB200_UP(ctx->dQcViews, contigSeq, sizeof(mb200_seqview) * (size_t)nSlots);
    MB200_UP(ctx->dQcSupport, supportCount, 4 * (size_t)nSlots);
    MB200_UP(ctx->dQcSlotLocus, slotLocus.data(), 4 * (size_t)nSlots);
  }
  MB200_UP(ctx->dQcRef, ref1, sizeof(mb200_seqview) * (size_t)nLoci);
  MB200_UP(ctx->dQcSlotFirst, slotFirst, 4 * (size_t)nLoci);
  MB200_UP(ctx->dQcNContigs, nContigs, 4 * (size_t)nLoci);
  MB200_UP(ctx->dQcIsJump, isJump, (size_t)nLoci);
#undef MB200_UP
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  const uint8_t* dp = reinterpret_cast<const uint8_t*>(ctx->dPool.ptr);
  rc = mb200_qc_run(ctx, opts, nLoci, nSlots, reinterpret_cast<const uint8_t*>(ctx->dQcIsJump.ptr),
                    reinterpret_cast<const uint32_t*>(ctx->dQcSlotFirst.ptr), reinterpret_cast<const uint32_t*>(ctx->dQcNContigs.ptr),
                    reinterpret_cast<const uint32_t*>(ctx->dQcSlotLocus.ptr), reinterpret_cast<const mb200_aln_result*>(ctx->dAln.ptr),
                    reinterpret_cast<const mb200_jump_result*>(ctx->dJump.ptr), reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr), dp,
                    reinterpret_cast<const mb200_seqview*>(ctx->dQcViews.ptr), dp, reinterpret_cast<const mb200_seqview*>(ctx->dQcRef.ptr),
                    reinterpret_cast<const uint32_t*>(ctx->dQcSupport.ptr), reinterpret_cast<mb200_contig_qc*>(ctx->dQc.ptr),
                    reinterpret_cast<int32_t*>(ctx->dQcBest.ptr), nullptr, 0, nullptr);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  if (nSlots) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(qc, ctx->dQc.ptr, sizeof(mb200_contig_qc) * (size_t)nSlots, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(best, ctx->dQcBest.ptr, 4 * (size_t)nLoci, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}

int mb200_small_sv_batch(mb200_ctx* ctx, uint32_t minCandidateVariantSize, uint32_t nLoci, const uint8_t* isJump,
                         const uint32_t* slotFirst, uint32_t nSlots, const int32_t* best, const mb200_contig_qc* qc,
                         const mb200_aln_result* aln, const uint32_t* cigarPool, uint64_t cigarLen, const uint8_t* pool,
                         uint64_t poolBytes, const mb200_seqview* contigSeq, const mb200_seqview* ref1, uint32_t* svFirst,
                         mb200_small_sv* svs, uint64_t svCap, uint64_t* svUsed)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!svFirst || !svUsed) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  *svUsed = 0;
  if (nLoci == 0) { svFirst[0] = 0; return MB200_OK; }
  if (!isJump || !slotFirst || !best || !ref1 || !pool || (svCap && !svs)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (nSlots && (!qc || !aln || !cigarPool || !contigSeq)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  for (uint32_t li = 0; li < nLoci; ++li) {
    if (ref1[li].off + ref1[li].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "view outside the pool");
    if (isJump[li] || best[li] < 0) continue;
    const uint64_t s = (uint64_t)slotFirst[li] + (uint32_t)best[li];
    if (s >= nSlots) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "best contig outside the result arrays");
    if (aln[s].cigarOff + aln[s].cigarLen > cigarLen) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "path outside the cigar pool");
    if (contigSeq[s].off + contigSeq[s].len > poolBytes) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "view outside the pool");
    if (aln[s].beginPos < 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "negative beginPos");
    uint64_t span = 0, rd = 0;
    for (uint32_t i = 0; i < aln[s].cigarLen; ++i) {
      const uint32_t seg = cigarPool[aln[s].cigarOff + i], t = MB200_SEG_TYPE(seg);
      if (t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH) span += MB200_SEG_LEN(seg);
      if (t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH) rd += MB200_SEG_LEN(seg);
    }
    if ((uint64_t)aln[s].beginPos + span > ref1[li].len) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "alignment runs past the end of ref1");
    if (rd > contigSeq[s].len) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "alignment path longer than the contig");
  }
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;  // shares device buffers with the fused path's resident results
  int          rc;
  const size_t ns = nSlots ? nSlots : 1;
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dCigar, 4 * (size_t)(cigarLen ? cigarLen : 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAln, sizeof(mb200_aln_result) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcViews, sizeof(mb200_seqview) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcRef, sizeof(mb200_seqview) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSlotFirst, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcIsJump, (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQc, sizeof(mb200_contig_qc) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcBest, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvFirst, 4 * (size_t)(nLoci + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvs, sizeof(mb200_small_sv) * (size_t)(svCap ? svCap : 1))) != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
#define MB200_UP(buf, src, bytes) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync((buf).ptr, (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
  MB200_UP(ctx->dPool, pool, poolBytes);
  if (cigarLen) MB200_UP(ctx->dCigar, cigarPool, 4 * (size_t)cigarLen);
  if (nSlots) {
    MB200_UP(ctx->dAln, aln, sizeof(mb200_aln_result) * (size_t)nSlots);
    MB200_UP(ctx->dQcViews, contigSeq, sizeof(mb200_seqview) * (size_t)nSlots);
    MB200_UP(ctx->dQc, qc, sizeof(mb200_contig_qc) * (size_t)nSlots);
  }
  MB200_UP(ctx->dQcRef, ref1, sizeof(mb200_seqview) * (size_t)nLoci);
  MB200_UP(ctx->dQcSlotFirst, slotFirst, 4 * (size_t)nLoci);
  MB200_UP(ctx->dQcIsJump, isJump, (size_t)nLoci);
  MB200_UP(ctx->dQcBest, best, 4 * (size_t)nLoci);
#undef MB200_UP
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  const uint8_t* dp = reinterpret_cast<const uint8_t*>(ctx->dPool.ptr);
  rc = mb200_small_sv_run(ctx, minCandidateVariantSize, nLoci, reinterpret_cast<const uint8_t*>(ctx->dQcIsJump.ptr),
                          reinterpret_cast<const uint32_t*>(ctx->dQcSlotFirst.ptr), reinterpret_cast<const int32_t*>(ctx->dQcBest.ptr),
                          reinterpret_cast<const mb200_contig_qc*>(ctx->dQc.ptr), reinterpret_cast<const mb200_aln_result*>(ctx->dAln.ptr),
                          reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr), dp, reinterpret_cast<const mb200_seqview*>(ctx->dQcViews.ptr), dp,
                          reinterpret_cast<const mb200_seqview*>(ctx->dQcRef.ptr), reinterpret_cast<uint32_t*>(ctx->dSvFirst.ptr),
                          reinterpret_cast<mb200_small_sv*>(ctx->dSvs.ptr), svCap);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(svFirst, ctx->dSvFirst.ptr, 4 * (size_t)(nLoci + 1), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t used = svFirst[nLoci];
  *svUsed             = used;
  if (used > svCap) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "small SV record array too small");
  if (used) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(svs, ctx->dSvs.ptr, sizeof(mb200_small_sv) * (size_t)used, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}

int mb200_jump_sv_batch(mb200_ctx* ctx, uint32_t nLoci, const uint8_t* isJump, const uint32_t* slotFirst, uint32_t nSlots,
                        const int32_t* best, const mb200_jump_result* jump, const uint32_t* cigarPool, uint64_t cigarLen,
                        const mb200_seqview* ref1, const mb200_seqview* ref2, const mb200_jump_orient* orient, mb200_jump_sv* out)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (nLoci == 0) return MB200_OK;
  if (!isJump || !slotFirst || !best || !ref1 || !ref2 || !orient || !out) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  if (nSlots && (!jump || !cigarPool)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  for (uint32_t li = 0; li < nLoci; ++li) {
    if (!isJump[li] || best[li] < 0) continue;
    const uint64_t s = (uint64_t)slotFirst[li] + (uint32_t)best[li];
    if (s >= nSlots) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "best contig outside the result arrays");
    if (jump[s].cigarOff1 + jump[s].cigarLen1 > cigarLen || jump[s].cigarOff2 + jump[s].cigarLen2 > cigarLen)
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "path outside the cigar pool");
  }
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;  // shares device buffers with the fused path's resident results
  int          rc;
  const size_t ns = nSlots ? nSlots : 1;
  if ((rc = mb200_reserve(ctx, ctx->dCigar, 4 * (size_t)(cigarLen ? cigarLen : 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dJump, sizeof(mb200_jump_result) * ns)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcRef, sizeof(mb200_seqview) * (size_t)nLoci * 2)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcSlotFirst, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcIsJump, (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dQcBest, 4 * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvFirst, sizeof(mb200_jump_orient) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dSvs, sizeof(mb200_jump_sv) * (size_t)nLoci)) != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
#define MB200_UP(dst, src, bytes) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
  mb200_seqview* dRef = reinterpret_cast<mb200_seqview*>(ctx->dQcRef.ptr);
  if (cigarLen) MB200_UP(ctx->dCigar.ptr, cigarPool, 4 * (size_t)cigarLen);
  if (nSlots) MB200_UP(ctx->dJump.ptr, jump, sizeof(mb200_jump_result) * (size_t)nSlots);
  MB200_UP(dRef, ref1, sizeof(mb200_seqview) * (size_t)nLoci);
  MB200_UP(dRef + nLoci, ref2, sizeof(mb200_seqview) * (size_t)nLoci);
  MB200_UP(ctx->dQcSlotFirst.ptr, slotFirst, 4 * (size_t)nLoci);
  MB200_UP(ctx->dQcIsJump.ptr, isJump, (size_t)nLoci);
  MB200_UP(ctx->dQcBest.ptr, best, 4 * (size_t)nLoci);
  MB200_UP(ctx->dSvFirst.ptr, orient, sizeof(mb200_jump_orient) * (size_t)nLoci);
#undef MB200_UP
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  rc = mb200_jump_sv_run(ctx, nLoci, reinterpret_cast<const uint8_t*>(ctx->dQcIsJump.ptr), reinterpret_cast<const uint32_t*>(ctx->dQcSlotFirst.ptr),
                         reinterpret_cast<const int32_t*>(ctx->dQcBest.ptr), reinterpret_cast<const mb200_jump_result*>(ctx->dJump.ptr),
                         reinterpret_cast<const uint32_t*>(ctx->dCigar.ptr), dRef, dRef + nLoci,
                         reinterpret_cast<const mb200_jump_orient*>(ctx->dSvFirst.ptr), reinterpret_cast<mb200_jump_sv*>(ctx->dSvs.ptr));
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(out, ctx->dSvs.ptr, sizeof(mb200_jump_sv) * (size_t)nLoci, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}

}  // extern "C"
