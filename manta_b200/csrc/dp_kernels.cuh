#pragma once
// dp_kernels.cuh -- batched contig-to-reference affine-gap aligners for sm_100a.
//
// Replaces (bit-exactly) the three full-matrix aligners of the reference:
//   GlobalAligner<int>::align            lib/alignment/GlobalAlignerImpl.hpp:29-181            (3 states, KIND_GLOBAL)
//   GlobalJumpAligner<int>::align        lib/alignment/GlobalJumpAlignerImpl.hpp:33-333        (4 states, KIND_JUMP)
//   GlobalLargeIndelAligner<int>::align  lib/alignment/GlobalLargeIndelAlignerImpl.hpp:35-225  (5 states, KIND_INDEL)
// and their back-traces (SingleRefAlignerSharedImpl.hpp:75-168, JumpAlignerBaseImpl.hpp:86-242) including
// apath_add_seqmatch (lib/blt_util/align_path_impl.hpp:33-75) and the jumpRange scan.
//
// Mapping to the machine (DESIGN.md "DP kernel"):
//   * one WARP per alignment, persistent warps pulling tasks from a global counter (longest first);
//   * query rows live in registers: thread `lane` owns RT consecutive rows, the warp covers a strip of
//     32*RT rows; queries longer than one strip are swept in passes with the strip-boundary row kept in a
//     per-warp L2-resident buffer;
//   * columns are processed as a skewed (anti-diagonal) wavefront: at step t lane l works on column
//     t-l; the values of its last row reach lane l+1 with one __shfl_up per state per step;
//   * the reference window is staged through shared memory in 2 KiB tiles by TMA bulk copies
//     (cp.async.bulk + mbarrier), double buffered ahead of the wavefront;
//   * back-pointers (2 or 3 bits per state, 1 or 2 bytes per cell, same information as the reference's
//     PtrVal) are written in the skewed [step][lane] order, so every store instruction of a warp covers
//     one contiguous 128..512-byte line -- this stream is the algorithmic HBM traffic of the kernel;
//   * the best back-trace start is tracked as a 64-bit (score, -sequence) key so that the reference's
//     first-strict-maximum rule (updateBacktrace, AlignerUtil.hpp:53-67) is order independent;
//   * the same warp then walks the pointers back (lane 0), emits the run-length path, splits M runs
//     into '='/'X', computes jumpRange and appends the segments to the caller's pool.
//
// Scores are int32 with the finite sentinel badVal = -10000 exactly as in the reference (no saturation).
#include <algorithm>
#include <cstdint>

#include "mb200_internal.h"

#ifndef MB200_USE_TMA
#define MB200_USE_TMA 1
#endif

namespace {

constexpr int BADVAL = -10000;

enum { KIND_GLOBAL = 0, KIND_JUMP = 1, KIND_INDEL = 2 };
enum { ST_MATCH = 0, ST_DELETE = 1, ST_INSERT = 2, ST_JUMP = 3, ST_JUMPINS = 4 };

constexpr int REF_TILE  = 2048;            // bytes per staged reference tile
constexpr int REF_RING  = 2 * REF_TILE;    // two tiles resident per warp
constexpr int WARPS_PER_CTA = 4;

struct DpPair {
  DpTask a, b;
  uint32_t dup;  // 1: b is a copy of a (unpaired task), only a's result is emitted
  uint32_t pad;
};

struct DpParams {
  int match, mismatch, open, extend, offEdge, allowEdgeIns;
  int extra;  // jumpScore (KIND_JUMP) or largeIndelScore (KIND_INDEL)
  int tagBias, tagBad;  // dpTagKernel: score offset and the sentinel's stand-in, in score units
  uint32_t one;         // 1, opaque to ptxas: keeps dpTagKernel's re-tagging adds on the FMA pipe as IMADs
  const DpTask*  tasks;
  uint32_t       nTasks;
  const DpPair*  pairs;   // pair kernel: two tasks per warp (int16x2 lanes)
  uint32_t       nPairs;
  uint32_t*      workCounter;
  uint8_t*       ptrScratch;
  size_t         ptrStride;  // bytes per warp slot
  uint32_t*      segScratch;
  size_t         segStride;  // uint32 per warp slot
  int*           bndScratch;
  size_t         bndStride;  // ints per warp slot
  void*          results;
  uint32_t*      cigarPool;
  unsigned long long cigarCap;
  unsigned long long* cigarUsed;
  int*           errFlag;
};

template <int KIND> struct KindTraits;
template <> struct KindTraits<KIND_GLOBAL> { static constexpr int NS = 3, BITS = 2, CELLB = 1; };
template <> struct KindTraits<KIND_JUMP>   { static constexpr int NS = 4, BITS = 2, CELLB = 1; };
template <> struct KindTraits<KIND_INDEL>  { static constexpr int NS = 5, BITS = 3, CELLB = 2; };

// ---- first-maximum-wins scans (AlignerBase.hpp:46-59, JumpAlignerBase.hpp:93-111, GlobalLargeIndelAligner.hpp:124-151)
__device__ __forceinline__ int max3p(int& m, int v0, int v1, int v2)
{
  int p = 0;
  m     = v0;
  if (v1 > m) { m = v1; p = 1; }
  if (v2 > m) { m = v2; p = 2; }
  return p;
}
__device__ __forceinline__ int max4p(int& m, int v0, int v1, int v2, int v3)
{
  int p = 0;
  m     = v0;
  if (v1 > m) { m = v1; p = 1; }
  if (v2 > m) { m = v2; p = 2; }
  if (v3 > m) { m = v3; p = 3; }
  return p;
}
__device__ __forceinline__ int max5p(int& m, int v0, int v1, int v2, int v3, int v4)
{
  int p = 0;
  m     = v0;
  if (v1 > m) { m = v1; p = 1; }
  if (v2 > m) { m = v2; p = 2; }
  if (v3 > m) { m = v3; p = 3; }
  if (v4 > m) { m = v4; p = 4; }
  return p;
}

__device__ __forceinline__ long long makeKey(int val, uint32_t seq)
{
  return (long long)val * 4294967296LL + (long long)(0x7fffffffu - seq);
}

// ---- shared-memory staging of the reference window ---------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct RefStage {
  // staged coordinates: ref1 occupies [shift1, shift1+R1), ref2 [base2+shift2, base2+shift2+R2)
  const uint8_t* src1;  // 16-byte aligned-down global address of ref1
  const uint8_t* src2;
  uint32_t       shift1, shift2, base2, total;  // total = staged bytes (multiple of 16)
  uint8_t*       ring;                           // REF_RING bytes of shared memory
  uint64_t*      bar;                            // two mbarriers (one per ring slot)
  uint32_t       nTiles;
};

__device__ __forceinline__ void stageIssue(const RefStage& st, uint32_t tile, int lane)
{
  const uint32_t a = tile * REF_TILE;
  uint32_t       b = a + REF_TILE;
  if (b > st.total) b = st.total;
  uint8_t* dst = st.ring + (tile & 1) * REF_TILE;
#if MB200_USE_TMA
  if (lane == 0) {
    const uint32_t barAddr = smemAddr(&st.bar[tile & 1]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(b - a) : "memory");
    // part inside ref1's staged range
    uint32_t e1 = b < st.base2 ? b : st.base2;
    if (a < e1) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smemAddr(dst)),
                   "l"(st.src1 + a), "r"(e1 - a), "r"(barAddr)
                   : "memory");
    }
    uint32_t s2 = a > st.base2 ? a : st.base2;
    if (s2 < b) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smemAddr(dst + (s2 - a))),
                   "l"(st.src2 + (s2 - st.base2)), "r"(b - s2), "r"(barAddr)
                   : "memory");
    }
  }
#else
  for (uint32_t o = a + lane * 16; o < b; o += 32 * 16) {
    const uint8_t* src = (o < st.base2) ? (st.src1 + o) : (st.src2 + (o - st.base2));
    *reinterpret_cast<uint4*>(dst + (o - a)) = *reinterpret_cast<const uint4*>(src);
  }
#endif
}

// ---- back-pointer addressing in the skewed layout --------------------------------------------------
template <int RT, int CELLB> struct PtrLayout {
  static constexpr int BYTES_PER_THREAD = ((RT * CELLB + 3) / 4) * 4;                // per step, padded to 32 bits
  // widest vector store (16, 8 or 4 bytes) that divides the per-thread bytes
  static constexpr int CHUNK_BYTES      = (BYTES_PER_THREAD % 16 == 0) ? 16 : ((BYTES_PER_THREAD % 8 == 0) ? 8 : 4);
  static constexpr int NCHUNK           = BYTES_PER_THREAD / CHUNK_BYTES;
  static constexpr int STEP_BYTES       = 32 * BYTES_PER_THREAD;
  __device__ static __forceinline__ size_t passBytes(uint32_t rtot) { return (size_t)(rtot + 31) * STEP_BYTES; }
  // cell (q>=1, c>=1) -> byte offset
  __device__ static __forceinline__ size_t cellOffset(uint32_t q, uint32_t c, size_t passStride)
  {
    const uint32_t row    = q - 1;
    const uint32_t pass   = row / (32 * RT);
    const uint32_t within = row % (32 * RT);
    const uint32_t lane   = within / RT;
    const uint32_t rr     = within % RT;
    const uint32_t t      = c - 1 + lane;
    const uint32_t boff   = rr * CELLB;
    const uint32_t chunk  = boff / CHUNK_BYTES;
    const uint32_t inner  = boff % CHUNK_BYTES;
    return (size_t)pass * passStride + ((size_t)(t * NCHUNK + chunk) * 32 + lane) * CHUNK_BYTES + inner;
  }
};

// one step's back-pointer words of a lane -> the skewed layout (chunk-major so each store instruction of the warp
// covers one contiguous line)
template <typename PL, int PTRWORDS>
__device__ __forceinline__ void storePtrWords(uint8_t* stepBase, int lane, const uint32_t* pw)
{
  constexpr int CW = PL::CHUNK_BYTES / 4;
#pragma unroll
  for (int ch = 0; ch < PL::NCHUNK; ++ch) {
    uint8_t* dst = stepBase + ((size_t)ch * 32 + lane) * PL::CHUNK_BYTES;
    if (CW == 4) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pw[(ch * 4 + 0) % PTRWORDS], pw[(ch * 4 + 1) % PTRWORDS],
                                                  pw[(ch * 4 + 2) % PTRWORDS], pw[(ch * 4 + 3) % PTRWORDS]);
    } else if (CW == 2) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(pw[(ch * 2 + 0) % PTRWORDS], pw[(ch * 2 + 1) % PTRWORDS]);
    } else {
      *reinterpret_cast<uint32_t*>(dst) = pw[ch % PTRWORDS];
    }
  }
}

// ---- path emission helpers (lane 0 only) -------------------------------------------------------------
struct SegWriter {
  uint32_t* buf;  // raw segments in back-trace (reverse) order
  uint32_t  n;
  uint32_t  type, length;  // the open segment `ps`
  __device__ void updatePath(uint32_t atype)
  {  // AlignerUtil::updatePath
    if (type == atype) return;
    if (type != MB200_SEG_NONE) buf[n++] = MB200_SEG_MAKE(length, type);
    type   = atype;
    length = 0;
  }
  __device__ void push(uint32_t t, uint32_t len) { buf[n++] = MB200_SEG_MAKE(len, t); }
};

__device__ __forceinline__ bool segIsReadLen(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP;
}
__device__ __forceinline__ bool segIsRefLen(uint32_t t) { return t == MB200_SEG_MATCH || t == MB200_SEG_DELETE; }

// Expand a raw path (given in REVERSE order in raw[0..n)) into the final '='/'X' form (apath_add_seqmatch,
// lib/blt_util/align_path_impl.hpp:33-75).  If out == nullptr only counts.  q/r point at the first query/ref base the path
// covers.  By the whole warp: the lanes compare 32 bases of a match run at once and every lane folds the ballot into the
// run-length state (identical in all lanes); lane 0 writes.  (Round 1 did this on one lane: ~ 1 000 dependent loads per
// alignment.)
__device__ uint32_t expandPathWarp(const uint32_t* raw, uint32_t n, const uint8_t* q, const uint8_t* r, uint32_t* out, int lane)
{
  uint32_t count = 0;
  uint32_t curT = MB200_SEG_NONE, curL = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t seg  = raw[n - 1 - k];
    const uint32_t type = MB200_SEG_TYPE(seg);
    const uint32_t len  = MB200_SEG_LEN(seg);
    if (type == MB200_SEG_MATCH) {
      for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t i  = base + (uint32_t)lane;
        bool           m  = false;
        if (i < len) {
          const uint8_t qc = q[i], rc = r[i];
          m = (qc == rc) && (qc != 'N') && (rc != 'N');
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, m);
        const uint32_t nb   = (len - base) < 32u ? (len - base) : 32u;
        uint32_t       pos  = 0;
        while (pos < nb) {
          const uint32_t bit = (bits >> pos) & 1u;
          const uint32_t t   = bit ? MB200_SEG_SEQ_MATCH : MB200_SEG_SEQ_MISMATCH;
          const uint32_t x   = (bit ? ~bits : bits) >> pos;  // first set bit: where the run of equal bits ends
          uint32_t       run = x ? (uint32_t)(__ffs(x) - 1) : (32u - pos);
          if (run > nb - pos) run = nb - pos;
          if (t == curT) {
            curL += run;
          } else {
            if (curT != MB200_SEG_NONE) {
              if (out && lane == 0) out[count] = MB200_SEG_MAKE(curL, curT);
              count++;
            }
            curT = t;
            curL = run;
          }
          pos += run;
        }
      }
      q += len;
      r += len;
    } else {
      if (curT != MB200_SEG_NONE) {
        if (out && lane == 0) out[count] = MB200_SEG_MAKE(curL, curT);
        count++;
        curT = MB200_SEG_NONE;
        curL = 0;
      }
      if (out && lane == 0) out[count] = seg;
      count++;
      if (segIsReadLen(type)) q += len;
      if (segIsRefLen(type)) r += len;
    }
  }
  if (curT != MB200_SEG_NONE) {
    if (out && lane == 0) out[count] = MB200_SEG_MAKE(curL, curT);
    count++;
  }
  return count;
}

__device__ void rawLengths(const uint32_t* raw, uint32_t n, uint32_t& readLen, uint32_t& refLen)
{
  readLen = refLen = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t type = MB200_SEG_TYPE(raw[k]);
    const uint32_t len  = MB200_SEG_LEN(raw[k]);
    if (segIsReadLen(type)) readLen += len;
    if (segIsRefLen(type)) refLen += len;
  }
}

// ---- trace-back + path emission of one finished alignment (whole warp) ----------------------------------
// ptrBuf holds the task's back-pointer cells in the skewed layout with `passStride` bytes per pass.
// The pointer walk is warp-cooperative: in state S the path keeps moving in S's direction (diagonal for MATCH,
// along the row for DELETE/JUMP, along the column for INSERT/JUMPINS) for as long as the visited cells point back
// to S, so lane i speculatively fetches the i-th cell ahead and one ballot finds where the run ends -- one memory
// round trip per <= 32 path steps instead of one per step.  All lanes keep identical copies of the scalar
// bookkeeping; only lane 0 writes segments.
// dpTagKernel keeps the cells of the two alignments of a pair interleaved in ONE stream (the packed registers are
// stored as they are): 2-byte cells as [row][task] 16-bit halves of one 32-bit word per row, 1-byte cells as
// [task][row parity] bytes of one 32-bit word per two rows.
template <int RT, int CELLB> struct PairLayout {
  static constexpr int WORDS            = (CELLB == 2) ? RT : (RT + 1) / 2;  // 32-bit words per thread per step
  static constexpr int BYTES_PER_THREAD = WORDS * 4;
  static constexpr int CHUNK_BYTES      = (BYTES_PER_THREAD % 16 == 0) ? 16 : ((BYTES_PER_THREAD % 8 == 0) ? 8 : 4);
  static constexpr int NCHUNK           = BYTES_PER_THREAD / CHUNK_BYTES;
  static constexpr int STEP_BYTES       = 32 * BYTES_PER_THREAD;
  __device__ static __forceinline__ size_t passBytes(uint32_t rtot) { return (size_t)(rtot + 31) * STEP_BYTES; }
  __device__ static __forceinline__ size_t cellOffset(uint32_t q, uint32_t c, size_t passStride, uint32_t half)
  {
    const uint32_t row    = q - 1;
    const uint32_t pass   = row / (32 * RT);
    const uint32_t within = row % (32 * RT);
    const uint32_t lane   = within / RT;
    const uint32_t rr     = within % RT;
    const uint32_t t      = c - 1 + lane;
    const uint32_t boff   = (CELLB == 2) ? (rr * 4 + half * 2) : ((rr >> 1) * 4 + half * 2 + (rr & 1));
    const uint32_t chunk  = boff / CHUNK_BYTES;
    const uint32_t inner  = boff % CHUNK_BYTES;
    return (size_t)pass * passStride + ((size_t)(t * NCHUNK + chunk) * 32 + lane) * CHUNK_BYTES + inner;
  }
};

// TAGGED: written by dpTagKernel -- interleaved pair layout (`half` selects the alignment), and the fields hold the
// winner's tie-break code: state = TOPC - code
template <int KIND, int RT, bool TAGGED>
__device__ __forceinline__ uint32_t loadPtrField(const uint8_t* ptrBuf, size_t passStride, uint32_t q, uint32_t c, uint32_t state,
                                                 uint32_t half)
{
  using KT = KindTraits<KIND>;
  using PL = PtrLayout<RT, KT::CELLB>;
  using PP = PairLayout<RT, KT::CELLB>;
  const size_t off = TAGGED ? PP::cellOffset(q, c, passStride, half) : PL::cellOffset(q, c, passStride);
  uint32_t cell;
  if (KT::CELLB == 1) cell = ptrBuf[off];
  else cell = *reinterpret_cast<const uint16_t*>(ptrBuf + off);
  const uint32_t f = (cell >> (KT::BITS * state)) & ((1u << KT::BITS) - 1);
  return TAGGED ? ((KIND == KIND_INDEL ? 4u : 3u) - f) : f;
}

template <int KIND, int RT, bool TAGGED = false>
__device__ void emitResult(const DpParams& P, const DpTask& task, const uint8_t* ptrBuf, size_t passStride,
                           uint32_t* segBuf, long long bestKey, bool stageOk, int lane, uint32_t pairHalf = 0)
{
  const uint32_t Q = task.Q, R1 = task.R1;
  const uint32_t R2   = (KIND == KIND_JUMP) ? task.R2 : 0u;
  const uint8_t* qSeq = task.q;
  const uint8_t* ref1 = task.r1;
  const uint8_t* ref2 = task.r2;
  const uint32_t FULL = 0xffffffffu;
  if (!stageOk && lane == 0) atomicExch(P.errFlag, MB200_ERR_CUDA);
  const int      bestVal = (int)(bestKey >> 32);
  const uint32_t seq     = 0x7fffffffu - (uint32_t)(bestKey & 0xffffffffLL);

  if (KIND == KIND_JUMP) {
    // ---------------- JumpAlignerBase::backTraceAlignment  (JumpAlignerBaseImpl.hpp:101-242)
    uint32_t refBegin, queryBegin;
    if (seq < R1) { refBegin = seq + 1; queryBegin = Q; }
    else if (seq < R1 + Q) { refBegin = R1; queryBegin = seq - R1; }
    else if (seq < R1 + Q + R2) { refBegin = R1 + (seq - R1 - Q) + 1; queryBegin = Q; }
    else { refBegin = R1 + R2; queryBegin = seq - (R1 + Q + R2); }
    uint32_t state = ST_MATCH;

    // raw paths in back-trace order: apath1 in segBuf[0..), apath2 in segBuf[half..)
    const uint32_t half = (uint32_t)(P.segStride / 2);
    uint32_t       n1 = 0, n2 = 0;
    uint32_t       psType = MB200_SEG_NONE, psLen = 0;
    uint32_t       jumpInsertSize = 0;
    int            beginPos1 = 0, beginPos2 = 0;
    if (queryBegin < Q) { psType = MB200_SEG_SOFT_CLIP; psLen = Q - queryBegin; }
    bool isRef2End = false;
#define PUSH_TO(isRef1, T, L) do { if (isRef1) { if (lane == 0) segBuf[n1] = MB200_SEG_MAKE(L, T); n1++; } else { if (lane == 0) segBuf[half + n2] = MB200_SEG_MAKE(L, T); n2++; } } while (0)
#define UPDATE_PATH(isRef1, AT) do { if (psType != (AT)) { if (psType != MB200_SEG_NONE) PUSH_TO(isRef1, psType, psLen); psType = (AT); psLen = 0; } } while (0)
    while (queryBegin > 0 && refBegin > 0) {
      if (isRef2End) break;
      const bool isRef1 = (refBegin <= R1);
      if (state == ST_JUMP) {
        if (psType != MB200_SEG_NONE) {
          // first visit of the jump state: close the open segment (no move)
          const uint32_t next = loadPtrField<KIND, RT, TAGGED>(ptrBuf, passStride, queryBegin, refBegin, ST_JUMP, pairHalf);
          beginPos2 = (int)(refBegin - R1);
          if (psType == MB200_SEG_INSERT) {
            jumpInsertSize += psLen;
            psType = MB200_SEG_NONE;
            psLen  = 0;
          } else {
            UPDATE_PATH(false, MB200_SEG_NONE);
          }
          state = next;
          psLen++;
          continue;
        }
        // walking left along the row while the cells keep pointing to JUMP
        const uint32_t avail = refBegin < 32u ? refBegin : 32u;
        uint32_t       nxt   = ST_JUMP;
        if ((uint32_t)lane < avail) nxt = loadPtrField<KIND, RT, TAGGED>(ptrBuf, passStride, queryBegin, refBegin - lane, ST_JUMP, pairHalf);
        const uint32_t diff = __ballot_sync(FULL, ((uint32_t)lane < avail) && nxt != ST_JUMP);
        if (diff) {
          const int n = __ffs(diff) - 1;  // cells 0..n-1 moved, cell n leaves the jump state without moving
          refBegin -= (uint32_t)n;
          state = __shfl_sync(FULL, nxt, n);
          psLen += (uint32_t)n + 1;
        } else {
          refBegin -= avail;
          psLen += avail;
        }
        continue;
      }
      // MATCH / DELETE / INSERT runs
      uint32_t dq = 0, dc = 0, atype, avail;
      if (state == ST_MATCH) { dq = 1; dc = 1; atype = MB200_SEG_MATCH; avail = queryBegin < refBegin ? queryBegin : refBegin; }
      else if (state == ST_DELETE) { dc = 1; atype = MB200_SEG_DELETE; avail = refBegin; }
      else { dq = 1; atype = MB200_SEG_INSERT; avail = queryBegin; }
      if (avail > 32u) avail = 32u;
      uint32_t nxt = state;
      const uint32_t cq = queryBegin - lane * dq, cc = refBegin - lane * dc;
      if ((uint32_t)lane < avail) nxt = loadPtrField<KIND, RT, TAGGED>(ptrBuf, passStride, cq, cc, state, pairHalf);
      const bool valid = (uint32_t)lane < avail;
      // a MATCH step at the first ref2 column that points back to MATCH ends the alignment (isRef2End)
      const bool stop  = valid && (state == ST_MATCH) && (cc == R1 + 1) && (nxt == ST_MATCH);
      const uint32_t ev = __ballot_sync(FULL, valid && (nxt != state || stop));
      UPDATE_PATH(isRef1, atype);
      uint32_t consumed;
      if (ev) {
        const int n = __ffs(ev) - 1;
        consumed    = (uint32_t)n + 1;
        state       = __shfl_sync(FULL, nxt, n);
        isRef2End   = (__shfl_sync(FULL, stop ? 1 : 0, n) != 0);
      } else {
        consumed = avail;
      }
      queryBegin -= consumed * dq;
      refBegin -= consumed * dc;
      psLen += consumed;
    }
    {
      const bool isRef1 = (refBegin < R1);
      if (psType != MB200_SEG_NONE) PUSH_TO(isRef1, psType, psLen);
      if (queryBegin != 0) PUSH_TO(isRef1, MB200_SEG_SOFT_CLIP, queryBegin);
      if (isRef1) beginPos1 = (int)refBegin; else beginPos2 = (int)(refBegin - R1);
    }
#undef PUSH_TO
#undef UPDATE_PATH
    __syncwarp();
    {
      // every lane keeps the same scalars; the '='/'X' expansion is warp-parallel, lane 0 writes
      uint32_t read1, refl1, read2, refl2;
      rawLengths(segBuf, n1, read1, refl1);
      rawLengths(segBuf + half, n2, read2, refl2);
      (void)read2; (void)refl2;
      uint32_t jumpRange = 0;
      if (n1 && n2) {
        uint32_t r1i = (uint32_t)beginPos1 + refl1, r2i = (uint32_t)beginPos2, qi = read1;
        uint32_t jic = jumpInsertSize;
        for (;;) {
          if (r1i == R1) break;
          if (jic > 0) {
            if (qi == Q) break;
            if (ref1[r1i] != qSeq[qi]) break;
          } else {
            if (r2i == R2) break;
            if (ref1[r1i] != ref2[r2i]) break;
          }
          jumpRange++;
          r1i++;
          if (jic > 0) { jic--; qi++; } else { r2i++; }
        }
      }
      const uint8_t* q2 = qSeq + read1 + jumpInsertSize;
      const uint32_t c1 = expandPathWarp(segBuf, n1, qSeq, ref1 + beginPos1, nullptr, lane);
      const uint32_t c2 = expandPathWarp(segBuf + half, n2, q2, ref2 + beginPos2, nullptr, lane);
      unsigned long long off = 0;
      if (lane == 0) off = atomicAdd(P.cigarUsed, (unsigned long long)(c1 + c2));
      off = __shfl_sync(FULL, off, 0);
      if (off + c1 + c2 <= P.cigarCap) {
        expandPathWarp(segBuf, n1, qSeq, ref1 + beginPos1, P.cigarPool + off, lane);
        expandPathWarp(segBuf + half, n2, q2, ref2 + beginPos2, P.cigarPool + off + c1, lane);
      }
      if (lane == 0) {
        mb200_jump_result* o = reinterpret_cast<mb200_jump_result*>(P.results) + task.outIndex;
        o->score = bestVal; o->jumpInsertSize = jumpInsertSize; o->jumpRange = jumpRange;
        o->beginPos1 = beginPos1; o->beginPos2 = beginPos2;
        o->cigarLen1 = c1; o->cigarLen2 = c2; o->reserved = 0;
        o->cigarOff1 = off; o->cigarOff2 = off + c1;
      }
    }
  } else {
    // ---------------- SingleRefAlignerBase::backTraceAlignment (SingleRefAlignerSharedImpl.hpp:75-168)
    const uint32_t R = R1;
    uint32_t       refBegin, queryBegin, state = ST_MATCH;
    if (seq < R) { refBegin = seq + 1; queryBegin = Q; }
    else if (seq == R) { refBegin = R; queryBegin = Q; state = ST_INSERT; }
    else { refBegin = R; queryBegin = seq - R - 1; }
    uint32_t n = 0, psType = MB200_SEG_NONE, psLen = 0, isJumped = 0;
    if (queryBegin < Q) { psType = MB200_SEG_SOFT_CLIP; psLen = Q - queryBegin; }
    for (;;) {
      uint32_t dq = 0, dc = 0, atype, avail;
      if (state == ST_MATCH) { dq = 1; dc = 1; atype = MB200_SEG_MATCH; avail = queryBegin < refBegin ? queryBegin : refBegin; }
      else if (state == ST_DELETE || state == ST_JUMP) { dc = 1; atype = MB200_SEG_DELETE; avail = refBegin; }
      else { dq = 1; atype = MB200_SEG_INSERT; avail = queryBegin; }
      if (avail == 0) break;  // the per-state break conditions of the reference (:114,:119,:124)
      if (avail > 32u) avail = 32u;
      const bool     valid = (uint32_t)lane < avail;
      const uint32_t cq = queryBegin - lane * dq, cc = refBegin - lane * dc;
      uint32_t       nxt = state;
      if (valid) {
        if (cq >= 1 && cc >= 1) nxt = loadPtrField<KIND, RT, TAGGED>(ptrBuf, passStride, cq, cc, state, pairHalf);
        else  // row 0 / column 0 pointers: all MATCH, except the column-0 insert state with edge insertions allowed
          nxt = (cc == 0 && state == ST_INSERT && P.allowEdgeIns) ? (uint32_t)ST_INSERT : (uint32_t)ST_MATCH;
      }
      const uint32_t ev = __ballot_sync(FULL, valid && nxt != state);
      if (psType != atype) {
        if (psType != MB200_SEG_NONE) { if (lane == 0) segBuf[n] = MB200_SEG_MAKE(psLen, psType); n++; }
        psType = atype;
        psLen  = 0;
      }
      if (state == ST_JUMP || state == ST_JUMPINS) isJumped = 1;
      uint32_t consumed;
      if (ev) {
        const int k = __ffs(ev) - 1;
        consumed    = (uint32_t)k + 1;
        state       = __shfl_sync(FULL, nxt, k);
      } else {
        consumed = avail;
      }
      queryBegin -= consumed * dq;
      refBegin -= consumed * dc;
      psLen += consumed;
    }
    if (psType != MB200_SEG_NONE) { if (lane == 0) segBuf[n] = MB200_SEG_MAKE(psLen, psType); n++; }
    if (queryBegin != 0) { if (lane == 0) segBuf[n] = MB200_SEG_MAKE(queryBegin, MB200_SEG_SOFT_CLIP); n++; }
    __syncwarp();
    {
      const int          beginPos = (int)refBegin;
      const uint32_t     cnt      = expandPathWarp(segBuf, n, qSeq, ref1 + beginPos, nullptr, lane);
      unsigned long long off      = 0;
      if (lane == 0) off = atomicAdd(P.cigarUsed, (unsigned long long)cnt);
      off = __shfl_sync(FULL, off, 0);
      if (off + cnt <= P.cigarCap) expandPathWarp(segBuf, n, qSeq, ref1 + beginPos, P.cigarPool + off, lane);
      if (lane == 0) {
        mb200_aln_result* o = reinterpret_cast<mb200_aln_result*>(P.results) + task.outIndex;
        o->score = bestVal; o->beginPos = beginPos; o->isJumped = isJumped; o->cigarLen = cnt; o->cigarOff = off;
      }
    }
  }
  __syncwarp();
}

// ======================================================================================================
template <int KIND, int RT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dpAlignKernel(const DpParams P)
{
  using KT               = KindTraits<KIND>;
  using PL               = PtrLayout<RT, KT::CELLB>;
  constexpr int NS       = KT::NS;
  constexpr int BITS     = KT::BITS;
  constexpr int STRIP    = 32 * RT;
  constexpr int PTRWORDS = (RT * KT::CELLB + 3) / 4;  // uint32 words of pointers per thread per step

  __shared__ __align__(128) uint8_t sRing[WARPS_PER_CTA][REF_RING];
  __shared__ __align__(8) uint64_t sBar[WARPS_PER_CTA][2];
  __shared__ uint32_t sTask[WARPS_PER_CTA];

  const int      lane   = threadIdx.x & 31;
  const int      warp   = threadIdx.x >> 5;
  const uint32_t slot   = blockIdx.x * WARPS_PER_CTA + warp;
  uint8_t*       ptrBuf = P.ptrScratch + (size_t)slot * P.ptrStride;
  uint32_t*      segBuf = P.segScratch + (size_t)slot * P.segStride;
  int*           bndBuf = P.bndScratch + (size_t)slot * P.bndStride;

#if MB200_USE_TMA
  uint32_t tilesUsed[2] = {0, 0};  // completed phases per ring slot (mbarrier parity bookkeeping)
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar[warp][0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar[warp][1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
#endif

  const int open = P.open, extend = P.extend, offEdge = P.offEdge, match = P.match, mismatch = P.mismatch;
  const int extra = P.extra;

  for (;;) {
    if (lane == 0) sTask[warp] = atomicAdd(P.workCounter, 1u);
    __syncwarp();
    const uint32_t taskIdx = sTask[warp];
    __syncwarp();
    if (taskIdx >= P.nTasks) break;
    const DpTask   task = P.tasks[taskIdx];
    const uint32_t Q = task.Q, R1 = task.R1;
    const uint32_t R2   = (KIND == KIND_JUMP) ? task.R2 : 0u;
    const uint32_t RTOT = R1 + R2;
    const uint8_t* qSeq = task.q;
    const uint8_t* ref1 = task.r1;
    const uint8_t* ref2 = task.r2;

    // ---- reference staging geometry
    RefStage st;
    st.shift1 = (uint32_t)(reinterpret_cast<uintptr_t>(ref1) & 15);
    st.src1   = ref1 - st.shift1;
    st.base2  = (st.shift1 + R1 + 15) & ~15u;
    if (KIND == KIND_JUMP) {
      st.shift2 = (uint32_t)(reinterpret_cast<uintptr_t>(ref2) & 15);
      st.src2   = ref2 - st.shift2;
      st.total  = st.base2 + ((st.shift2 + R2 + 15) & ~15u);
    } else {
      st.shift2 = 0;
      st.src2   = st.src1;
      st.total  = st.base2;
    }
    st.ring   = sRing[warp];
    st.bar    = sBar[warp];
    st.nTiles = (st.total + REF_TILE - 1) / REF_TILE;

    long long bestKey = (long long)0x8000000000000000LL;
    bool      stageOk = true;

    const uint32_t nPass = (Q + STRIP - 1) / STRIP;
    for (uint32_t pass = 0; pass < nPass; ++pass) {
      const uint32_t q0 = pass * STRIP + lane * RT;  // rows q0+1 .. q0+RT (1-based)
      // query characters of my rows (0 where beyond the query: such rows never influence valid rows)
      uint32_t qch[(RT + 3) / 4];
#pragma unroll
      for (int w = 0; w < (RT + 3) / 4; ++w) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t row = q0 + w * 4 + b;  // 0-based query index
          const uint32_t ch  = (row < Q) ? (uint32_t)__ldg(qSeq + row) : 0u;
          v |= ch << (8 * b);
        }
        qch[w] = v;
      }

      // restart the reference stream for this pass
#if MB200_USE_TMA
      // (mbarrier phases continue across passes and tasks: parity = number of completed uses of the slot)
#endif
      uint32_t nextIssue = 0;
      __syncwarp();
      // tiles 0 and 1
      {
        stageIssue(st, 0, lane);
        nextIssue = 1;
        if (st.nTiles > 1) {
          stageIssue(st, 1, lane);
          nextIssue = 2;
        }
      }

      int pM[RT], pD[RT], pI[RT], pJ[RT], pK[RT];
      int lastM = 0, lastD = BADVAL, lastI = BADVAL, lastJ = BADVAL, lastK = BADVAL;
      int dgM = 0, dgD = BADVAL, dgI = BADVAL, dgJ = BADVAL, dgK = BADVAL;
#pragma unroll
      for (int r = 0; r < RT; ++r) { pM[r] = 0; pD[r] = pI[r] = pJ[r] = pK[r] = BADVAL; }

      const uint32_t nSteps   = RTOT + 31;
      uint8_t*       passPtr  = ptrBuf + (size_t)pass * PL::passBytes(RTOT);
      uint32_t       waitTile = 0;  // next tile lane 0 has to wait for

      for (uint32_t t = 0; t < nSteps; ++t) {
        // ---- reference tiles: wait when the head of the wavefront enters a tile, refill behind the tail
        {
          // highest staged coordinate touched this step is that of lane 0 (or of the last column)
          uint32_t cz0 = t < RTOT ? t : RTOT - 1;
          uint32_t s0  = (KIND == KIND_JUMP && cz0 >= R1) ? (st.base2 + st.shift2 + (cz0 - R1)) : (st.shift1 + cz0);
          const uint32_t needTile = s0 / REF_TILE;
          while (waitTile <= needTile) {
#if MB200_USE_TMA
            const uint32_t barAddr = smemAddr(&st.bar[waitTile & 1]);
            const uint32_t parity  = tilesUsed[waitTile & 1] & 1;
            uint32_t       ok      = 0;
            for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
              asm volatile(
                  "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                  : "=r"(ok)
                  : "r"(barAddr), "r"(parity)
                  : "memory");
              if (ok) break;
            }
            if (!ok) stageOk = false;
            tilesUsed[waitTile & 1]++;
#else
            __syncwarp();
#endif
            waitTile++;
          }
          // lowest staged coordinate still needed is that of lane 31 (column t-31)
          if (nextIssue < st.nTiles) {
            const uint32_t czT = t >= 31 ? t - 31 : 0;
            uint32_t sT = (KIND == KIND_JUMP && czT >= R1) ? (st.base2 + st.shift2 + (czT - R1)) : (st.shift1 + czT);
            // tile (nextIssue) reuses the ring slot of tile (nextIssue-2): free once the tail has left it
            if (sT / REF_TILE >= nextIssue - 1) {
              __syncwarp();
              stageIssue(st, nextIssue, lane);
              nextIssue++;
            }
          }
        }

        // ---- values of the row above (row q0) for this step's column
        int upM = __shfl_up_sync(0xffffffffu, lastM, 1);
        int upD = __shfl_up_sync(0xffffffffu, lastD, 1);
        int upI = __shfl_up_sync(0xffffffffu, lastI, 1);
        int upJ = (NS >= 4) ? __shfl_up_sync(0xffffffffu, lastJ, 1) : BADVAL;
        int upK = (NS >= 5) ? __shfl_up_sync(0xffffffffu, lastK, 1) : BADVAL;

        const int  c      = (int)t - lane + 1;  // 1-based combined column
        const bool active = (c >= 1) && (c <= (int)RTOT);
        if (lane == 0) {
          if (pass == 0) {
            upM = 0;
            upD = upI = upJ = upK = BADVAL;  // row 0 of every column >= 1
          } else if (active) {
            const int* b = bndBuf + (size_t)c * 5;
            upM = b[0]; upD = b[1]; upI = b[2]; upJ = b[3]; upK = b[4];
          }
        }

        if (active) {
          const bool phase2 = (KIND == KIND_JUMP) && ((uint32_t)c > R1);
          const uint32_t cz = (uint32_t)c - 1;
          const uint32_t sidx = phase2 ? (st.base2 + st.shift2 + (cz - R1)) : (st.shift1 + cz);
          const uint32_t refch = st.ring[sidx & (REF_RING - 1)];

          if (c == 1) {
            // column 0: M=q*offEdge, D=bad, I=(allowEdgeIns ? open+q*extend : bad), others bad
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              const int q = (int)q0 + r + 1;
              pM[r] = q * offEdge;
              pD[r] = BADVAL;
              pI[r] = (KIND != KIND_JUMP && P.allowEdgeIns) ? (open + q * extend) : BADVAL;
              pJ[r] = BADVAL;
              pK[r] = BADVAL;
            }
            dgM = (int)q0 * offEdge;
            dgD = BADVAL;
            dgI = (KIND != KIND_JUMP && P.allowEdgeIns) ? (open + (int)q0 * extend) : BADVAL;
            dgJ = dgK = BADVAL;
          }
          if (KIND == KIND_JUMP && (uint32_t)c == R1 + 1) {
            // re-initialisation between ref1 and ref2 (GlobalJumpAlignerImpl.hpp:197-204): jump is preserved
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              pM[r] = ((int)q0 + r + 1) * offEdge;
              pD[r] = pI[r] = BADVAL;
            }
            dgM = (int)q0 * offEdge;
            dgD = dgI = BADVAL;
            if (q0 == 0) dgJ = BADVAL;
          }

          int dM = dgM, dD = dgD, dI = dgI, dJ = dgJ, dK = dgK;   // row above, previous column
          int vM = upM, vI = upI, vJ = upJ, vK = upK;             // row above, this column
          uint32_t ptrw[PTRWORDS];
#pragma unroll
          for (int w = 0; w < PTRWORDS; ++w) ptrw[w] = 0;

          const bool isLastCol1 = (KIND == KIND_JUMP) && ((uint32_t)c == R1);
          const bool isLastCol  = ((uint32_t)c == RTOT);

#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const uint32_t q   = q0 + r + 1;  // 1-based row
            const uint32_t qc  = (qch[r / 4] >> (8 * (r % 4))) & 0xffu;
            const int      s   = (qc == refch) ? match : mismatch;
            const bool     row1 = (r == 0) && (q0 == 0);
            int nM, nD, nI, nJ = BADVAL, nK = BADVAL;
            int pm, pd, pi, pj = 0, pk = 0;
            if (KIND == KIND_GLOBAL) {
              pm = max3p(nM, dM, dD, dI);
              nM += s;
              pd = max3p(nD, pM[r] + open, pD[r], pI[r]);
              nD += extend;
              if (row1) nD = BADVAL;
              pi = max3p(nI, vM + open, BADVAL, vI);
              nI += extend;
              if (row1) nI = BADVAL;
            } else if (KIND == KIND_INDEL) {
              pm = max5p(nM, dM, dD, dI, dJ, dK);
              nM += s;
              pd = max5p(nD, pM[r] + open, pD[r], pI[r], BADVAL, pK[r]);
              nD += extend;
              if (row1) nD = BADVAL;
              pi = max5p(nI, vM + open, BADVAL, vI, BADVAL, BADVAL);
              nI += extend;
              if (row1) nI = BADVAL;
              pj = max5p(nJ, pM[r] + extra, BADVAL, pI[r] + extra - open, pJ[r], pK[r] + extra);
              if (row1) nJ = BADVAL;
              pk = max5p(nK, vM + extra, BADVAL, BADVAL, BADVAL, vK);
              if (row1) nK = BADVAL;
            } else {  // KIND_JUMP
              if (!phase2) {
                pm = max3p(nM, dM, dD, dI);
                nM += s;
                pd = max3p(nD, pM[r] + open, pD[r], pI[r]);
                nD += extend;
                if (row1) nD = BADVAL;
                pi = max3p(nI, vM + open, BADVAL, vI);
                nI += extend;
                if (row1) nI = BADVAL;
                pj = max4p(nJ, nM + extra, BADVAL, nI + extra, pJ[r]);
              } else {
                pm = max4p(nM, dM, dD, dI, dJ);
                nM += s;
                pd = max3p(nD, pM[r] + open, pD[r], pI[r]);
                nD += extend;
                pi = max4p(nI, vM + open, BADVAL, vI, vJ);
                nI += extend;
                nJ = pJ[r];
                pj = ST_JUMP;
              }
            }
            uint32_t cell = (uint32_t)pm | ((uint32_t)pd << BITS) | ((uint32_t)pi << (2 * BITS));
            if (NS >= 4) cell |= ((uint32_t)pj << (3 * BITS));
            if (NS >= 5) cell |= ((uint32_t)pk << (4 * BITS));
            if (KT::CELLB == 1) {
              ptrw[r / 4] |= cell << (8 * (r % 4));
            } else {
              ptrw[r / 2] |= cell << (16 * (r % 2));
            }

            // shift the stencil down one row
            dM = pM[r]; dD = pD[r]; dI = pI[r]; dJ = pJ[r]; dK = pK[r];
            pM[r] = nM; pD[r] = nD; pI[r] = nI; pJ[r] = nJ; pK[r] = nK;
            vM = nM; vI = nI; vJ = nJ; vK = nK;

            // ---- back-trace start candidates (updateBacktrace call sites of the three aligners)
            if (q <= Q) {
              if (KIND == KIND_JUMP) {
                if (q == Q) {
                  const uint32_t seq = phase2 ? (R1 + Q + ((uint32_t)c - R1) - 1) : ((uint32_t)c - 1);
                  const long long k  = makeKey(nM, seq);
                  bestKey            = k > bestKey ? k : bestKey;
                } else {
                  if (isLastCol1) {
                    const long long k = makeKey(nM + (int)(Q - q) * offEdge, R1 + q);
                    bestKey           = k > bestKey ? k : bestKey;
                  }
                  if (isLastCol) {
                    const long long k = makeKey(nM + (int)(Q - q) * offEdge, R1 + Q + R2 + q);
                    bestKey           = k > bestKey ? k : bestKey;
                  }
                }
              } else {
                if (q == Q) {
                  long long k = makeKey(nM, (uint32_t)c - 1);
                  bestKey     = k > bestKey ? k : bestKey;
                  if (isLastCol && P.allowEdgeIns) {
                    k       = makeKey(nI, R1);
                    bestKey = k > bestKey ? k : bestKey;
                  }
                }
                if (isLastCol && (q < Q || KIND == KIND_INDEL)) {
                  const long long k = makeKey(nM + (int)(Q - q) * offEdge, R1 + 1 + q);
                  bestKey           = k > bestKey ? k : bestKey;
                }
              }
            }
          }

          // row 0 takes part in the off-edge scans as well (queryIndex 0): tM[0] == 0
          if (q0 == 0) {
            if (KIND == KIND_JUMP) {
              if (isLastCol1) { const long long k = makeKey((int)Q * offEdge, R1); bestKey = k > bestKey ? k : bestKey; }
              if (isLastCol) { const long long k = makeKey((int)Q * offEdge, R1 + Q + R2); bestKey = k > bestKey ? k : bestKey; }
            } else if (isLastCol) {
              const long long k = makeKey((int)Q * offEdge, R1 + 1);
              bestKey           = k > bestKey ? k : bestKey;
            }
          }

          // diagonal for the next column = this column's row-above values
          dgM = upM; dgD = upD; dgI = upI; dgJ = upJ; dgK = upK;
          lastM = vM; lastD = pD[RT - 1]; lastI = vI; lastJ = vJ; lastK = vK;

          // ---- back-pointers of this (lane, step): contiguous per warp in the skewed layout
          {
            uint8_t* stepBase = passPtr + (size_t)t * PL::STEP_BYTES;
            storePtrWords<PL, PTRWORDS>(stepBase, lane, ptrw);
          }
          // strip boundary row for the next pass
          if (lane == 31 && pass + 1 < nPass) {
            int* b = bndBuf + (size_t)c * 5;
            b[0] = lastM; b[1] = lastD; b[2] = lastI; b[3] = lastJ; b[4] = lastK;
          }
        }
      }  // steps
      // drain outstanding tile loads of this pass (all issued tiles were waited for above: nextIssue == waitTile
      // unless the reference is shorter than the issued prefetch)
      while (waitTile < nextIssue) {
#if MB200_USE_TMA
        const uint32_t barAddr = smemAddr(&st.bar[waitTile & 1]);
        const uint32_t parity  = tilesUsed[waitTile & 1] & 1;
        uint32_t       ok      = 0;
        for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
          asm volatile(
              "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
              : "=r"(ok)
              : "r"(barAddr), "r"(parity)
              : "memory");
          if (ok) break;
        }
        if (!ok) stageOk = false;
        tilesUsed[waitTile & 1]++;
#endif
        waitTile++;
      }
      __syncwarp();
    }  // passes

    // ---- best back-trace start over the warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, bestKey, off);
      bestKey           = o > bestKey ? o : bestKey;
    }
    __threadfence_block();
    __syncwarp();

    emitResult<KIND, RT>(P, task, ptrBuf, PL::passBytes(RTOT), segBuf, bestKey, stageOk, lane);
    __syncwarp();
  }  // tasks
}


// ======================================================================================================
// int16x2 pair kernel: one warp fills TWO alignments at once, task A in the low and task B in the high 16-bit half
// of every register (inter-task SIMD).  sm_100a executes the packed max with per-half predicate outputs
// (VIMNMX.S16x2 R, P, P, a, b -- __vibmax_s16x2) and the packed add (VIADD.16x2 -- __vadd2) natively, so one
// instruction advances two cells.  Legal only while every score stays inside int16 (the host checks the bound from
// the scores and the query length; otherwise the int32 kernel above runs) -- within the bound 16-bit two's
// complement arithmetic is the reference's int arithmetic, sentinel included.  Jump pairs must share R1.
__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }
__device__ __forceinline__ int lo16(uint32_t x) { return (int)(short)(x & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t x) { return ((int)x) >> 16; }

// m = max(m, v) per half; where v wins strictly the half's pointer field becomes `code` (pre-shifted)
// (`code` is a compile-time constant after unrolling: field position inside the cell + the cell's position inside
// its 32-bit pointer word are folded into it, so the SEL result needs no further shifting)
__device__ __forceinline__ void vmaxp(uint32_t code, uint32_t& m, uint32_t v, uint32_t& fa, uint32_t& fb)
{
  bool ph, pl;
  m  = __vibmax_s16x2(m, v, &ph, &pl);
  fa = pl ? fa : code;
  fb = ph ? fb : code;
}

struct StageCursor {
  RefStage st;
  uint32_t nextIssue, waitTile;
  uint32_t used0, used1;  // completed uses of ring slot 0 / 1 (mbarrier parity); two scalars: an indexed array would
                          // push the whole cursor into local memory
};

__device__ __forceinline__ uint32_t stagedIndex(const RefStage& st, uint32_t cz, uint32_t R1, bool isJump)
{
  return (isJump && cz >= R1) ? (st.base2 + st.shift2 + (cz - R1)) : (st.shift1 + cz);
}

__device__ __forceinline__ bool cursorWait(StageCursor& cur)
{
#if MB200_USE_TMA
  const bool     odd     = (cur.waitTile & 1) != 0;
  const uint32_t barAddr = smemAddr(cur.st.bar) + (odd ? 8u : 0u);
  const uint32_t parity  = (odd ? cur.used1 : cur.used0) & 1;
  uint32_t       ok      = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(barAddr), "r"(parity)
        : "memory");
    if (ok) break;
  }
  if (odd) cur.used1++; else cur.used0++;
  cur.waitTile++;
  return ok != 0;
#else
  __syncwarp();
  cur.waitTile++;
  return true;
#endif
}

// NOBAD: the host proved that the sentinel can never win a max strictly (every running maximum stays above
// badVal for this score set and query length), so the comparisons against badVal -- and their pointer codes --
// are dropped.
// Strips of up to 12 rows per lane are held to 168 registers (3 CTAs/SM).  Left alone, ptxas spends 176-255 registers
// on them and the occupancy drops to 2 CTAs/SM for nothing: with the cap the spills are 0-12 bytes.  Measured: -7 %
// jump time for the cap up to 8 rows, another -8 % up to 12 rows; 128 registers (4 CTAs/SM) gains nothing for 4-6 rows
// and loses for 8; 16 rows spill 130+ bytes at 168 and stay uncapped.
template <int KIND, int RT, bool NOBAD>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, (RT <= 12) ? 3 : 1) dpPairKernel(const DpParams P)
{
  using KT               = KindTraits<KIND>;
  using PL               = PtrLayout<RT, KT::CELLB>;
  constexpr int NS       = KT::NS;
  constexpr int BITS     = KT::BITS;
  constexpr int STRIP    = 32 * RT;
  constexpr int PTRWORDS = (RT * KT::CELLB + 3) / 4;
  constexpr int ROWBITS  = 8 * KT::CELLB;  // bits of one cell in the pointer words
  constexpr int ROWSPW   = 32 / ROWBITS;   // cells per 32-bit pointer word

  __shared__ __align__(128) uint8_t sRing[WARPS_PER_CTA][2][REF_RING];
  __shared__ __align__(8) uint64_t sBar[WARPS_PER_CTA][2][2];
  __shared__ uint32_t sTask[WARPS_PER_CTA];
  // M of every row of a lane that owns the last query row of A or B (read back with a run-time row index: registers
  // cannot be indexed dynamically, and a SEL chain in the row loop would sit on the saturated ALU pipe)
  // ALU-trimming variants (row-Q capture through shared memory, pointer codes pre-shifted to the row's position,
  // staging checks once per 32 steps).  The 5-state kernel sits on the register cliff and ptxas spills inside the
  // row loop with them (measured: 112 ms -> 128..207 ms per 20k loci), so it keeps the plain forms.
  // instruction trimming per aligner kind.  The 5-state kind sits at the register limit of 3 CTAs/SM (165 of 168):
  // pre-shifted pointer codes spill there and cost more than they save (measured: +3 %); checking the reference tiles
  // once per 32 steps saves 6 %, capturing the last query row through shared memory another 4 %.
  constexpr bool PRESHIFT = (KIND != KIND_INDEL);  // pointer codes pre-shifted to their row slot
  constexpr bool STAGE32  = true;                  // reference-tile checks once per 32 steps
  constexpr bool SMEM_CAP = true;                  // last-query-row capture through shared memory
  __shared__ uint32_t sCap[SMEM_CAP ? WARPS_PER_CTA : 1][SMEM_CAP ? 32 : 1][RT];

  const int      lane    = threadIdx.x & 31;
  const int      warp    = threadIdx.x >> 5;
  const uint32_t slot    = blockIdx.x * WARPS_PER_CTA + warp;
  const size_t   halfPtr = P.ptrStride / 2;
  uint8_t*       ptrBufA = P.ptrScratch + (size_t)slot * P.ptrStride;
  uint8_t*       ptrBufB = ptrBufA + halfPtr;
  uint32_t*      segBufA = P.segScratch + (size_t)slot * 2 * P.segStride;
  uint32_t*      segBufB = segBufA + P.segStride;
  int*           bndBuf  = P.bndScratch + (size_t)slot * P.bndStride;

  StageCursor curA, curB;
  curA.used0 = curA.used1 = curB.used0 = curB.used1 = 0;
#if MB200_USE_TMA
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar[warp][h][b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
#endif

  const int      offEdge = P.offEdge;
  const uint32_t open2 = pack2(P.open), ext2 = pack2(P.extend), match2 = pack2(P.match);
  const uint32_t negDelta = (uint32_t)(P.mismatch - P.match) & 0xffffu;  // per-half (mismatch - match)
  const uint32_t extra2 = pack2(P.extra), BAD2 = pack2(BADVAL);
  const uint32_t extraMinusOpen2 = pack2(P.extra - P.open);
  const bool     allowEdgeIns = (KIND != KIND_JUMP) && P.allowEdgeIns;

  for (;;) {
    if (lane == 0) sTask[warp] = atomicAdd(P.workCounter, 1u);
    __syncwarp();
    const uint32_t pairIdx = sTask[warp];
    __syncwarp();
    if (pairIdx >= P.nPairs) break;
    const DpPair   pr = P.pairs[pairIdx];
    const DpTask   tA = pr.a, tB = pr.b;
    const uint32_t QA = tA.Q, QB = tB.Q;
    const uint32_t Q  = QA > QB ? QA : QB;
    const uint32_t R1 = tA.R1;  // == tB.R1 for jump pairs; single-ref kinds: R1 is the whole reference
    const uint32_t R1A = tA.R1, R1B = tB.R1;
    const uint32_t R2A = (KIND == KIND_JUMP) ? tA.R2 : 0u, R2B = (KIND == KIND_JUMP) ? tB.R2 : 0u;
    const uint32_t RTA = R1A + R2A, RTB = R1B + R2B;
    const uint32_t RTOT = RTA > RTB ? RTA : RTB;
    const size_t   passStride = PL::passBytes(RTOT);

    // ---- staging geometry of both halves
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      StageCursor&   cur  = h ? curB : curA;
      const DpTask&  tk   = h ? tB : tA;
      RefStage&      st   = cur.st;
      st.shift1 = (uint32_t)(reinterpret_cast<uintptr_t>(tk.r1) & 15);
      st.src1   = tk.r1 - st.shift1;
      st.base2  = (st.shift1 + tk.R1 + 15) & ~15u;
      if (KIND == KIND_JUMP) {
        st.shift2 = (uint32_t)(reinterpret_cast<uintptr_t>(tk.r2) & 15);
        st.src2   = tk.r2 - st.shift2;
        st.total  = st.base2 + ((st.shift2 + tk.R2 + 15) & ~15u);
      } else {
        st.shift2 = 0;
        st.src2   = st.src1;
        st.total  = st.base2;
      }
      st.ring   = sRing[warp][h];
      st.bar    = sBar[warp][h];
      st.nTiles = (st.total + REF_TILE - 1) / REF_TILE;
    }

    long long bestA = (long long)0x8000000000000000LL, bestB = bestA;
    bool      stageOk = true;

    const uint32_t nPass = (Q + STRIP - 1) / STRIP;
    for (uint32_t pass = 0; pass < nPass; ++pass) {
      const uint32_t q0 = pass * STRIP + lane * RT;
      // query characters: low half task A, high half task B (0 beyond the end)
      uint32_t qpair[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const uint32_t row = q0 + r;
        const uint32_t ca  = (row < QA) ? (uint32_t)__ldg(tA.q + row) : 0u;
        const uint32_t cb  = (row < QB) ? (uint32_t)__ldg(tB.q + row) : 0u;
        qpair[r]           = ca | (cb << 16);
      }
      // which of my rows is the last query row of A / of B (RT == none)
      const int rQA = (QA > q0 && QA <= q0 + RT) ? (int)(QA - q0 - 1) : RT;
      const int rQB = (QB > q0 && QB <= q0 + RT) ? (int)(QB - q0 - 1) : RT;

      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        StageCursor& cur = h ? curB : curA;
        stageIssue(cur.st, 0, lane);
        cur.nextIssue = 1;
        if (cur.st.nTiles > 1) {
          stageIssue(cur.st, 1, lane);
          cur.nextIssue = 2;
        }
        cur.waitTile = 0;
      }

      uint32_t pM[RT], pD[RT], pI[RT], pJ[RT], pK[RT];
      uint32_t lastM = 0, lastD = BAD2, lastI = BAD2, lastJ = BAD2, lastK = BAD2;
      uint32_t dgM = 0, dgD = BAD2, dgI = BAD2, dgJ = BAD2, dgK = BAD2;
#pragma unroll
      for (int r = 0; r < RT; ++r) { pM[r] = 0; pD[r] = pI[r] = pJ[r] = pK[r] = BAD2; }

      const uint32_t nSteps = RTOT + 31;
      uint8_t*       passA  = ptrBufA + (size_t)pass * passStride;
      uint8_t*       passB  = ptrBufB + (size_t)pass * passStride;

      for (uint32_t t = 0; t < nSteps; ++t) {
        // ---- reference tiles of both halves: checked once per 32 steps with a 32-column look-ahead for the head
        // (tiles are 2 KiB, so a tile the head enters inside this block is waited for at the block start)
        if (!STAGE32 || (t & 31u) == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            StageCursor&   cur = h ? curB : curA;
            const uint32_t rt  = h ? RTB : RTA;
            const uint32_t r1  = h ? R1B : R1A;
            const uint32_t ahead = STAGE32 ? 31u : 0u;
            const uint32_t cz0 = (t + ahead < rt) ? t + ahead : rt - 1;
            const uint32_t need = stagedIndex(cur.st, cz0, r1, KIND == KIND_JUMP) / REF_TILE;
            while (cur.waitTile <= need) stageOk = cursorWait(cur) && stageOk;
            if (cur.nextIssue < cur.st.nTiles) {
              uint32_t czT = t >= 31 ? t - 31 : 0;
              if (czT >= rt) czT = rt - 1;
              if (stagedIndex(cur.st, czT, r1, KIND == KIND_JUMP) / REF_TILE >= cur.nextIssue - 1) {
                __syncwarp();
                stageIssue(cur.st, cur.nextIssue, lane);
                cur.nextIssue++;
              }
            }
          }
        }

        uint32_t upM = __shfl_up_sync(0xffffffffu, lastM, 1);
        uint32_t upD = __shfl_up_sync(0xffffffffu, lastD, 1);
        uint32_t upI = __shfl_up_sync(0xffffffffu, lastI, 1);
        uint32_t upJ = (NS >= 4) ? __shfl_up_sync(0xffffffffu, lastJ, 1) : BAD2;
        uint32_t upK = (NS >= 5) ? __shfl_up_sync(0xffffffffu, lastK, 1) : BAD2;

        const int  c      = (int)t - lane + 1;
        const bool active = (c >= 1) && (c <= (int)RTOT);
        if (lane == 0) {
          if (pass == 0) {
            upM = 0;
            upD = upI = upJ = upK = BAD2;
          } else if (active) {
            const uint32_t* b = reinterpret_cast<const uint32_t*>(bndBuf) + (size_t)c * 5;
            upM = b[0]; upD = b[1]; upI = b[2]; upJ = b[3]; upK = b[4];
          }
        }

        if (active) {
          const bool     phase2 = (KIND == KIND_JUMP) && ((uint32_t)c > R1);
          const uint32_t cz  = (uint32_t)c - 1;
          const uint32_t czA = cz < RTA ? cz : RTA - 1;
          const uint32_t czB = cz < RTB ? cz : RTB - 1;
          const uint32_t refA = curA.st.ring[stagedIndex(curA.st, czA, R1A, KIND == KIND_JUMP) & (REF_RING - 1)];
          const uint32_t refB = curB.st.ring[stagedIndex(curB.st, czB, R1B, KIND == KIND_JUMP) & (REF_RING - 1)];
          const uint32_t refpair = refA | (refB << 16);

          if (c == 1) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              const int q = (int)q0 + r + 1;
              pM[r] = pack2(q * offEdge);
              pD[r] = BAD2;
              pI[r] = allowEdgeIns ? pack2(P.open + q * P.extend) : BAD2;
              pJ[r] = BAD2;
              pK[r] = BAD2;
            }
            dgM = pack2((int)q0 * offEdge);
            dgD = BAD2;
            dgI = allowEdgeIns ? pack2(P.open + (int)q0 * P.extend) : BAD2;
            dgJ = dgK = BAD2;
          }
          if (KIND == KIND_JUMP && (uint32_t)c == R1 + 1) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              pM[r] = pack2(((int)q0 + r + 1) * offEdge);
              pD[r] = pI[r] = BAD2;
            }
            dgM = pack2((int)q0 * offEdge);
            dgD = dgI = BAD2;
            if (q0 == 0) dgJ = BAD2;
          }

          uint32_t dM = dgM, dD = dgD, dI = dgI, dJ = dgJ, dK = dgK;
          uint32_t vM = upM, vI = upI, vJ = upJ, vK = upK;
          uint32_t ptrwA[PTRWORDS], ptrwB[PTRWORDS];
#pragma unroll
          for (int w = 0; w < PTRWORDS; ++w) ptrwA[w] = ptrwB[w] = 0;
          uint32_t capA = 0, capB = 0;

#pragma unroll
          for (int r = 0; r < RT; ++r) {
            constexpr uint32_t F0 = 0;  // field shifts inside the 32-bit pointer word are added below
            const int      shRow = ROWBITS * (r % ROWSPW);
            const int      sh    = PRESHIFT ? shRow : 0;  // folded into the SEL constants, or applied once below
            (void)F0;
            // substitution score of both halves: match where the characters are equal, mismatch elsewhere
            const uint32_t d  = qpair[r] ^ refpair;
            const uint32_t nz = ((d + 0x7fff7fffu) & 0x80008000u) >> 15;  // 1 per half where the characters differ
            const uint32_t s2 = __vadd2(match2, nz * negDelta);
            const bool     row1 = (r == 0) && (q0 == 0);
            uint32_t nM, nD, nI, nJ = BAD2, nK = BAD2;
            uint32_t fa = 0, fb = 0;  // pointer fields of this cell, A and B
            if (KIND == KIND_GLOBAL) {
              { uint32_t a = 0, b = 0; nM = dM; vmaxp((1u) << sh, nM, dD, a, b); vmaxp((2u) << sh, nM, dI, a, b); nM = __vadd2(nM, s2); fa |= a; fb |= b; }
              { uint32_t a = 0, b = 0; nD = __vadd2(pM[r], open2); vmaxp((1u << BITS) << sh, nD, pD[r], a, b); vmaxp((2u << BITS) << sh, nD, pI[r], a, b); nD = __vadd2(nD, ext2); if (row1) nD = BAD2; fa |= a; fb |= b; }
              { uint32_t a = 0, b = 0; nI = __vadd2(vM, open2); if (!NOBAD) vmaxp((1u << (2 * BITS)) << sh, nI, BAD2, a, b); vmaxp((2u << (2 * BITS)) << sh, nI, vI, a, b); nI = __vadd2(nI, ext2); if (row1) nI = BAD2; fa |= a; fb |= b; }
            } else if (KIND == KIND_INDEL) {
              { uint32_t a = 0, b = 0; nM = dM; vmaxp((1u) << sh, nM, dD, a, b); vmaxp((2u) << sh, nM, dI, a, b); vmaxp((3u) << sh, nM, dJ, a, b); vmaxp((4u) << sh, nM, dK, a, b); nM = __vadd2(nM, s2); fa |= a; fb |= b; }
              { uint32_t a = 0, b = 0; nD = __vadd2(pM[r], open2); vmaxp((1u << BITS) << sh, nD, pD[r], a, b); vmaxp((2u << BITS) << sh, nD, pI[r], a, b); if (!NOBAD) vmaxp((3u << BITS) << sh, nD, BAD2, a, b); vmaxp((4u << BITS) << sh, nD, pK[r], a, b); nD = __vadd2(nD, ext2); if (row1) nD = BAD2; fa |= a; fb |= b; }
              { uint32_t a = 0, b = 0; nI = __vadd2(vM, open2); if (!NOBAD) vmaxp((1u << (2 * BITS)) << sh, nI, BAD2, a, b); vmaxp((2u << (2 * BITS)) << sh, nI, vI, a, b); nI = __vadd2(nI, ext2); if (row1) nI = BAD2; fa |= a; fb |= b; }
              { uint32_t a = 0, b = 0; nJ = __vadd2(pM[r], extra2); if (!NOBAD) vmaxp((1u << (3 * BITS)) << sh, nJ, BAD2, a, b); vmaxp((2u << (3 * BITS)) << sh, nJ, __vadd2(pI[r], extraMinusOpen2), a, b); vmaxp((3u << (3 * BITS)) << sh, nJ, pJ[r], a, b); vmaxp((4u << (3 * BITS)) << sh, nJ, __vadd2(pK[r], extra2), a, b); if (row1) nJ = BAD2; fa |= a; fb |= b; }
              { uint32_t a = 0, b = 0; nK = __vadd2(vM, extra2); if (!NOBAD) vmaxp((1u << (4 * BITS)) << sh, nK, BAD2, a, b); vmaxp((4u << (4 * BITS)) << sh, nK, vK, a, b); if (row1) nK = BAD2; fa |= a; fb |= b; }
            } else {  // KIND_JUMP
              if (!phase2) {
                { uint32_t a = 0, b = 0; nM = dM; vmaxp((1u) << sh, nM, dD, a, b); vmaxp((2u) << sh, nM, dI, a, b); nM = __vadd2(nM, s2); fa |= a; fb |= b; }
                { uint32_t a = 0, b = 0; nD = __vadd2(pM[r], open2); vmaxp((1u << BITS) << sh, nD, pD[r], a, b); vmaxp((2u << BITS) << sh, nD, pI[r], a, b); nD = __vadd2(nD, ext2); if (row1) nD = BAD2; fa |= a; fb |= b; }
                { uint32_t a = 0, b = 0; nI = __vadd2(vM, open2); if (!NOBAD) vmaxp((1u << (2 * BITS)) << sh, nI, BAD2, a, b); vmaxp((2u << (2 * BITS)) << sh, nI, vI, a, b); nI = __vadd2(nI, ext2); if (row1) nI = BAD2; fa |= a; fb |= b; }
                { uint32_t a = 0, b = 0; nJ = __vadd2(nM, extra2); if (!NOBAD) vmaxp((1u << (3 * BITS)) << sh, nJ, BAD2, a, b); vmaxp((2u << (3 * BITS)) << sh, nJ, __vadd2(nI, extra2), a, b); vmaxp((3u << (3 * BITS)) << sh, nJ, pJ[r], a, b); fa |= a; fb |= b; }
              } else {
                { uint32_t a = 0, b = 0; nM = dM; vmaxp((1u) << sh, nM, dD, a, b); vmaxp((2u) << sh, nM, dI, a, b); vmaxp((3u) << sh, nM, dJ, a, b); nM = __vadd2(nM, s2); fa |= a; fb |= b; }
                { uint32_t a = 0, b = 0; nD = __vadd2(pM[r], open2); vmaxp((1u << BITS) << sh, nD, pD[r], a, b); vmaxp((2u << BITS) << sh, nD, pI[r], a, b); nD = __vadd2(nD, ext2); fa |= a; fb |= b; }
                { uint32_t a = 0, b = 0; nI = __vadd2(vM, open2); if (!NOBAD) vmaxp((1u << (2 * BITS)) << sh, nI, BAD2, a, b); vmaxp((2u << (2 * BITS)) << sh, nI, vI, a, b); vmaxp((3u << (2 * BITS)) << sh, nI, vJ, a, b); nI = __vadd2(nI, ext2); fa |= a; fb |= b; }
                nJ = pJ[r];
                fa |= ((uint32_t)ST_JUMP << (3 * BITS)) << sh;
                fb |= ((uint32_t)ST_JUMP << (3 * BITS)) << sh;
              }
            }
            ptrwA[r / ROWSPW] |= fa << (shRow - sh);
            ptrwB[r / ROWSPW] |= fb << (shRow - sh);

            dM = pM[r]; dD = pD[r]; dI = pI[r]; dJ = pJ[r]; dK = pK[r];
            pM[r] = nM; pD[r] = nD; pI[r] = nI; pJ[r] = nJ; pK[r] = nK;
            vM = nM; vI = nI; vJ = nJ; vK = nK;
            if (!SMEM_CAP) {
              if (r == rQA) capA = nM;
              if (r == rQB) capB = nM;
            }
          }

          // ---- back-trace start candidates (same enumeration as the int32 kernel)
          const bool validA = (uint32_t)c <= RTA, validB = (uint32_t)c <= RTB;
          if (SMEM_CAP && (rQA < RT || rQB < RT)) {
            uint32_t* cp = sCap[warp][lane];
#pragma unroll
            for (int r = 0; r < RT; ++r) cp[r] = pM[r];
            capA = cp[rQA < RT ? rQA : 0];
            capB = cp[rQB < RT ? rQB : 0];
          }
          if (rQA < RT && validA) {
            const uint32_t seq = (KIND == KIND_JUMP && phase2) ? (R1 + QA + ((uint32_t)c - R1) - 1) : ((uint32_t)c - 1);
            const long long k  = makeKey(lo16(capA), seq);
            bestA              = k > bestA ? k : bestA;
          }
          if (rQB < RT && validB) {
            const uint32_t seq = (KIND == KIND_JUMP && phase2) ? (R1 + QB + ((uint32_t)c - R1) - 1) : ((uint32_t)c - 1);
            const long long k  = makeKey(hi16(capB), seq);
            bestB              = k > bestB ? k : bestB;
          }
          const bool edge1 = (KIND == KIND_JUMP) && ((uint32_t)c == R1);
          const bool edgeA = ((uint32_t)c == RTA), edgeB = ((uint32_t)c == RTB);
          if (edge1 || edgeA || edgeB) {
            // off-edge scans after the last column of a reference (rare: not in the hot path)
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              const uint32_t q = q0 + r + 1;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t Qh = h ? QB : QA;
                if (q > Qh) continue;
                const int  m    = h ? hi16(pM[r]) : lo16(pM[r]);
                const bool edge = h ? edgeB : edgeA;
                long long& best = h ? bestB : bestA;
                if (KIND == KIND_JUMP) {
                  const uint32_t R2h = h ? R2B : R2A;
                  if (q < Qh) {
                    if (edge1) { const long long k = makeKey(m + (int)(Qh - q) * offEdge, R1 + q); best = k > best ? k : best; }
                    if (edge) { const long long k = makeKey(m + (int)(Qh - q) * offEdge, R1 + Qh + R2h + q); best = k > best ? k : best; }
                  }
                } else if (edge) {
                  const uint32_t Rh = h ? R1B : R1A;
                  if (q == Qh && allowEdgeIns) {
                    const long long k = makeKey(h ? hi16(pI[r]) : lo16(pI[r]), Rh);
                    best              = k > best ? k : best;
                  }
                  if (q < Qh || KIND == KIND_INDEL) {
                    const long long k = makeKey(m + (int)(Qh - q) * offEdge, Rh + 1 + q);
                    best              = k > best ? k : best;
                  }
                }
              }
            }
            if (q0 == 0) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t Qh = h ? QB : QA;
                const bool edge = h ? edgeB : edgeA;
                long long& best = h ? bestB : bestA;
                if (KIND == KIND_JUMP) {
                  const uint32_t R2h = h ? R2B : R2A;
                  if (edge1) { const long long k = makeKey((int)Qh * offEdge, R1); best = k > best ? k : best; }
                  if (edge) { const long long k = makeKey((int)Qh * offEdge, R1 + Qh + R2h); best = k > best ? k : best; }
                } else if (edge) {
                  const long long k = makeKey((int)Qh * offEdge, (h ? R1B : R1A) + 1);
                  best              = k > best ? k : best;
                }
              }
            }
          }

          dgM = upM; dgD = upD; dgI = upI; dgJ = upJ; dgK = upK;
          lastM = vM; lastD = pD[RT - 1]; lastI = vI; lastJ = vJ; lastK = vK;

          // ---- back-pointers: one contiguous line per warp and task
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint8_t*        stepBase = (h ? passB : passA) + (size_t)t * PL::STEP_BYTES;
            const uint32_t* pw       = h ? ptrwB : ptrwA;
            storePtrWords<PL, PTRWORDS>(stepBase, lane, pw);
          }
          if (lane == 31 && pass + 1 < nPass) {
            uint32_t* b = reinterpret_cast<uint32_t*>(bndBuf) + (size_t)c * 5;
            b[0] = lastM; b[1] = lastD; b[2] = lastI; b[3] = lastJ; b[4] = lastK;
          }
        }
      }  // steps
      while (curA.waitTile < curA.nextIssue) stageOk = cursorWait(curA) && stageOk;
      while (curB.waitTile < curB.nextIssue) stageOk = cursorWait(curB) && stageOk;
      __syncwarp();
    }  // passes

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const long long oa = __shfl_xor_sync(0xffffffffu, bestA, off);
      const long long ob = __shfl_xor_sync(0xffffffffu, bestB, off);
      bestA              = oa > bestA ? oa : bestA;
      bestB              = ob > bestB ? ob : bestB;
    }
    __threadfence_block();
    __syncwarp();
    emitResult<KIND, RT>(P, tA, ptrBufA, passStride, segBufA, bestA, stageOk, lane);
    if (!pr.dup) emitResult<KIND, RT>(P, tB, ptrBufB, passStride, segBufB, bestB, stageOk, lane);
    __syncwarp();
  }
}

// ======================================================================================================
// Tagged-score pair kernel (the default when its range check holds).
//
// Same skewed wavefront, staging, scratch layout and trace-back as dpPairKernel, but the cell update carries the
// back-pointer INSIDE the score instead of deriving it from compare predicates: every state value is held per 16-bit
// half as  u = (score + bias) * SC + code(state)  with SC = 8 (5-state kind) or 4, unsigned, and
// code(MATCH) > code(DELETE) > code(INSERT) > code(JUMP) > code(JUMPINS).  All maxima of the recurrences list their
// candidates in exactly that state order and keep the first maximum (AlignerBase.hpp:46-59, JumpAlignerBase.hpp:93-111,
// GlobalLargeIndelAligner.hpp:124-151), so ONE unsigned maximum over the tagged candidates yields both the score and,
// in its low bits, the code of the first-winning candidate -- i.e. the reference's pointer.  The maxima map to the
// 3-input / add-fused packed instructions of sm_100a (VIMNMX3.U16x2, VIADDMNMX.U16x2: two alignments per
// instruction), the tag is split off with one LOP3, and everything else (re-tagging, the substitution score, the
// pointer-word assembly) is 32-bit integer arithmetic that ptxas can place on the FMA pipe (IMAD) next to the
// saturated ALU pipe.  Because every per-half value provably stays inside [0, 65535] (host check tagRangeOk), 32-bit
// modular arithmetic on the packed word is exact for both halves.  Requires the NOBAD regime (the sentinel never wins,
// so it never has to be an exact -10000: it is any value below every real score).
// Pointer fields hold the code; trace-back decodes state = TOPC - code.
// a * b + c on the FMA pipe (IMAD); asm so that the compiler cannot turn it back into ALU-pipe shifts/logic
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int KIND> struct TagTraits {
  static constexpr int      SH    = (KIND == KIND_INDEL) ? 3 : 2;
  static constexpr int      SC    = 1 << SH;
  static constexpr int      TOPC  = (KIND == KIND_INDEL) ? 4 : 3;
  static constexpr uint32_t MASK2 = (uint32_t)(SC - 1) * 0x00010001u;
  static constexpr int      UNITS = 65536 / SC;  // representable score span
};

#ifndef MB200_TAG_OCC4_MAXRT
#define MB200_TAG_OCC4_MAXRT 0  // strips up to this many rows per lane are held to 128 registers (4 CTAs/SM)
#endif
template <int KIND, int RT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, (RT <= MB200_TAG_OCC4_MAXRT) ? 4 : ((RT <= 12) ? 3 : 1)) dpTagKernel(const DpParams P)
{
  using KT               = KindTraits<KIND>;
  using TT               = TagTraits<KIND>;
  using PP               = PairLayout<RT, KT::CELLB>;
  constexpr int NS       = KT::NS;
  constexpr int STRIP    = 32 * RT;
  constexpr int SH       = TT::SH;
  constexpr int SC       = TT::SC;
  constexpr uint32_t CLEAN = ~TT::MASK2;
  constexpr int cM = TT::TOPC - ST_MATCH, cD = TT::TOPC - ST_DELETE, cI = TT::TOPC - ST_INSERT;
  constexpr int cJ = TT::TOPC - ST_JUMP;  // code of JUMPINS is 0

  __shared__ __align__(128) uint8_t sRing[WARPS_PER_CTA][2][REF_RING];
  __shared__ __align__(8) uint64_t sBar[WARPS_PER_CTA][2][2];
  __shared__ uint32_t sTask[WARPS_PER_CTA];
  __shared__ uint32_t sCap[WARPS_PER_CTA][32][RT];

  const int      lane    = threadIdx.x & 31;
  const int      warp    = threadIdx.x >> 5;
  const uint32_t slot    = blockIdx.x * WARPS_PER_CTA + warp;
  uint8_t*       ptrBuf  = P.ptrScratch + (size_t)slot * P.ptrStride;  // both alignments, interleaved (PairLayout)
  uint32_t*      segBufA = P.segScratch + (size_t)slot * 2 * P.segStride;
  uint32_t*      segBufB = segBufA + P.segStride;
  int*           bndBuf  = P.bndScratch + (size_t)slot * P.bndStride;

  StageCursor curA, curB;
  curA.used0 = curA.used1 = curB.used0 = curB.used1 = 0;
#if MB200_USE_TMA
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar[warp][h][b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
#endif

  const int      offEdge = P.offEdge;
  const int      bias    = P.tagBias;
  const uint32_t ONE = P.one, NEG1 = 0u - P.one;
  // T(v, code): both halves hold the tagged score v
  auto T = [bias](int v, int code) -> uint32_t { return (uint32_t)((v + bias) * SC + code) * 0x00010001u; };
  // per-half two's-complement constants for the packed add-max instructions
  const uint32_t open16 = pack2(P.open * SC), extra16 = pack2(P.extra * SC), emo16 = pack2((P.extra - P.open) * SC);
  // 32-bit modular constants: (delta * SC + code) on both halves (exact while every half stays in range)
  const uint32_t matchC = (uint32_t)(P.match * SC + cM) * 0x00010001u;
  const uint32_t negDeltaT = (uint32_t)((P.mismatch - P.match) * SC);  // times (nzA + 65536 nzB)
  const uint32_t extD = (uint32_t)(P.extend * SC + cD) * 0x00010001u;
  const uint32_t extI = (uint32_t)(P.extend * SC + cI) * 0x00010001u;
  const uint32_t codeJ2 = (uint32_t)cJ * 0x00010001u;
  const uint32_t BADT = (uint32_t)(P.tagBad * SC) * 0x00010001u;
  const bool     allowEdgeIns = (KIND != KIND_JUMP) && P.allowEdgeIns;

  for (;;) {
    if (lane == 0) sTask[warp] = atomicAdd(P.workCounter, 1u);
    __syncwarp();
    const uint32_t pairIdx = sTask[warp];
    __syncwarp();
    if (pairIdx >= P.nPairs) break;
    const DpPair   pr = P.pairs[pairIdx];
    const DpTask   tA = pr.a, tB = pr.b;
    const uint32_t QA = tA.Q, QB = tB.Q;
    const uint32_t Q  = QA > QB ? QA : QB;
    const uint32_t R1 = tA.R1;
    const uint32_t R1A = tA.R1, R1B = tB.R1;
    const uint32_t R2A = (KIND == KIND_JUMP) ? tA.R2 : 0u, R2B = (KIND == KIND_JUMP) ? tB.R2 : 0u;
    const uint32_t RTA = R1A + R2A, RTB = R1B + R2B;
    const uint32_t RTOT = RTA > RTB ? RTA : RTB;
    const size_t   passStride = PP::passBytes(RTOT);

#pragma unroll
    for (int h = 0; h < 2; ++h) {
      StageCursor&   cur  = h ? curB : curA;
      const DpTask&  tk   = h ? tB : tA;
      RefStage&      st   = cur.st;
      st.shift1 = (uint32_t)(reinterpret_cast<uintptr_t>(tk.r1) & 15);
      st.src1   = tk.r1 - st.shift1;
      st.base2  = (st.shift1 + tk.R1 + 15) & ~15u;
      if (KIND == KIND_JUMP) {
        st.shift2 = (uint32_t)(reinterpret_cast<uintptr_t>(tk.r2) & 15);
        st.src2   = tk.r2 - st.shift2;
        st.total  = st.base2 + ((st.shift2 + tk.R2 + 15) & ~15u);
      } else {
        st.shift2 = 0;
        st.src2   = st.src1;
        st.total  = st.base2;
      }
      st.ring   = sRing[warp][h];
      st.bar    = sBar[warp][h];
      st.nTiles = (st.total + REF_TILE - 1) / REF_TILE;
    }

    long long bestA = (long long)0x8000000000000000LL, bestB = bestA;
    bool      stageOk = true;
    // tagged half -> score
    auto valLo = [bias](uint32_t x) -> int { return (int)((x & 0xffffu) >> SH) - bias; };
    auto valHi = [bias](uint32_t x) -> int { return (int)(x >> (16 + SH)) - bias; };

    const uint32_t nPass = (Q + STRIP - 1) / STRIP;
    uint8_t* const ringA = sRing[warp][0];
    uint8_t* const ringB = sRing[warp][1];
    // staged coordinate of column c (0-based cz) = cz + off1 in ref1, cz + off2 in ref2 (jump pairs share R1)
    const uint32_t off1A = curA.st.shift1, off1B = curB.st.shift1;
    const uint32_t off2A = curA.st.base2 + curA.st.shift2 - R1A, off2B = curB.st.base2 + curB.st.shift2 - R1B;
    for (uint32_t pass = 0; pass < nPass; ++pass) {
      const uint32_t q0 = pass * STRIP + lane * RT;
      uint32_t qpair[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const uint32_t row = q0 + r;
        const uint32_t ca  = (row < QA) ? (uint32_t)__ldg(tA.q + row) : 0u;
        const uint32_t cb  = (row < QB) ? (uint32_t)__ldg(tB.q + row) : 0u;
        qpair[r]           = ca | (cb << 16);
      }
      const int  rQA = (QA > q0 && QA <= q0 + RT) ? (int)(QA - q0 - 1) : RT;
      const int  rQB = (QB > q0 && QB <= q0 + RT) ? (int)(QB - q0 - 1) : RT;
      const bool ownsQ = (rQA < RT) || (rQB < RT);
      // running maxima of the last query row over the columns: (tagged M << 16) | (0xffff - column), so that the unsigned
      // maximum keeps the first (lowest-column) maximum -- the start candidates enumerate in column order in both
      // phases.  Turned into the 64-bit (score, -sequence) key once per pass.
      uint32_t keyA = 0, keyB = 0;

      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        StageCursor& cur = h ? curB : curA;
        stageIssue(cur.st, 0, lane);
        cur.nextIssue = 1;
        if (cur.st.nTiles > 1) {
          stageIssue(cur.st, 1, lane);
          cur.nextIssue = 2;
        }
        cur.waitTile = 0;
      }

      // column 0 of my rows (a lane's registers are first touched at its column 1, step t == lane)
      uint32_t pM[RT], pD[RT], pI[RT], pJ[RT], pK[RT];
      const uint32_t M0 = T(0, cM);
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const int q = (int)q0 + r + 1;
        pM[r] = T(q * offEdge, cM);
        pD[r] = BADT;
        pI[r] = allowEdgeIns ? T(P.open + q * P.extend, cI) : BADT;
        pJ[r] = BADT;
        pK[r] = BADT;
      }
      uint32_t lastM = M0, lastD = BADT, lastI = BADT, lastJ = BADT, lastK = BADT;
      uint32_t dgM = T((int)q0 * offEdge, cM), dgD = BADT, dgI = allowEdgeIns ? T(P.open + (int)q0 * P.extend, cI) : BADT;
      uint32_t dgJ = BADT, dgK = BADT;

      const uint32_t nSteps = RTOT + 31;
      const bool     isTop  = (lane == 0) && (pass == 0);
      uint8_t*       stp    = ptrBuf + (size_t)pass * passStride + (size_t)lane * PP::CHUNK_BYTES;  // my chunk 0 of step t
      uint32_t       cdn    = 0xffffu + (uint32_t)lane;                                              // 0xffff - c + 1 at step 0

      // One anti-diagonal step.  The wavefront is peeled (below): in the long middle of a strip every lane is active and, in a
      // jump alignment, every lane is on the same side of the jump column, so those steps run a copy of this body in which
      // the per-lane tests are compile-time constants (ALLACT: cz < RTOT holds; PH = 1 / 2: all lanes in ref1 / ref2 and not
      // at the jump column; PH = 0: decide per lane).
      auto step = [&](const uint32_t t, auto hotTag, auto phTag) {
        constexpr bool HOT    = decltype(hotTag)::value;  // a step of the strip's interior: every lane active, none at an edge
        constexpr bool ALLACT = HOT;
        constexpr int  PH     = decltype(phTag)::value;
        if ((t & 31u) == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            StageCursor&   cur = h ? curB : curA;
            const uint32_t rt  = h ? RTB : RTA;
            const uint32_t r1  = h ? R1B : R1A;
            const uint32_t cz0 = (t + 31u < rt) ? t + 31u : rt - 1;
            const uint32_t need = stagedIndex(cur.st, cz0, r1, KIND == KIND_JUMP) / REF_TILE;
            while (cur.waitTile <= need) stageOk = cursorWait(cur) && stageOk;
            if (cur.nextIssue < cur.st.nTiles) {
              uint32_t czT = t >= 31 ? t - 31 : 0;
              if (czT >= rt) czT = rt - 1;
              if (stagedIndex(cur.st, czT, r1, KIND == KIND_JUMP) / REF_TILE >= cur.nextIssue - 1) {
                __syncwarp();
                stageIssue(cur.st, cur.nextIssue, lane);
                cur.nextIssue++;
              }
            }
          }
        }

        uint32_t upM = __shfl_up_sync(0xffffffffu, lastM, 1);
        uint32_t upD = __shfl_up_sync(0xffffffffu, lastD, 1);
        uint32_t upI = __shfl_up_sync(0xffffffffu, lastI, 1);
        uint32_t upJ = (NS >= 4) ? __shfl_up_sync(0xffffffffu, lastJ, 1) : BADT;
        uint32_t upK = (NS >= 5) ? __shfl_up_sync(0xffffffffu, lastK, 1) : BADT;
        if (isTop) {  // row 0 of every column
          upM = M0;
          upD = upI = upJ = upK = BADT;
        }
        cdn--;
        const uint32_t cz     = t - (uint32_t)lane;  // 0-based column; wraps for lanes that have not started
        const bool     active = ALLACT || (cz < RTOT);
        if (pass != 0 && lane == 0 && active) {
          const uint32_t* b = reinterpret_cast<const uint32_t*>(bndBuf) + (size_t)(cz + 1) * 5;
          upM = b[0]; upD = b[1]; upI = b[2]; upJ = b[3]; upK = b[4];
        }

        if (active) {
          const uint32_t c      = cz + 1;
          const bool     phase2 = (KIND == KIND_JUMP) && (PH == 2 || (PH == 0 && cz >= R1));
          // reference characters of this column.  A column past the shorter alignment's own end reads whatever the ring
          // holds there: those cells are never used (no start candidate, no trace-back), and the halves do not interact.
          const uint32_t refA = ringA[(cz + (phase2 ? off2A : off1A)) & (REF_RING - 1)];
          const uint32_t refB = ringB[(cz + (phase2 ? off2B : off1B)) & (REF_RING - 1)];
          const uint32_t refpair = refA | (refB << 16);

          if (KIND == KIND_JUMP && PH == 0 && cz == R1) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              pM[r] = T(((int)q0 + r + 1) * offEdge, cM);
              pD[r] = pI[r] = BADT;
            }
            dgM = T((int)q0 * offEdge, cM);
            dgD = dgI = BADT;
            if (q0 == 0) dgJ = BADT;
          }

          uint32_t dM = dgM, dD = dgD, dI = dgI, dJ = dgJ, dK = dgK;
          uint32_t vM = upM, vI = upI, vJ = upJ, vK = upK;
          uint32_t ptrw[PP::WORDS];  // pointer cells of both alignments, as stored

#pragma unroll
          for (int r = 0; r < RT; ++r) {
            // substitution score of both halves with MATCH's code folded in (1 per half where the characters differ)
            const uint32_t nz = __vminu2(qpair[r] ^ refpair, 0x00010001u);
            const uint32_t sC = imad(nz, negDeltaT, matchC);
            const bool     row1 = (r == 0) && (q0 == 0);
            uint32_t nM, nD, nI, nJ = BADT, nK = BADT;
            uint32_t w, cl, cell;  // cell = [B cell : A cell], field k at bits k*SH of each half
            // per state: w = tagged maximum (ALU: packed 3-input / add-fused max), cl = w without its tag (ALU: LOP3);
            // the tag w - cl goes into the cell and cl is re-tagged with the state's own code by IMADs (FMA pipe)
#define MB200_FIELD0() do { cl = w & CLEAN; cell = imad(cl, NEG1, w); } while (0)
#define MB200_FIELD(K) do { cl = w & CLEAN; cell = imad(w, 1u << ((K) * SH), cell); cell = imad(cl, 0u - (1u << ((K) * SH)), cell); } while (0)
            if (KIND == KIND_GLOBAL) {
              w = __vimax3_u16x2(dM, dD, dI);
              MB200_FIELD0(); nM = imad(cl, ONE, sC);
              w = __vmaxu2(__viaddmax_u16x2(pM[r], open16, pD[r]), pI[r]);
              MB200_FIELD(1); nD = imad(cl, ONE, extD); if (row1) nD = BADT;
              w = __viaddmax_u16x2(vM, open16, vI);
              MB200_FIELD(2); nI = imad(cl, ONE, extI); if (row1) nI = BADT;
            } else if (KIND == KIND_INDEL) {
              w = __vimax3_u16x2(__vimax3_u16x2(dM, dD, dI), dJ, dK);
              MB200_FIELD0(); nM = imad(cl, ONE, sC);
              w = __vimax3_u16x2(__viaddmax_u16x2(pM[r], open16, pD[r]), pI[r], pK[r]);
              MB200_FIELD(1); nD = imad(cl, ONE, extD); if (row1) nD = BADT;
              w = __viaddmax_u16x2(vM, open16, vI);
              MB200_FIELD(2); nI = imad(cl, ONE, extI); if (row1) nI = BADT;
              w = __viaddmax_u16x2(pI[r], emo16, __viaddmax_u16x2(__vmaxu2(pM[r], pK[r]), extra16, pJ[r]));
              MB200_FIELD(3); nJ = imad(cl, ONE, codeJ2); if (row1) nJ = BADT;
              w = __viaddmax_u16x2(vM, extra16, vK);
              MB200_FIELD(4); nK = cl; if (row1) nK = BADT;
            } else {  // KIND_JUMP
              if (!phase2) {
                w = __vimax3_u16x2(dM, dD, dI);
                MB200_FIELD0(); nM = imad(cl, ONE, sC);
                w = __vmaxu2(__viaddmax_u16x2(pM[r], open16, pD[r]), pI[r]);
                MB200_FIELD(1); nD = imad(cl, ONE, extD); if (row1) nD = BADT;
                w = __viaddmax_u16x2(vM, open16, vI);
                MB200_FIELD(2); nI = imad(cl, ONE, extI); if (row1) nI = BADT;
                w = __viaddmax_u16x2(__vmaxu2(nM, nI), extra16, pJ[r]);
                MB200_FIELD(3); nJ = imad(cl, ONE, codeJ2);
              } else {
                w = __vmaxu2(__vimax3_u16x2(dM, dD, dI), dJ);
                MB200_FIELD0(); nM = imad(cl, ONE, sC);
                w = __vmaxu2(__viaddmax_u16x2(pM[r], open16, pD[r]), pI[r]);
                MB200_FIELD(1); nD = imad(cl, ONE, extD);
                w = __vmaxu2(__viaddmax_u16x2(vM, open16, vI), vJ);
                MB200_FIELD(2); nI = imad(cl, ONE, extI);
                nJ = pJ[r];  // the JUMP field of a ref2 cell is the constant JUMP, code 0
              }
            }
#undef MB200_FIELD0
#undef MB200_FIELD
            if (KT::CELLB == 2) ptrw[r] = cell;
            else if ((r & 1) == 0) ptrw[r / 2] = cell;
            else ptrw[r / 2] = imad(cell, 256u, ptrw[r / 2]);

            dM = pM[r]; dD = pD[r]; dI = pI[r]; dJ = pJ[r]; dK = pK[r];
            pM[r] = nM; pD[r] = nD; pI[r] = nI; pJ[r] = nJ; pK[r] = nK;
            vM = nM; vI = nI; vJ = nJ; vK = nK;
          }

          // ---- back-trace start candidates: the last query row of A / B in this lane
          if (ownsQ) {
            uint32_t* cp = sCap[warp][lane];
#pragma unroll
            for (int r = 0; r < RT; ++r) cp[r] = pM[r];
            if (rQA < RT && c <= RTA) {
              const uint32_t k = (cp[rQA] << 16) | cdn;
              keyA             = k > keyA ? k : keyA;
            }
            if (rQB < RT && c <= RTB) {
              const uint32_t k = (cp[rQB] & 0xffff0000u) | cdn;
              keyB             = k > keyB ? k : keyB;
            }
          }
          const bool edge1 = (KIND == KIND_JUMP) && PH == 0 && (c == R1);
          const bool edgeA = !HOT && (c == RTA), edgeB = !HOT && (c == RTB);
          if (edge1 || edgeA || edgeB) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              const uint32_t q = q0 + r + 1;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t Qh = h ? QB : QA;
                if (q > Qh) continue;
                const int  m    = h ? valHi(pM[r]) : valLo(pM[r]);
                const bool edge = h ? edgeB : edgeA;
                long long& best = h ? bestB : bestA;
                if (KIND == KIND_JUMP) {
                  const uint32_t R2h = h ? R2B : R2A;
                  if (q < Qh) {
                    if (edge1) { const long long k = makeKey(m + (int)(Qh - q) * offEdge, R1 + q); best = k > best ? k : best; }
                    if (edge) { const long long k = makeKey(m + (int)(Qh - q) * offEdge, R1 + Qh + R2h + q); best = k > best ? k : best; }
                  }
                } else if (edge) {
                  const uint32_t Rh = h ? R1B : R1A;
                  if (q == Qh && allowEdgeIns) {
                    const long long k = makeKey(h ? valHi(pI[r]) : valLo(pI[r]), Rh);
                    best              = k > best ? k : best;
                  }
                  if (q < Qh || KIND == KIND_INDEL) {
                    const long long k = makeKey(m + (int)(Qh - q) * offEdge, Rh + 1 + q);
                    best              = k > best ? k : best;
                  }
                }
              }
            }
            if (q0 == 0) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t Qh = h ? QB : QA;
                const bool edge = h ? edgeB : edgeA;
                long long& best = h ? bestB : bestA;
                if (KIND == KIND_JUMP) {
                  const uint32_t R2h = h ? R2B : R2A;
                  if (edge1) { const long long k = makeKey((int)Qh * offEdge, R1); best = k > best ? k : best; }
                  if (edge) { const long long k = makeKey((int)Qh * offEdge, R1 + Qh + R2h); best = k > best ? k : best; }
                } else if (edge) {
                  const long long k = makeKey((int)Qh * offEdge, (h ? R1B : R1A) + 1);
                  best              = k > best ? k : best;
                }
              }
            }
          }

          dgM = upM; dgD = upD; dgI = upI; dgJ = upJ; dgK = upK;
          lastM = vM; lastD = pD[RT - 1]; lastI = vI; lastJ = vJ; lastK = vK;

          // ---- back-pointers: chunk ch of this step at stp + ch * 32 * CHUNK (one contiguous line per warp and chunk)
          {
            constexpr int CW = PP::CHUNK_BYTES / 4;
#pragma unroll
            for (int ch = 0; ch < PP::NCHUNK; ++ch) {
              uint8_t* dst = stp + (size_t)ch * 32 * PP::CHUNK_BYTES;
              if (CW == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(ptrw[ch * 4], ptrw[ch * 4 + 1], ptrw[ch * 4 + 2], ptrw[ch * 4 + 3]);
              else if (CW == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(ptrw[ch * 2], ptrw[ch * 2 + 1]);
              else *reinterpret_cast<uint32_t*>(dst) = ptrw[ch];
            }
          }
          if (lane == 31 && pass + 1 < nPass) {
            uint32_t* b = reinterpret_cast<uint32_t*>(bndBuf) + (size_t)c * 5;
            b[0] = lastM; b[1] = lastD; b[2] = lastI; b[3] = lastJ; b[4] = lastK;
          }
        }
      };  // step
      {
        using std::integral_constant;
        const integral_constant<bool, true>  hot;
        const integral_constant<bool, false> generic;
        const integral_constant<int, 0>      phAny;
        const integral_constant<int, 1>      ph1;
        const integral_constant<int, 2>      ph2;
        // The wavefront is peeled.  A step is "interior" when all 32 lanes are active (31 <= t < RTOT) and no lane sits at an
        // event column: the last column of A or B (c == RTA / RTB: t in [RT? - 1, RT? + 31)) or, in a jump alignment, the
        // jump column (t in [R1 - 1, R1 + 32)).  Interior steps run the specialised copy (per-lane tests and the whole
        // edge block compiled out; jump: one copy per side of the jump column), everything else the generic copy.
#ifndef MB200_DP_PEEL_KINDS
#define MB200_DP_PEEL_KINDS 7  // bit per kind; measured at 50 000 loci: jump 81 -> 54 ms, large-indel 82 -> 78 ms
#endif
        constexpr bool PEEL = ((MB200_DP_PEEL_KINDS >> KIND) & 1) != 0;
        uint32_t t = 0;
        if (PEEL) {
          while (t < nSteps) {
            // boundaries of the event windows, ascending search for the first one beyond t
            const uint32_t bnd[8] = {31u, RTOT, RTA >= 1 ? RTA - 1 : 0u, RTA + 31u, RTB >= 1 ? RTB - 1 : 0u, RTB + 31u,
                                     (KIND == KIND_JUMP && R1 >= 1) ? R1 - 1 : 0u, (KIND == KIND_JUMP) ? R1 + 32u : 0u};
            uint32_t next = nSteps;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (bnd[i] > t && bnd[i] < next) next = bnd[i];
            const bool inA = (t + 1 >= RTA) && (t < RTA + 31u), inB = (t + 1 >= RTB) && (t < RTB + 31u);
            const bool inJ = (KIND == KIND_JUMP) && (t + 1 >= R1) && (t < R1 + 32u);
            const bool interior = (t >= 31u) && (t < RTOT) && !inA && !inB && !inJ;
            if (!interior) {
              for (; t < next; ++t, stp += PP::STEP_BYTES) step(t, generic, phAny);
            } else if (KIND == KIND_JUMP && t >= R1) {
              for (; t < next; ++t, stp += PP::STEP_BYTES) step(t, hot, ph2);
            } else if (KIND == KIND_JUMP) {
              for (; t < next; ++t, stp += PP::STEP_BYTES) step(t, hot, ph1);
            } else {
              // two steps per iteration let the row registers alternate instead of being moved back at the loop edge
              // (large-indel 77.6 -> 73.8 ms; the two jump loops lose with it: 54 -> 61 ms)
#pragma unroll 2
              for (; t < next; ++t, stp += PP::STEP_BYTES) step(t, hot, ph1);
            }
          }
        } else {
          for (; t < nSteps; ++t, stp += PP::STEP_BYTES) step(t, generic, phAny);
        }
      }
      // this pass's best last-row cells as (score, -sequence) keys
      if (keyA) {
        const uint32_t c   = 0xffffu - (keyA & 0xffffu);
        const uint32_t seq = (KIND == KIND_JUMP && c > R1) ? (R1 + QA + (c - R1) - 1) : (c - 1);
        const long long k  = makeKey((int)(keyA >> (16 + SH)) - bias, seq);
        bestA              = k > bestA ? k : bestA;
      }
      if (keyB) {
        const uint32_t c   = 0xffffu - (keyB & 0xffffu);
        const uint32_t seq = (KIND == KIND_JUMP && c > R1) ? (R1 + QB + (c - R1) - 1) : (c - 1);
        const long long k  = makeKey((int)(keyB >> (16 + SH)) - bias, seq);
        bestB              = k > bestB ? k : bestB;
      }
      while (curA.waitTile < curA.nextIssue) stageOk = cursorWait(curA) && stageOk;
      while (curB.waitTile < curB.nextIssue) stageOk = cursorWait(curB) && stageOk;
      __syncwarp();
    }  // passes

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const long long oa = __shfl_xor_sync(0xffffffffu, bestA, off);
      const long long ob = __shfl_xor_sync(0xffffffffu, bestB, off);
      bestA              = oa > bestA ? oa : bestA;
      bestB              = ob > bestB ? ob : bestB;
    }
    __threadfence_block();
    __syncwarp();
    emitResult<KIND, RT, true>(P, tA, ptrBuf, passStride, segBufA, bestA, stageOk, lane, 0);
    if (!pr.dup) emitResult<KIND, RT, true>(P, tB, ptrBuf, passStride, segBufB, bestB, stageOk, lane, 1);
    __syncwarp();
  }
}

// ======================================================================================================
// host side
template <int KIND, int RT> size_t ptrBytesFor(uint32_t Q, uint32_t rtot)
{
  using PL = PtrLayout<RT, KindTraits<KIND>::CELLB>;
  const size_t nPass = (Q + 32 * RT - 1) / (32 * RT);
  return nPass * ((size_t)(rtot + 31) * PL::STEP_BYTES);
}

struct ClassPlan {
  std::vector<DpTask> tasks;
  size_t              maxPtr = 0, maxSeg = 0, maxBnd = 0;
};

template <int KIND, int RT>
int launchClass(mb200_ctx* ctx, const DpParams& base, const ClassPlan& plan, size_t taskOffset, uint32_t* dCounter)
{
  if (plan.tasks.empty()) return MB200_OK;
  int ctasPerSm = 0;
  MB200_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, dpAlignKernel<KIND, RT>,
                                                                      WARPS_PER_CTA * 32, 0));
  if (ctasPerSm < 1) ctasPerSm = 1;
  if (ctx->cfg.ctasPerSm && (int)ctx->cfg.ctasPerSm < ctasPerSm) ctasPerSm = (int)ctx->cfg.ctasPerSm;
  size_t   nSlotsWanted = (size_t)ctx->smCount * ctasPerSm * WARPS_PER_CTA;
  uint32_t grid         = (uint32_t)(ctx->smCount * ctasPerSm);
  // no more warps than tasks
  if (nSlotsWanted > plan.tasks.size()) {
    grid = (uint32_t)((plan.tasks.size() + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
  }
  // scratch budget: shrink the grid if the back-pointer scratch would exceed the configured budget
  const size_t budget   = ctx->cfg.scratchBytes ? (size_t)ctx->cfg.scratchBytes : ((size_t)48 << 30);
  const size_t ptrSlot  = (plan.maxPtr + 255) & ~(size_t)255;
  while (grid > 1 && (size_t)grid * WARPS_PER_CTA * ptrSlot > budget) grid = (grid + 1) / 2;
  const size_t nSlots = (size_t)grid * WARPS_PER_CTA;

  int rc;
  if ((rc = mb200_reserve(ctx, ctx->dPtrScratch, nSlots * ptrSlot)) != MB200_OK) return rc;
  const size_t segSlot = (plan.maxSeg + 63) & ~(size_t)63;
  if ((rc = mb200_reserve(ctx, ctx->dSegScratch, nSlots * segSlot * sizeof(uint32_t))) != MB200_OK) return rc;
  const size_t bndSlot = (plan.maxBnd + 63) & ~(size_t)63;
  if ((rc = mb200_reserve(ctx, ctx->dBndScratch, nSlots * bndSlot * sizeof(int) + 256)) != MB200_OK) return rc;

  DpParams P    = base;
  P.tasks       = reinterpret_cast<const DpTask*>(ctx->dTasks.ptr) + taskOffset;
  P.nTasks      = (uint32_t)plan.tasks.size();
  P.workCounter = dCounter;
  P.ptrScratch  = reinterpret_cast<uint8_t*>(ctx->dPtrScratch.ptr);
  P.ptrStride   = ptrSlot;
  P.segScratch  = reinterpret_cast<uint32_t*>(ctx->dSegScratch.ptr);
  P.segStride   = segSlot;
  P.bndScratch  = reinterpret_cast<int*>(ctx->dBndScratch.ptr);
  P.bndStride   = bndSlot;
  dpAlignKernel<KIND, RT><<<grid, WARPS_PER_CTA * 32, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

struct PairPlan {
  size_t   nPairs = 0;
  size_t   maxPtr = 0, maxSeg = 0, maxBnd = 0;
  uint32_t maxQ = 0, maxR = 0;
  bool     noBad = false, tag = false;
  int      tagBias = 0, tagBad = 0;
};

// can badVal ever win a max?  M[q][c] >= -L*q with L = max(|match|,|mismatch|,|offEdge|) (induction over the
// recurrences, column 0 and the ref2 re-initialisation included); every max that has badVal as a candidate starts
// from M + open / M + extra / M + jump, so below this bound the sentinel never compares strictly greater.
inline bool sentinelNeverWins(const mb200_scores* sc, int extra, uint32_t maxQ)
{
  auto ab = [](int v) { return (long long)(v < 0 ? -v : v); };
  const long long L = std::max(std::max(ab(sc->match), ab(sc->mismatch)), ab(sc->offEdge));
  return L * (long long)(maxQ + 1) + std::max(ab(sc->open), ab(extra)) + ab(sc->extend) + 8 < 10000;
}

// dpTagKernel's range: with match >= 0 and every other score <= 0 all real state values lie in
// [-L*maxQ - G, match*maxQ], G = |open| + |extend| + 2|extra| (a state other than MATCH is at least a MATCH value plus
// one gap/jump constant, a candidate adds at most one more), and the sentinel's stand-in needs G of room below it and
// 2G between it and the lowest real candidate.  `units` = 65536 / SC score units fit one 16-bit half.
inline bool tagRangeOk(const mb200_scores* sc, int extra, uint32_t maxQ, int units, int* bias, int* bad)
{
  if (sc->match < 0 || sc->mismatch > 0 || sc->open > 0 || sc->extend > 0 || sc->offEdge > 0 || extra > 0) return false;
  auto ab = [](int v) { return (long long)(v < 0 ? -v : v); };
  const long long L = std::max(std::max(ab(sc->match), ab(sc->mismatch)), ab(sc->offEdge));
  const long long G = ab(sc->open) + ab(sc->extend) + 2 * ab(extra);
  const long long b = G + 8;
  const long long off = b + 2 * G + L * (long long)(maxQ + 1) + 16;
  if (off + (long long)sc->match * (maxQ + 2) + 16 >= units) return false;
  *bias = (int)off;
  *bad  = (int)b;
  return true;
}

// which kernel runs a class: the tagged-score kernel (TAG), else the predicate/SEL pair kernel with or without the
// sentinel comparisons (NOBAD)
template <int KIND, int RT, bool NOBAD, bool TAG>
int launchPairClass(mb200_ctx* ctx, const DpParams& base, const PairPlan& plan, size_t pairOffset, uint32_t* dCounter,
                    int cls, int nActive)
{
  if (plan.nPairs == 0) return MB200_OK;
  auto kernel = TAG ? dpTagKernel<KIND, RT> : dpPairKernel<KIND, RT, NOBAD>;
  DeviceBuffer& dPtr = ctx->dPtrC[cls];
  DeviceBuffer& dSeg = ctx->dSegC[cls];
  DeviceBuffer& dBnd = ctx->dBndC[cls];
  cudaStream_t  strm = ctx->dpStream[cls];
  int ctasPerSm = 0;
  MB200_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, WARPS_PER_CTA * 32, 0));
  if (ctasPerSm < 1) ctasPerSm = 1;
  if (ctx->cfg.ctasPerSm && (int)ctx->cfg.ctasPerSm < ctasPerSm) ctasPerSm = (int)ctx->cfg.ctasPerSm;
  uint32_t grid = (uint32_t)(ctx->smCount * ctasPerSm);
  if ((size_t)grid * WARPS_PER_CTA > plan.nPairs) grid = (uint32_t)((plan.nPairs + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
  const size_t budget  = (ctx->cfg.scratchBytes ? (size_t)ctx->cfg.scratchBytes : ((size_t)96 << 30)) / (size_t)(nActive > 0 ? nActive : 1);
  const size_t ptrSlot = (2 * plan.maxPtr + 511) & ~(size_t)511;  // both tasks of a pair
  while (grid > 1 && (size_t)grid * WARPS_PER_CTA * ptrSlot > budget) grid = (grid + 1) / 2;
  const size_t nSlots = (size_t)grid * WARPS_PER_CTA;
  int rc;
  if ((rc = mb200_reserve(ctx, dPtr, nSlots * ptrSlot)) != MB200_OK) return rc;
  const size_t segSlot = (plan.maxSeg + 63) & ~(size_t)63;
  if ((rc = mb200_reserve(ctx, dSeg, nSlots * 2 * segSlot * sizeof(uint32_t))) != MB200_OK) return rc;
  const size_t bndSlot = (plan.maxBnd + 63) & ~(size_t)63;
  if ((rc = mb200_reserve(ctx, dBnd, nSlots * bndSlot * sizeof(int) + 256)) != MB200_OK) return rc;
  DpParams P    = base;
  P.tagBias     = plan.tagBias;
  P.tagBad      = plan.tagBad;
  P.pairs       = reinterpret_cast<const DpPair*>(ctx->dTasksK[KIND].ptr) + pairOffset;
  P.nPairs      = (uint32_t)plan.nPairs;
  P.tasks       = nullptr;
  P.nTasks      = 0;
  P.workCounter = dCounter;
  P.ptrScratch  = reinterpret_cast<uint8_t*>(dPtr.ptr);
  P.ptrStride   = ptrSlot;
  P.segScratch  = reinterpret_cast<uint32_t*>(dSeg.ptr);
  P.segStride   = segSlot;
  P.bndScratch  = reinterpret_cast<int*>(dBnd.ptr);
  P.bndStride   = bndSlot;
  MB200_CUDA_CHECK(ctx, cudaStreamWaitEvent(strm, ctx->dpFork, 0));
  MB200_CUDA_CHECK(ctx, cudaStreamWaitEvent(strm, ctx->copyDone[KIND], 0));
  kernel<<<grid, WARPS_PER_CTA * 32, 0, strm>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->dpDone[cls], strm));
  MB200_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->dpDone[cls], 0));
  ctx->timing.launches++;
  return MB200_OK;
}

// every intermediate of the recurrences stays inside int16 for this score set and query length?
inline bool fitsInt16(const mb200_scores* sc, int extra, uint32_t maxQ)
{
  auto ab = [](int v) { return (long long)(v < 0 ? -v : v); };
  const long long perBase = std::max(std::max(ab(sc->match), ab(sc->mismatch)), std::max(ab(sc->offEdge), ab(sc->extend)));
  const long long bound   = 10000 + perBase * (long long)(maxQ + 2) + 2 * ab(sc->open) + 2 * ab(extra) + 4 * ab(sc->extend) + 64;
  return bound < 32000;
}

template <int KIND>
int dpRunPairs(mb200_ctx* ctx, const mb200_scores* scores, int extra, std::vector<DpTask>& tasks, const DpParams& base,
               uint32_t* dWorkCounters, uint64_t* cellsOut, uint64_t* algoBytesOut)
{
  // classes by query length: RT rows per lane, strip = 32*RT rows (padding rows are wasted work, so the classes
  // are fine-grained where contigs live: 130..400 bases)
  constexpr int NCLS = 7;
  static const int RTS[NCLS] = {4, 6, 8, 9, 10, 12, 16};
  // (an earlier version kept the 2-byte kind on multiples of 4 only; with the classes running concurrently the finer
  // strips pay there too: -9 % large-indel time on the config-2 contig lengths, padding waste 1.30 -> 1.12)
  auto classOf = [](uint32_t Q) {
    for (int c = 0; c < NCLS - 1; ++c)
      if (Q <= (uint32_t)(32 * RTS[c])) return c;
    return NCLS - 1;
  };
  // Order of work: class, then (jump) R1 descending so that pair members share it, then cost descending.  This runs on
  // the critical path between the assembler and the DP kernels, so it is one LSD radix sort over a 64-bit key
  // (class:3 | ~R1:32 (jump only) | quantised cost, inverted:29) instead of a comparison sort; equal keys keep task order.
  uint64_t maxCost = 1;
  for (const DpTask& t : tasks) maxCost = std::max(maxCost, (uint64_t)t.Q * ((uint64_t)t.R1 + t.R2));
  int costShift = 0;
  while ((maxCost >> costShift) >= (1ull << 29)) ++costShift;
  const size_t          n = tasks.size();
  std::vector<uint64_t> key(n), key2(n);
  std::vector<uint32_t> idx(n), idx2(n);
  for (size_t i = 0; i < n; ++i) {
    const DpTask&  t = tasks[i];
    const uint64_t cost = ((uint64_t)t.Q * ((uint64_t)t.R1 + t.R2)) >> costShift;
    key[i] = ((uint64_t)classOf(t.Q) << 61) | (KIND == KIND_JUMP ? ((uint64_t)(uint32_t)(~t.R1) << 29) : 0u) | ((1ull << 29) - 1 - cost);
    idx[i] = (uint32_t)i;
  }
  for (int shift = 0; shift < 64; shift += 11) {
    size_t hist[2049];
    memset(hist, 0, sizeof(hist));
    for (size_t i = 0; i < n; ++i) hist[((key[i] >> shift) & 2047) + 1]++;
    bool single = false;
    for (int d = 0; d < 2048; ++d)
      if (hist[d + 1] == n) single = true;
    if (single) continue;  // every key has the same digit here
    for (int d = 0; d < 2048; ++d) hist[d + 1] += hist[d];
    for (size_t i = 0; i < n; ++i) {
      const size_t o = hist[(key[i] >> shift) & 2047]++;
      key2[o] = key[i];
      idx2[o] = idx[i];
    }
    key.swap(key2);
    idx.swap(idx2);
  }
  // pair neighbours; pairs come out class by class, most expensive first (dynamic hand-out, longest processing time first)
  PairPlan plans[NCLS];
  size_t   offsets[NCLS];
  int      rc;
  if ((rc = mb200_reserve_pinned_buf(ctx, ctx->hTasksK[KIND], (n + 1) * sizeof(DpPair))) != MB200_OK) return rc;
  DpPair* all     = reinterpret_cast<DpPair*>(ctx->hTasksK[KIND].ptr);
  size_t  nAll    = 0;
  int     lastCls = -1;
  for (size_t i = 0; i < n;) {
    DpPair pr;
    memset(&pr, 0, sizeof(pr));
    pr.a           = tasks[idx[i]];
    const int cls  = (int)(key[i] >> 61);
    const bool can = (i + 1 < n) && (int)(key[i + 1] >> 61) == cls && (KIND != KIND_JUMP || tasks[idx[i + 1]].R1 == pr.a.R1);
    if (can) {
      pr.b   = tasks[idx[i + 1]];
      pr.dup = 0;
      i += 2;
    } else {
      pr.b   = pr.a;
      pr.dup = 1;
      i += 1;
    }
    if (cls != lastCls) {
      offsets[cls] = nAll;
      lastCls      = cls;
    }
    const uint32_t Q    = std::max(pr.a.Q, pr.b.Q);
    const uint32_t rtot = std::max(pr.a.R1 + pr.a.R2, pr.b.R1 + pr.b.R2);
    const int      RT   = RTS[cls];
    const size_t   nPass = (Q + 32 * RT - 1) / (32 * RT);
    const size_t   pb    = nPass * ((size_t)(rtot + 31) * 32 * (size_t)(((RT * KindTraits<KIND>::CELLB + 3) / 4) * 4));  // PtrLayout::STEP_BYTES
    PairPlan&      p     = plans[cls];
    p.nPairs++;
    p.maxQ   = std::max(p.maxQ, Q);
    p.maxR   = std::max(p.maxR, rtot);
    p.maxPtr = std::max(p.maxPtr, pb);
    p.maxSeg = std::max(p.maxSeg, 2 * ((size_t)Q + rtot + 8));
    p.maxBnd = std::max(p.maxBnd, (Q > (uint32_t)(32 * RT)) ? ((size_t)rtot + 2) * 5 : (size_t)8);
    for (int h = 0; h < (pr.dup ? 1 : 2); ++h) {
      const DpTask& t = h ? pr.b : pr.a;
      if (cellsOut) *cellsOut += (uint64_t)t.Q * (t.R1 + t.R2);
      if (algoBytesOut) *algoBytesOut += (uint64_t)t.Q * (t.R1 + t.R2) * KindTraits<KIND>::CELLB + t.Q + t.R1 + t.R2;
    }
    all[nAll++] = pr;
  }
  for (int c = 0; c < NCLS; ++c)
    if (plans[c].nPairs == 0) offsets[c] = 0;
  if ((rc = mb200_reserve(ctx, ctx->dTasksK[KIND], nAll * sizeof(DpPair) + 64)) != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dTasksK[KIND].ptr, all, nAll * sizeof(DpPair), cudaMemcpyHostToDevice, ctx->copyStream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->copyDone[KIND], ctx->copyStream));
  int nActive = 0;
  for (int c = 0; c < NCLS; ++c) nActive += plans[c].nPairs == 0 ? 0 : 1;
  // every class runs on its own stream with its own scratch, forked from / joined to the ctx stream, so that the
  // tail of one class overlaps the body of the next (largest strips first).  Kernel choice per class, from the
  // longest query of the class: tagged scores where the range allows, else predicates/SEL with or without the sentinel.
  for (int c = 0; c < NCLS; ++c) {
    PairPlan& p = plans[c];
    p.noBad = sentinelNeverWins(scores, extra, p.maxQ);
    p.tag   = p.noBad && !ctx->noTagKernel && p.maxR < 65535u /* 16-bit column in the start-candidate keys */ && tagRangeOk(scores, extra, p.maxQ, TagTraits<KIND>::UNITS, &p.tagBias, &p.tagBad);
  }
#define MB200_LAUNCH_CLASS(RTV, C)                                                                                   \
  do {                                                                                                               \
    if (plans[C].tag) rc = launchPairClass<KIND, RTV, true, true>(ctx, base, plans[C], offsets[C], dWorkCounters + C, C, nActive); \
    else if (plans[C].noBad) rc = launchPairClass<KIND, RTV, true, false>(ctx, base, plans[C], offsets[C], dWorkCounters + C, C, nActive); \
    else rc = launchPairClass<KIND, RTV, false, false>(ctx, base, plans[C], offsets[C], dWorkCounters + C, C, nActive); \
    if (rc != MB200_OK) return rc;                                                                                   \
  } while (0)
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->dpFork, ctx->stream));
  MB200_LAUNCH_CLASS(16, 6);
  MB200_LAUNCH_CLASS(12, 5);
  MB200_LAUNCH_CLASS(10, 4);
  MB200_LAUNCH_CLASS(9, 3);
  MB200_LAUNCH_CLASS(8, 2);
  MB200_LAUNCH_CLASS(6, 1);
  MB200_LAUNCH_CLASS(4, 0);
#undef MB200_LAUNCH_CLASS
  return MB200_OK;
}

template <int KIND>
int dpRunKind(mb200_ctx* ctx, const mb200_scores* scores, int extra, std::vector<DpTask>& tasks, void* dResults,
              uint32_t* dCigar, uint64_t cigarCap, unsigned long long* dCigarUsed, int* dErrFlag,
              uint32_t* dWorkCounters, uint64_t* cellsOut, uint64_t* algoBytesOut)
{
  DpParams P;
  memset(&P, 0, sizeof(P));
  P.match = scores->match; P.mismatch = scores->mismatch; P.open = scores->open; P.extend = scores->extend;
  P.offEdge = scores->offEdge; P.allowEdgeIns = scores->isAllowEdgeInsertion ? 1 : 0; P.extra = extra;
  P.one       = 1;
  P.results   = dResults;
  P.cigarPool = dCigar;
  P.cigarCap  = cigarCap;
  P.errFlag   = dErrFlag;
  P.cigarUsed = dCigarUsed;

  uint32_t maxQ = 0;
  for (const DpTask& t : tasks) maxQ = std::max(maxQ, t.Q);
  if (!ctx->forceInt32 && fitsInt16(scores, extra, maxQ)) return dpRunPairs<KIND>(ctx, scores, extra, tasks, P, dWorkCounters, cellsOut, algoBytesOut);

  // ---- general int32 path: one alignment per warp, classes by query length, longest tasks first
  ClassPlan plans[3];
  for (const DpTask& t : tasks) {
    const uint32_t rtot = t.R1 + t.R2;
    const int      cls  = (t.Q <= 128) ? 0 : (t.Q <= 256 ? 1 : 2);
    ClassPlan&     p    = plans[cls];
    p.tasks.push_back(t);
    size_t pb = cls == 0 ? ptrBytesFor<KIND, 4>(t.Q, rtot) : (cls == 1 ? ptrBytesFor<KIND, 8>(t.Q, rtot) : ptrBytesFor<KIND, 16>(t.Q, rtot));
    if (pb > p.maxPtr) p.maxPtr = pb;
    const size_t seg = 2 * ((size_t)t.Q + rtot + 8);
    if (seg > p.maxSeg) p.maxSeg = seg;
    const size_t bnd = (t.Q > 512) ? ((size_t)rtot + 2) * 5 : 8;
    if (bnd > p.maxBnd) p.maxBnd = bnd;
    if (cellsOut) *cellsOut += (uint64_t)t.Q * rtot;
    // inputs + one packed back-pointer cell per DP cell (1 B jump/global, 2 B large-indel)
    if (algoBytesOut) *algoBytesOut += (uint64_t)t.Q * rtot * KindTraits<KIND>::CELLB + t.Q + rtot;
  }
  std::vector<DpTask> all;
  all.reserve(tasks.size());
  size_t offsets[3];
  for (int c = 0; c < 3; ++c) {
    std::vector<DpTask>& v = plans[c].tasks;
    std::sort(v.begin(), v.end(), [](const DpTask& a, const DpTask& b) {
      const uint64_t ca = (uint64_t)a.Q * (a.R1 + a.R2), cb = (uint64_t)b.Q * (b.R1 + b.R2);
      return ca != cb ? ca > cb : a.outIndex < b.outIndex;
    });
    offsets[c] = all.size();
    all.insert(all.end(), v.begin(), v.end());
  }
  int rc;
  if ((rc = mb200_reserve(ctx, ctx->dTasks, all.size() * sizeof(DpTask) + 64)) != MB200_OK) return rc;
  // pageable source: the runtime stages it before returning, so `all` may go out of scope
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dTasks.ptr, all.data(), all.size() * sizeof(DpTask),
                                        cudaMemcpyHostToDevice, ctx->stream));
  // the classes run back to back on the ctx stream (they share the scratch buffers)
  if ((rc = launchClass<KIND, 4>(ctx, P, plans[0], offsets[0], dWorkCounters + 0)) != MB200_OK) return rc;
  if ((rc = launchClass<KIND, 8>(ctx, P, plans[1], offsets[1], dWorkCounters + 1)) != MB200_OK) return rc;
  if ((rc = launchClass<KIND, 16>(ctx, P, plans[2], offsets[2], dWorkCounters + 2)) != MB200_OK) return rc;
  return MB200_OK;
}

}  // namespace

// one instantiation per aligner kind lives in its own translation unit (dp_global.cu / dp_jump.cu / dp_indel.cu)
// so that the three sets of kernels compile in parallel
template <int KIND>
int mb200_dp_run_kind(mb200_ctx* ctx, const mb200_scores* scores, int extra, std::vector<DpTask>& tasks, void* dResults,
                      uint32_t* dCigar, uint64_t cigarCap, unsigned long long* dCigarUsed, int* dErrFlag,
                      uint32_t* dWorkCounters, uint64_t* cellsOut, uint64_t* algoBytesOut)
{
  return dpRunKind<KIND>(ctx, scores, extra, tasks, dResults, dCigar, cigarCap, dCigarUsed, dErrFlag, dWorkCounters,
                         cellsOut, algoBytesOut);
}
