// assemble.cu -- batched iterative-k de-Bruijn assembler for sm_100a.
//
// Replaces (bit-exactly) runIterativeAssembler, lib/assembly/IterativeAssembler.cpp:844-931, with
//   getKmerCounts :506-550, getRepeatKmers/searchRepeats :555-642, buildContigs :644-720, walk :149-501,
//   selectContigs :722-842.
//
// Two kernels run the same algorithm: assembleKernel below keeps its tables in global memory (one warp per
// locus, 32 loci per SM: the throughput kernel of large batches) and assembleChipKernel further down keeps them in shared
// memory (one 128-thread CTA per locus, 3 loci per SM: a sixth of the latency, taken for batches of up to one wave; what it
// cannot hold is retried on the first).  Mapping of assembleKernel: one CTA per locus (persistent CTAs, global work counter).
//   * reads are packed once to 2 bits/base (MSB first, so integer order == lexicographic order A<C<G<T) plus
//     an N-mask bit plane; a k-mer (k <= 96) is three 64-bit words cut out of the packed stream;
//   * word table: open addressing keyed by the packed k-mer; the word id is the index of the window that
//     claimed the slot (one atomicCAS, no locks; only a window that tries to claim stores its key); per word a
//     read-support bitset (<= 32 words of 32 bits == <= 1024 reads incl. pseudo reads) built with one atomicOr per
//     window, so that a word repeated inside a read counts once (std::set<std::string> readWords, :524-535); count
//     and first read are derived from the bitset afterwards;
//   * successor / predecessor word ids are resolved once per word (4 + 4 lookups, all threads) so that the
//     greedy walk and the cycle search are pure integer graph walks;
//   * repeat words: homopolymers directly; other cycles are detected by parallel trimming (Kahn).  Only if a
//     cycle exists is the reference's order-dependent Tarjan replayed exactly: the iteration order of the two
//     nested std::unordered_map<std::string,...> is reproduced with a parallel formulation of libstdc++'s
//     hashtable (per rehash level: groups by bucket, ordered by descending first-arrival, members by descending
//     arrival), then one thread runs the DFS with the same index/low-link/"small circle <= 50" rules;
//   * seeds: CTA-wide maximum count, then the lexicographically smallest word among those that have it;
//   * walk: warp 0, read sets live one 32-bit word per lane, every std::set operation of the reference is a
//     lane-wise bit operation + warp vote / __reduce_add;
//   * contigs of a failed word length are re-packed as pseudo reads for the next one; selectContigs runs on
//     warp 0; results are bump-allocated into the output pools.
//
// Restrictions (documented in DESIGN.md): input bases must be in {A,C,G,T,N} (MB200_ERR_BAD_BASE otherwise),
// at most 1024 - 2*maxAssemblyCount reads per locus, word lengths 2..96, contigs up to cfg.maxContigLen.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "mb200_internal.h"

namespace {

#ifndef MB200_ASM_THREADS
#define MB200_ASM_THREADS 32
#endif
constexpr int      ASM_THREADS = MB200_ASM_THREADS;  // one warp per locus: the greedy walk is serial, so more loci in flight beat wider CTAs
constexpr uint32_t NONE32      = 0xffffffffu;
constexpr int      MAX_SETW    = 32;

// libstdc++ std::unordered_map bucket growth (see oracle/orc_assemble.c STL_BUCKETS)
__constant__ uint32_t STL_BUCKETS_D[22] = {13,     29,     59,      127,     257,     541,      1109,    2357,
                                           5087,   10273,  20753,   42043,   85229,   172933,   351061,  712697,
                                           1447153, 2938679, 5967347, 12117689, 24607243, 49969847};

struct WorkLayout {  // byte offsets inside one CTA's workspace
  size_t readLen, readPackOff, readMaskOff, winStart;
  size_t packed, nmask;
  size_t key0, key1, key2;
  size_t count, firstRead, bits, slot, wordList, succ, pred, indeg, flags, inContig;
  size_t hashv, seqA, arr, pos, bktHead, bktMin, bktCnt, chainNext, gs, orderA, tmpA;
  size_t idx, low, wstack, frame;
  size_t ctgSeq, ctgSets, ctgMeta;
  size_t total;
  uint32_t maxReads;   // incl. pseudo reads
  uint32_t nWinMax, slotsMax, packWords, maskWords, bktMax;
};

struct AsmParams {
  mb200_asm_opts       opt;
  const uint8_t*       pool;
  const uint64_t*      readOff;
  const mb200_seqview* views;
  uint32_t             nLoci;
  uint32_t*            lociCounter;
  uint8_t*             work;
  WorkLayout           L;
  uint32_t             maxContigLen;
  mb200_asm_locus*     loci;
  mb200_contig*        contigs;      // fixed slots: locus*maxAssemblyCount + c
  mb200_seqview*       pseudo;       // fixed slots: locus*2*maxAssemblyCount + p
  uint8_t*             contigPool;
  unsigned long long   contigPoolCap;
  unsigned long long*  contigPoolUsed;
  uint32_t*            setPool;
  unsigned long long   setPoolCap;
  unsigned long long*  setPoolUsed;
  int*                 errFlag;
  // list mode (second pass over the loci the on-chip kernel could not hold): process lociList[0 .. *lociListCount)
  const uint32_t*      lociList;
  const uint32_t*      lociListCount;
};

enum { FLAG_UNUSED = 1, FLAG_REPEAT = 2, FLAG_PEELED = 4, FLAG_ONSTACK = 8, FLAG_ONE_PRED = 16, FLAG_ONE_SUCC = 32 };  // the last two: on-chip kernel only

// -DMB200_ASM_PROFILE: per-phase clock64 totals and event counts of assembleKernel (tools/asm_profile.py); never in
// the product build
#ifdef MB200_ASM_PROFILE
__device__ unsigned long long gAsmProf[32];
#define PROF_DECL() long long profT = clock64()
#define PROF_MARK(i) do { const long long _t = clock64(); if (threadIdx.x == 0) atomicAdd(&gAsmProf[i], (unsigned long long)(_t - profT)); profT = _t; } while (0)
#define PROF_COUNT(i, v) do { if (threadIdx.x == 0) atomicAdd(&gAsmProf[i], (unsigned long long)(v)); } while (0)
#else
#define PROF_DECL() do { } while (0)
#define PROF_MARK(i) do { } while (0)
#define PROF_COUNT(i, v) do { } while (0)
#endif

struct CtgMeta {
  uint32_t lo, hi;  // contig == ctgSeq[lo, hi)
  int32_t  consBegin, consEnd;
};

// ---- packed sequence helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint64_t getBits(const uint64_t* s, uint32_t bitPos, uint32_t nbits)
{
  const uint32_t w = bitPos >> 6, off = bitPos & 63;
  const uint64_t hi = s[w], lo = s[w + 1];
  const uint64_t v  = off ? ((hi << off) | (lo >> (64 - off))) : hi;
  return v >> (64 - nbits);
}

struct Key {
  uint64_t k0, k1, k2;
};
__device__ __forceinline__ bool keyEq(const Key& a, const Key& b) { return a.k0 == b.k0 && a.k1 == b.k1 && a.k2 == b.k2; }
__device__ __forceinline__ bool keyLess(const Key& a, const Key& b)
{
  if (a.k2 != b.k2) return a.k2 < b.k2;
  if (a.k1 != b.k1) return a.k1 < b.k1;
  return a.k0 < b.k0;
}
__device__ __forceinline__ uint64_t lowMask(uint32_t nbits) { return nbits >= 64 ? ~0ull : ((1ull << nbits) - 1); }

__device__ __forceinline__ Key extractKey(const uint64_t* s, uint32_t base, uint32_t k)
{
  const uint32_t B = 2 * k, bp = 2 * base;
  Key            r;
  if (B <= 64) {
    r.k0 = getBits(s, bp, B);
    r.k1 = r.k2 = 0;
  } else if (B <= 128) {
    r.k1 = getBits(s, bp, B - 64);
    r.k0 = getBits(s, bp + B - 64, 64);
    r.k2 = 0;
  } else {
    r.k2 = getBits(s, bp, B - 128);
    r.k1 = getBits(s, bp + B - 128, 64);
    r.k0 = getBits(s, bp + B - 64, 64);
  }
  return r;
}
__device__ __forceinline__ Key succKey(const Key& a, uint32_t sym, uint32_t k)
{
  const uint32_t B = 2 * k;
  Key            r;
  r.k2 = (a.k2 << 2) | (a.k1 >> 62);
  r.k1 = (a.k1 << 2) | (a.k0 >> 62);
  r.k0 = (a.k0 << 2) | sym;
  if (B <= 64) {
    r.k0 &= lowMask(B);
    r.k1 = r.k2 = 0;
  } else if (B <= 128) {
    r.k1 &= lowMask(B - 64);
    r.k2 = 0;
  } else {
    r.k2 &= lowMask(B - 128);
  }
  return r;
}
__device__ __forceinline__ Key predKey(const Key& a, uint32_t sym, uint32_t k)
{
  const uint32_t B = 2 * k;
  Key            r;
  r.k0 = (a.k0 >> 2) | (a.k1 << 62);
  r.k1 = (a.k1 >> 2) | (a.k2 << 62);
  r.k2 = a.k2 >> 2;
  if (B <= 64) r.k0 |= (uint64_t)sym << (B - 2);
  else if (B <= 128) r.k1 |= (uint64_t)sym << (B - 66);
  else r.k2 |= (uint64_t)sym << (B - 130);
  return r;
}
__device__ __forceinline__ uint32_t baseAt(const Key& a, uint32_t i, uint32_t k)
{  // base i (0 = first) of the word
  const uint32_t pos = 2 * (k - 1 - i);
  const uint64_t w   = pos < 64 ? a.k0 : (pos < 128 ? a.k1 : a.k2);
  return (uint32_t)(w >> (pos & 63)) & 3u;
}
__device__ __forceinline__ uint32_t hashKey(const Key& a)
{
  uint64_t h = a.k0 * 0x9E3779B97F4A7C15ull;
  h ^= (a.k1 + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
  h ^= (a.k2 + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (uint32_t)h;
}

// std::hash<std::string> of the ASCII spelling of a word (libstdc++ _Hash_bytes, see oracle/orc_assemble.c)
__device__ uint64_t stlHashWord(const Key& a, uint32_t k)
{
  const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
  uint64_t       h   = 0xc70f6907ull ^ ((uint64_t)k * mul);
  const uint32_t alen = k & ~7u;
  for (uint32_t p = 0; p < alen; p += 8) {
    uint64_t d = 0;
    for (uint32_t t = 0; t < 8; ++t) d |= (uint64_t)(uint8_t)("ACGT"[baseAt(a, p + t, k)]) << (8 * t);
    d *= mul;
    d ^= d >> 47;
    d *= mul;
    h ^= d;
    h *= mul;
  }
  if (k & 7) {
    uint64_t d = 0;
    for (uint32_t t = 0; t < (k & 7); ++t) d |= (uint64_t)(uint8_t)("ACGT"[baseAt(a, alen + t, k)]) << (8 * t);
    h ^= d;
    h *= mul;
  }
  h ^= h >> 47;
  h *= mul;
  h ^= h >> 47;
  return h;
}

// bit j of x -> bit 2j (Morton spread), used to interleave the two bit planes of 32 packed bases
__device__ __forceinline__ uint64_t spread32(uint32_t x)
{
  uint64_t v = x;
  v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
  v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
  v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
  v = (v | (v << 2)) & 0x3333333333333333ull;
  v = (v | (v << 1)) & 0x5555555555555555ull;
  return v;
}

// Pack one read: 2 bits/base MSB-first (32 bases per word, nPk words incl. 4 zero words of padding) and the N-mask
// bit plane (64 bases per word, nMk words).  One lane per base: three ballots per 32 bases give the two code bit
// planes and the mask; lane c of a group of 32 chunks then assembles word c.  Returns true if a byte outside
// {A,C,G,T,N} was seen (only checked when `checkAlphabet`).
__device__ bool packRead(const uint8_t* src, uint32_t len, uint64_t* pk, uint64_t* mk, uint32_t nPk, uint32_t nMk, bool checkAlphabet,
                         int lane)
{
  bool           bad     = false;
  const uint32_t nChunks = (len + 31) / 32;
  uint32_t       maskDone = 0;
  for (uint32_t c0 = 0; c0 < nPk; c0 += 32) {
    uint32_t       myHi = 0, myLo = 0, myN = 0;
    const uint32_t cEnd = min(nChunks, c0 + 32);
    for (uint32_t c = c0; c < cEnd; ++c) {
      const uint32_t i    = c * 32 + lane;
      uint32_t       code = 0;
      bool           isN  = false;
      if (i < len) {
        const uint8_t ch = src[i];
        code             = (ch == 'C') ? 1u : ((ch == 'G') ? 2u : ((ch == 'T') ? 3u : 0u));
        isN              = (ch != 'A') && (code == 0);
        if (checkAlphabet && isN && ch != 'N') bad = true;
      }
      const uint32_t hi = __ballot_sync(0xffffffffu, (code & 2u) != 0);
      const uint32_t lo = __ballot_sync(0xffffffffu, (code & 1u) != 0);
      const uint32_t nn = __ballot_sync(0xffffffffu, isN);
      if ((uint32_t)lane == c - c0) { myHi = hi; myLo = lo; myN = nn; }
    }
    const uint32_t w = c0 + lane;
    if (w < nPk) pk[w] = (spread32(__brev(myHi)) << 1) | spread32(__brev(myLo));
    const uint32_t nA = __shfl_sync(0xffffffffu, myN, (2 * lane) & 31);
    const uint32_t nB = __shfl_sync(0xffffffffu, myN, (2 * lane + 1) & 31);
    const uint32_t wm = c0 / 2 + lane;
    if (lane < 16 && wm < nMk) mk[wm] = ((uint64_t)__brev(nA) << 32) | (uint64_t)__brev(nB);
    maskDone = c0 / 2 + 16;
  }
  for (uint32_t wm = maskDone + lane; wm < nMk; wm += 32) mk[wm] = 0;
  return __any_sync(0xffffffffu, bad);
}

// ---- per-CTA view of the workspace -----------------------------------------------------------------
struct Work {
  uint32_t *readLen, *readPackOff, *readMaskOff, *winStart;
  uint64_t *packed, *nmask;
  uint64_t *key0, *key1, *key2;  // key of a word, at the index of the window that claimed its slot
  uint32_t *count, *firstRead, *bits, *slot, *wordList, *succ, *pred, *indeg;
  uint8_t * flags, *inContig;
  uint64_t* hashv;
  uint32_t *seqA, *arr, *pos, *bktHead, *bktMin, *bktCnt, *chainNext, *gs, *orderA, *tmpA;
  uint32_t *idx, *low, *wstack, *frame;
  uint8_t*  ctgSeq;
  uint32_t* ctgSets;
  CtgMeta*  ctgMeta;
};

__device__ __forceinline__ Key loadKey(const Work& W, uint32_t id, uint32_t /*k*/)
{
  Key k;
  k.k0 = W.key0[id];
  k.k1 = W.key1[id];
  k.k2 = W.key2[id];
  return k;
}

__device__ uint32_t findWord(const Work& W, const Key& key, uint32_t k, uint32_t slotMask)
{
  uint32_t h = hashKey(key) & slotMask;
  for (;;) {
    const uint32_t cur = W.slot[h];
    if (cur == 0) return NONE32;
    if (keyEq(loadKey(W, cur - 1, k), key)) return cur - 1;
    h = (h + 1) & slotMask;
  }
}

// exclusive SUFFIX sum of a[0..m) into out[0..m): out[i] = sum_{j>i} a[j]   (whole CTA, uses sh[ASM_THREADS])
__device__ void suffixSumExclusive(const uint32_t* a, uint32_t* out, uint32_t m, uint32_t* sh)
{
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t per = (m + nt - 1) / nt;
  // thread t owns the chunk [m - (t+1)*per, m - t*per) counted from the END
  const long long hiL = (long long)m - (long long)tid * per;
  const long long loL = hiL - per;
  const uint32_t  hi  = hiL > 0 ? (uint32_t)hiL : 0;
  const uint32_t  lo  = loL > 0 ? (uint32_t)loL : 0;
  uint32_t        sum = 0;
  for (uint32_t i = lo; i < hi; ++i) sum += a[i];
  sh[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t t = 0; t < nt; ++t) {
      const uint32_t v = sh[t];
      sh[t]            = run;  // sum of everything after this thread's chunk
      run += v;
    }
  }
  __syncthreads();
  uint32_t run = sh[tid];
  for (uint32_t i = hi; i > lo; --i) {
    out[i - 1] = run;
    run += a[i - 1];
  }
  __syncthreads();
}

// Iteration order of a libstdc++ std::unordered_map<std::string,T> filled by operator[] with n keys whose
// std::hash values are hashIn[0..n) (insertion order).  order[p] = insertion index of the p-th iterated element.
__device__ void unorderedOrder(const Work& W, const uint64_t* hashIn, uint32_t n, uint32_t* order, uint32_t* sh)
{
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  uint32_t       prevB = 0;
  for (int level = 0; level < 22; ++level) {
    const uint32_t B = STL_BUCKETS_D[level];
    const uint32_t m = n < B ? n : B;
    for (uint32_t b = tid; b < B; b += nt) {
      W.bktHead[b] = NONE32;
      W.bktMin[b]  = NONE32;
      W.bktCnt[b]  = 0;
    }
    for (uint32_t e = tid; e < m; e += nt) {
      W.gs[e] = 0;
      // arrival time: position in the previous level's list for the old elements, own index for the new ones
      W.arr[e] = (e < prevB) ? W.pos[e] : e;
    }
    __syncthreads();
    for (uint32_t e = tid; e < m; e += nt) {
      const uint32_t b = (uint32_t)(hashIn[e] % B);
      W.chainNext[e]   = atomicExch(&W.bktHead[b], e);
      atomicMin(&W.bktMin[b], W.arr[e]);
      atomicAdd(&W.bktCnt[b], 1u);
    }
    __syncthreads();
    for (uint32_t e = tid; e < m; e += nt) {
      const uint32_t b = (uint32_t)(hashIn[e] % B);
      if (W.arr[e] == W.bktMin[b]) W.gs[W.arr[e]] = W.bktCnt[b];
    }
    __syncthreads();
    suffixSumExclusive(W.gs, W.tmpA, m, sh);  // tmpA[a] = number of elements in groups created after arrival a
    for (uint32_t e = tid; e < m; e += nt) {
      const uint32_t b = (uint32_t)(hashIn[e] % B);
      const uint32_t a = W.arr[e];
      uint32_t       rank = 0;
      for (uint32_t x = W.bktHead[b]; x != NONE32; x = W.chainNext[x])
        if (W.arr[x] > a) rank++;
      W.orderA[e] = W.tmpA[W.bktMin[b]] + rank;
    }
    __syncthreads();
    for (uint32_t e = tid; e < m; e += nt) W.pos[e] = W.orderA[e];
    __syncthreads();
    prevB = B;
    if (n <= B) break;
  }
  for (uint32_t e = tid; e < n; e += nt) order[W.pos[e]] = e;
  __syncthreads();
}

// ======================================================================================================
__global__ void __launch_bounds__(ASM_THREADS, 32) assembleKernel(const AsmParams P)
{
  __shared__ uint32_t sh[ASM_THREADS < 64 ? 64 : ASM_THREADS];
  __shared__ uint32_t shU[16];

  const uint32_t tid  = threadIdx.x;
  const uint32_t nt   = blockDim.x;
  const uint32_t lane = tid & 31;
  const uint32_t warp = tid >> 5;

  uint8_t* base = P.work + (size_t)blockIdx.x * P.L.total;
  Work     W;
  W.readLen     = reinterpret_cast<uint32_t*>(base + P.L.readLen);
  W.readPackOff = reinterpret_cast<uint32_t*>(base + P.L.readPackOff);
  W.readMaskOff = reinterpret_cast<uint32_t*>(base + P.L.readMaskOff);
  W.winStart    = reinterpret_cast<uint32_t*>(base + P.L.winStart);
  W.packed      = reinterpret_cast<uint64_t*>(base + P.L.packed);
  W.nmask       = reinterpret_cast<uint64_t*>(base + P.L.nmask);
  W.key0        = reinterpret_cast<uint64_t*>(base + P.L.key0);
  W.key1        = reinterpret_cast<uint64_t*>(base + P.L.key1);
  W.key2        = reinterpret_cast<uint64_t*>(base + P.L.key2);
  W.count       = reinterpret_cast<uint32_t*>(base + P.L.count);
  W.firstRead   = reinterpret_cast<uint32_t*>(base + P.L.firstRead);
  W.bits        = reinterpret_cast<uint32_t*>(base + P.L.bits);
  W.slot        = reinterpret_cast<uint32_t*>(base + P.L.slot);
  W.wordList    = reinterpret_cast<uint32_t*>(base + P.L.wordList);
  W.succ        = reinterpret_cast<uint32_t*>(base + P.L.succ);
  W.pred        = reinterpret_cast<uint32_t*>(base + P.L.pred);
  W.indeg       = reinterpret_cast<uint32_t*>(base + P.L.indeg);
  W.flags       = reinterpret_cast<uint8_t*>(base + P.L.flags);
  W.inContig    = reinterpret_cast<uint8_t*>(base + P.L.inContig);
  W.hashv       = reinterpret_cast<uint64_t*>(base + P.L.hashv);
  W.seqA        = reinterpret_cast<uint32_t*>(base + P.L.seqA);
  W.arr         = reinterpret_cast<uint32_t*>(base + P.L.arr);
  W.pos         = reinterpret_cast<uint32_t*>(base + P.L.pos);
  W.bktHead     = reinterpret_cast<uint32_t*>(base + P.L.bktHead);
  W.bktMin      = reinterpret_cast<uint32_t*>(base + P.L.bktMin);
  W.bktCnt      = reinterpret_cast<uint32_t*>(base + P.L.bktCnt);
  W.chainNext   = reinterpret_cast<uint32_t*>(base + P.L.chainNext);
  W.gs          = reinterpret_cast<uint32_t*>(base + P.L.gs);
  W.orderA      = reinterpret_cast<uint32_t*>(base + P.L.orderA);
  W.tmpA        = reinterpret_cast<uint32_t*>(base + P.L.tmpA);
  W.idx         = reinterpret_cast<uint32_t*>(base + P.L.idx);
  W.low         = reinterpret_cast<uint32_t*>(base + P.L.low);
  W.wstack      = reinterpret_cast<uint32_t*>(base + P.L.wstack);
  W.frame       = reinterpret_cast<uint32_t*>(base + P.L.frame);
  W.ctgSeq      = base + P.L.ctgSeq;
  W.ctgSets     = reinterpret_cast<uint32_t*>(base + P.L.ctgSets);
  W.ctgMeta     = reinterpret_cast<CtgMeta*>(base + P.L.ctgMeta);

  const mb200_asm_opts opt     = P.opt;
  const uint32_t       maxIter = 2 * opt.maxAssemblyCount;  // contigs per word length
  const uint32_t       ctgBuf  = 2 * P.maxContigLen;        // bytes per contig buffer

  const uint32_t lociTotal = P.lociListCount ? *P.lociListCount : P.nLoci;
  for (;;) {
    if (tid == 0) {
      const uint32_t idx = atomicAdd(P.lociCounter, 1u);
      shU[0]             = (idx < lociTotal) ? (P.lociList ? P.lociList[idx] : idx) : NONE32;
    }
    __syncthreads();
    const uint32_t li = shU[0];
    __syncthreads();
    if (li == NONE32) break;

    PROF_DECL();
    PROF_COUNT(19, 1);
    const uint64_t r0              = P.readOff[li];
    const uint32_t normalReadCount = (uint32_t)(P.readOff[li + 1] - r0);
    const uint32_t SW              = (normalReadCount + maxIter + 31) / 32;  // set words of this locus
    int            status          = MB200_OK;

    // ---------------- pack the normal reads (2 bits/base MSB-first + N mask)
    if (tid == 0) {
      uint32_t po = 0, mo = 0;
      for (uint32_t r = 0; r < normalReadCount; ++r) {
        const uint32_t len = P.views[r0 + r].len;
        W.readLen[r]       = len;
        W.readPackOff[r]   = po;
        W.readMaskOff[r]   = mo;
        po += (len + 31) / 32 + 4;
        mo += (len + 63) / 64 + 3;
      }
      shU[1] = po;
      shU[2] = mo;
    }
    __syncthreads();
    const uint32_t normPackWords = shU[1], normMaskWords = shU[2];
    bool           badBase       = false;
    for (uint32_t r = warp; r < normalReadCount; r += nt / 32) {
      const uint8_t* src = P.pool + P.views[r0 + r].off;
      const uint32_t len = W.readLen[r];
      uint64_t*      pk  = W.packed + W.readPackOff[r];
      uint64_t*      mk  = W.nmask + W.readMaskOff[r];
      const uint32_t nPk = (len + 31) / 32 + 4, nMk = (len + 63) / 64 + 3;
      if (packRead(src, len, pk, mk, nPk, nMk, true, lane)) badBase = true;
    }
    if (__syncthreads_or(badBase ? 1 : 0)) status = MB200_ERR_BAD_BASE;
    if (normalReadCount + maxIter > 32 * MAX_SETW) status = MB200_ERR_UNSUPPORTED;

    uint32_t nReads   = normalReadCount;  // normal + pseudo
    uint32_t nCtg     = 0;                // contigs of the last buildContigs call
    uint32_t lastK    = 0;

    if (status == MB200_OK) {
      PROF_MARK(0);
      for (uint32_t k = opt.minWordLength; k <= opt.maxWordLength; k += opt.wordStepSize) {
        lastK = k;
        PROF_COUNT(16, 1);
        // ============================ buildContigs(k) ============================
        // ---- windows
        if (tid == 0) {
          uint32_t acc = 0;
          for (uint32_t r = 0; r < nReads; ++r) {
            W.winStart[r] = acc;
            const uint32_t len = W.readLen[r];
            if (len >= k) acc += len - k + 1;
          }
          W.winStart[nReads] = acc;
          shU[3]             = acc;
        }
        __syncthreads();
        const uint32_t nWin = shU[3];
        uint32_t       slots = 64;
        while (slots < 2 * nWin) slots <<= 1;
        const uint32_t slotMask = slots - 1;
        for (uint32_t i = tid; i < slots; i += nt) W.slot[i] = 0;
        for (uint32_t i = tid; i < nWin; i += nt) {
          W.count[i]     = 0;
          W.firstRead[i] = NONE32;
          W.flags[i]     = 0;
          W.inContig[i]  = 0;
          W.indeg[i]     = 0;
        }
        for (uint32_t i = tid; i < nWin * SW; i += nt) W.bits[i] = 0;
        if (tid == 0) shU[4] = 0;  // nWords
        __syncthreads();
        PROF_MARK(1);
        PROF_COUNT(21, nWin);

        // ---- getKmerCounts: every N-free window claims or finds the slot of its word and sets its read's support bit
        // (std::set<std::string> readWords: a word repeated inside a read counts once, :524-535)
        PROF_MARK(2);
        for (uint32_t r = warp; r < nReads; r += nt / 32) {
          const uint32_t len = W.readLen[r];
          if (len < k) continue;
          const uint64_t* pk = W.packed + W.readPackOff[r];
          const uint64_t* mk = W.nmask + W.readMaskOff[r];
          const uint32_t  ws = W.winStart[r];
          const uint32_t  bit = 1u << (r & 31);
          for (uint32_t j = lane; j + k <= len; j += 32) {
            const uint32_t w = ws + j;
            bool hasN = getBits(mk, j, k < 64 ? k : 64) != 0;
            if (k > 64) hasN = hasN || (getBits(mk, j + 64, k - 64) != 0);
            if (hasN) continue;
            const Key key = extractKey(pk, j, k);
            uint32_t  h   = hashKey(key) & slotMask;
            uint32_t  owner;
            bool      stored = false;
            for (;;) {
              uint32_t cur = W.slot[h];
              if (cur == 0) {
                if (!stored) {  // only a window that may become the word's name stores its key (before the slot is published)
                  W.key0[w] = key.k0;
                  W.key1[w] = key.k1;
                  W.key2[w] = key.k2;
                  __threadfence_block();
                  stored = true;
                }
                cur = atomicCAS(&W.slot[h], 0u, w + 1);
                if (cur == 0) {
                  owner = w;
                  W.wordList[atomicAdd(&shU[4], 1u)] = w;
                  break;
                }
              }
              if (keyEq(loadKey(W, cur - 1, k), key)) {
                owner = cur - 1;
                break;
              }
              h = (h + 1) & slotMask;
            }
            atomicOr(&W.bits[(size_t)owner * SW + (r >> 5)], bit);
          }
        }
        __syncthreads();
        const uint32_t nWords = shU[4];
        // counts (1 per normal read, minCoverage per pseudo read, :541) and first reads, from the bitsets
        for (uint32_t i = tid; i < nWords; i += nt) {
          const uint32_t id = W.wordList[i];
          uint32_t       c = 0, first = NONE32;
          for (uint32_t sw = 0; sw < SW; ++sw) {
            const uint32_t b2  = W.bits[(size_t)id * SW + sw];
            const uint32_t lo2 = sw * 32;
            uint32_t       pm;  // pseudo-read mask of this set word
            if (normalReadCount <= lo2) pm = 0xffffffffu;
            else if (normalReadCount >= lo2 + 32) pm = 0;
            else pm = ~((1u << (normalReadCount - lo2)) - 1);
            c += (uint32_t)__popc(b2 & ~pm) + opt.minCoverage * (uint32_t)__popc(b2 & pm);
            if (b2 && first == NONE32) first = lo2 + (uint32_t)__ffs((int)b2) - 1;
          }
          W.count[id]     = c;
          W.firstRead[id] = first;
        }
        __syncthreads();
        PROF_MARK(3);
        PROF_COUNT(20, nWords);

        // ---- successor / predecessor ids, homopolymer repeats, in-degrees for the trimming pass
        for (uint32_t i = tid; i < nWords * 4; i += nt) {
          const uint32_t id  = W.wordList[i >> 2];
          const uint32_t sym = i & 3;
          const Key      key = loadKey(W, id, k);
          const uint32_t s   = findWord(W, succKey(key, sym, k), k, slotMask);
          W.succ[(size_t)id * 4 + sym] = s;
          W.pred[(size_t)id * 4 + sym] = findWord(W, predKey(key, sym, k), k, slotMask);
          if (s == id) {
            atomicOr(reinterpret_cast<unsigned int*>(W.flags + (id & ~3u)), (unsigned)FLAG_REPEAT << (8 * (id & 3)));
          } else if (s != NONE32) {
            atomicAdd(&W.indeg[s], 1u);
          }
        }
        __syncthreads();
        PROF_MARK(4);

        // ---- is there any cycle besides homopolymer self loops?  (Kahn trimming, level synchronous)
        {
          uint32_t* queue = W.seqA;  // reuse
          if (tid == 0) { shU[5] = 0; shU[6] = 0; }
          __syncthreads();
          for (uint32_t i = tid; i < nWords; i += nt) {
            const uint32_t id = W.wordList[i];
            if (W.indeg[id] == 0) queue[atomicAdd(&shU[5], 1u)] = id;
          }
          __syncthreads();
          uint32_t head = 0;
          for (;;) {
            const uint32_t tail = shU[5];
            if (head == tail) break;
            __syncthreads();
            for (uint32_t i = head + tid; i < tail; i += nt) {
              const uint32_t id = queue[i];
              const uint4    sc = *reinterpret_cast<const uint4*>(W.succ + (size_t)id * 4);
              const uint32_t ss[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint32_t nx = ss[s];
                if (nx == NONE32 || nx == id) continue;
                if (atomicSub(&W.indeg[nx], 1u) == 1u) queue[atomicAdd(&shU[5], 1u)] = nx;
              }
            }
            head = tail;
            __syncthreads();
          }
          // shU[5] == number of peeled words
        }
        const bool hasCycle = (shU[5] != nWords);
        __syncthreads();
        PROF_MARK(5);
        PROF_COUNT(22, hasCycle ? 1 : 0);

        if (hasCycle) {
          // ================= exact replay of getRepeatKmers (hash-order DFS roots) =================
          // (a) wordCount insertion order: by first read, then lexicographically inside the read
          uint32_t* bucketStart = W.tmpA;   // [nReads+1]
          uint32_t* members     = W.orderA; // words grouped by first read (unordered inside the group)
          for (uint32_t r = tid; r <= nReads; r += nt) W.gs[r] = 0;
          __syncthreads();
          for (uint32_t i = tid; i < nWords; i += nt) atomicAdd(&W.gs[W.firstRead[W.wordList[i]]], 1u);
          __syncthreads();
          if (tid == 0) {
            uint32_t acc = 0;
            for (uint32_t r = 0; r < nReads; ++r) {
              bucketStart[r] = acc;
              acc += W.gs[r];
              W.gs[r] = 0;
            }
            bucketStart[nReads] = acc;
          }
          __syncthreads();
          for (uint32_t i = tid; i < nWords; i += nt) {
            const uint32_t id = W.wordList[i];
            const uint32_t r  = W.firstRead[id];
            members[bucketStart[r] + atomicAdd(&W.gs[r], 1u)] = id;
          }
          __syncthreads();
          for (uint32_t i = tid; i < nWords; i += nt) {
            const uint32_t id  = W.wordList[i];
            const uint32_t r   = W.firstRead[id];
            const Key      key = loadKey(W, id, k);
            uint32_t       rank = 0;
            for (uint32_t m = bucketStart[r]; m < bucketStart[r + 1]; ++m)
              if (keyLess(loadKey(W, members[m], k), key)) rank++;
            const uint32_t p = bucketStart[r] + rank;
            W.seqA[p]        = id;                 // insertion sequence of wordCount
            W.hashv[p]       = stlHashWord(key, k);
          }
          __syncthreads();
          // (b) iteration order of wordCount, then of wordIndices (filled while iterating wordCount)
          uint32_t* order1 = W.idx;  // temporarily
          unorderedOrder(W, W.hashv, nWords, order1, sh);
          uint64_t* hash2 = reinterpret_cast<uint64_t*>(W.wstack);  // wstack+frame are contiguous uint32[nWin] x2
          for (uint32_t i = tid; i < nWords; i += nt) hash2[i] = W.hashv[order1[i]];
          __syncthreads();
          uint32_t* order2 = W.low;
          unorderedOrder(W, hash2, nWords, order2, sh);
          // root sequence: word id of the i-th element of wordIndices
          for (uint32_t i = tid; i < nWords; i += nt) W.chainNext[i] = W.seqA[order1[order2[i]]];
          __syncthreads();
          for (uint32_t i = tid; i < nWords; i += nt) {
            const uint32_t id = W.wordList[i];
            W.idx[id] = 0;
            W.low[id] = 0;
          }
          __syncthreads();
          // (c) searchRepeats, recursion unrolled (one thread; the DFS order is inherently sequential)
          if (tid == 0) {
            const uint32_t* roots  = W.chainNext;
            uint32_t        index  = 1;
            uint32_t        nStack = 0;
            for (uint32_t ri = 0; ri < nWords; ++ri) {
              const uint32_t root = roots[ri];
              if (W.idx[root] != 0) continue;
              uint32_t depth = 1;
              W.frame[0]     = root << 3;  // (word << 3) | next symbol
              W.idx[root] = W.low[root] = index++;
              W.wstack[nStack++]        = root;
              W.flags[root] |= FLAG_ONSTACK;
              while (depth) {
                const uint32_t fr  = W.frame[depth - 1];
                const uint32_t w   = fr >> 3;
                const uint32_t sym = fr & 7;
                if (sym < 4) {
                  W.frame[depth - 1] = fr + 1;
                  const uint32_t nw  = W.succ[(size_t)w * 4 + sym];
                  if (nw == w) continue;  // homopolymer: already marked
                  if (nw == NONE32) continue;
                  if (W.idx[nw] == 0) {
                    W.frame[depth++] = nw << 3;
                    W.idx[nw] = W.low[nw] = index++;
                    W.wstack[nStack++]    = nw;
                    W.flags[nw] |= FLAG_ONSTACK;
                  } else if (W.flags[nw] & FLAG_ONSTACK) {
                    if (W.idx[nw] < W.low[w]) W.low[w] = W.idx[nw];
                  }
                } else {
                  if (W.low[w] == W.idx[w]) {
                    const uint32_t lastWord = W.wstack[nStack - 1];
                    if (lastWord == w) {
                      nStack--;
                      W.flags[w] &= ~FLAG_ONSTACK;
                    } else {
                      const bool isSmallCircle = (W.idx[lastWord] - W.idx[w]) <= 50;
                      for (;;) {
                        const uint32_t rw = W.wstack[--nStack];
                        W.flags[rw] &= ~FLAG_ONSTACK;
                        if (isSmallCircle) W.flags[rw] |= FLAG_REPEAT;
                        if (rw == w) break;
                      }
                    }
                  }
                  depth--;
                  if (depth) {
                    const uint32_t parent = W.frame[depth - 1] >> 3;
                    if (W.low[w] < W.low[parent]) W.low[parent] = W.low[w];
                  }
                }
              }
            }
          }
          __syncthreads();
        }

        PROF_MARK(6);
        // ---- unusedWords = {count >= minCoverage}
        for (uint32_t i = tid; i < nWords; i += nt) {
          const uint32_t id = W.wordList[i];
          if (W.count[id] >= opt.minCoverage) W.flags[id] |= FLAG_UNUSED;
        }
        __syncthreads();

        // ---- contig loop (:685-716)
        bool isAssemblySuccess = true;
        nCtg                   = 0;
        while (nCtg < maxIter) {
          // seed: max count, ties -> lexicographically smallest word (first in the std::set scan, :686-696).
          // Stage 1: the maximum count; stage 2: keys only of the words that have it.
          uint32_t enc = 0;  // count + 1 of the best unused word seen, 0 = none
          for (uint32_t i = tid; i < nWords; i += nt) {
            const uint32_t id = W.wordList[i];
            if (!(W.flags[id] & FLAG_UNUSED)) continue;
            const uint32_t c = W.count[id];
            const uint32_t e = c >= 0xfffffffeu ? 0xffffffffu : c + 1;
            if (e > enc) enc = e;
          }
          enc = __reduce_max_sync(0xffffffffu, enc);
          if (lane == 0) sh[warp] = enc;
          __syncthreads();
          uint32_t encAll = 0;
          for (uint32_t w = 0; w < nt / 32; ++w) encAll = max(encAll, sh[w]);
          __syncthreads();
          Key      bestKey = {~0ull, ~0ull, ~0ull};
          uint32_t bestId  = NONE32;
          if (encAll != 0) {
            for (uint32_t i = tid; i < nWords; i += nt) {
              const uint32_t id = W.wordList[i];
              if (!(W.flags[id] & FLAG_UNUSED)) continue;
              const uint32_t c = W.count[id];
              if ((c >= 0xfffffffeu ? 0xffffffffu : c + 1) != encAll) continue;
              const Key key = loadKey(W, id, k);
              if (bestId == NONE32 || keyLess(key, bestKey)) {
                bestKey = key;
                bestId  = id;
              }
            }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const uint32_t oid = __shfl_xor_sync(0xffffffffu, bestId, off);
            Key            ok;
            ok.k0 = __shfl_xor_sync(0xffffffffu, bestKey.k0, off);
            ok.k1 = __shfl_xor_sync(0xffffffffu, bestKey.k1, off);
            ok.k2 = __shfl_xor_sync(0xffffffffu, bestKey.k2, off);
            if (oid != NONE32 && (bestId == NONE32 || keyLess(ok, bestKey))) {
              bestKey = ok;
              bestId  = oid;
            }
          }
          if (lane == 0) sh[warp * 8 + 0] = bestId;
          __syncthreads();
          if (tid == 0) {
            uint32_t bid = NONE32;
            Key      bk  = {0, 0, 0};
            for (uint32_t w = 0; w < nt / 32; ++w) {
              const uint32_t id = sh[w * 8 + 0];
              if (id == NONE32) continue;
              const Key kk = loadKey(W, id, k);
              if (bid == NONE32 || keyLess(kk, bk)) {
                bid = id;
                bk  = kk;
              }
            }
            shU[7] = bid;
          }
          __syncthreads();
          const uint32_t seed = shU[7];
          __syncthreads();
          PROF_MARK(7);
          if (seed == NONE32) break;  // unusedWords.empty()
          PROF_COUNT(17, 1);

          // ---------------- walk(seed)  (warp 0; IterativeAssembler.cpp:149-501)
          if (warp == 0) {
            uint8_t*       seq     = W.ctgSeq + (size_t)nCtg * ctgBuf;
            uint32_t*      sets    = W.ctgSets + (size_t)nCtg * 2 * MAX_SETW;
            const uint32_t serial  = nCtg + 1;
            const Key      seedKey = loadKey(W, seed, k);
            uint32_t       lo = P.maxContigLen - k / 2, hi;
            hi = lo + k;
            for (uint32_t i = lane; i < k; i += 32) seq[lo + i] = "ACGT"[baseAt(seedKey, i, k)];
            uint32_t support = (lane < SW) ? W.bits[(size_t)seed * SW + lane] : 0u;
            uint32_t reject  = 0;
            if (lane == 0) W.flags[seed] &= ~FLAG_UNUSED;
            int  consBegin = 0, consEnd = 0;
            bool isRepeatFound = false;
            bool overflow      = false;
            if (W.flags[seed] & FLAG_REPEAT) {
              consBegin     = 0;
              consEnd       = (int)k;
              isRepeatFound = true;
            } else {
              if (lane == 0) W.inContig[seed] = (uint8_t)serial;
              // rejecting reads from the seed's sibling words (same k-1 prefix)  :185-210
              {
                const uint32_t lastSym = (uint32_t)seedKey.k0 & 3u;
                Key            trunkK  = seedKey;
                trunkK.k0 &= ~3ull;
                for (uint32_t s = 0; s < 4; ++s) {
                  if (s == lastSym) continue;
                  Key kk = trunkK;
                  kk.k0 |= s;
                  const uint32_t id = findWord(W, kk, k, slotMask);
                  if (id == NONE32) continue;
                  if (lane < SW) reject |= W.bits[(size_t)id * SW + lane];
                }
              }
              __syncwarp();
              uint32_t curWord = seed;
              for (int mode = 0; mode < 2; ++mode) {
                const bool isEnd = (mode == 0);
                uint32_t   conservativeEndOffset = 0;
                curWord = seed;  // mode 0: last word of the contig == seed; mode 1: first word == seed
                for (;;) {
                  const uint4 cand4 = *reinterpret_cast<const uint4*>((isEnd ? W.succ : W.pred) + (size_t)curWord * 4);
                  const uint32_t cand[4] = {cand4.x, cand4.y, cand4.z, cand4.w};
                  uint32_t maxBaseCount = 0, maxCnt = 0, maxBase = 0, maxWord = NONE32;
                  uint32_t maxW = 0, maxCW = 0, sup2Remove = 0, rej2Add = 0;
#pragma unroll
                  for (int s = 0; s < 4; ++s) {
                    const uint32_t id = cand[s];
                    if (id == NONE32) continue;
                    const uint32_t cw         = (lane < SW) ? W.bits[(size_t)id * SW + lane] : 0u;
                    const uint32_t contigWord = support & cw;
                    const uint32_t shared     = maxCW & cw;
                    if (!__any_sync(0xffffffffu, contigWord != 0)) continue;
                    const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(contigWord));
                    if (c > maxCnt) {
                      sup2Remove |= maxCW & ~shared;
                      rej2Add |= maxW & ~shared;
                      maxW         = cw;
                      maxCnt       = c;
                      maxCW        = contigWord;
                      maxBaseCount = W.count[id];
                      maxBase      = (uint32_t)s;
                      maxWord      = id;
                    } else {
                      sup2Remove |= contigWord & ~shared;
                      rej2Add |= cw & ~shared;
                    }
                  }
                  if (maxBaseCount < opt.minCoverage) break;
                  if (W.inContig[maxWord] == (uint8_t)serial) {
                    isRepeatFound = true;
                    break;
                  }
                  // walk backwards one step at the branching point (:375-427); previousWordReads is always empty
                  // (the symbol the reference skips is the one that leads back to the current word; links are built from
                  // the same table, so that neighbour IS the current word and the test is on the id -- no key load)
                  if (__any_sync(0xffffffffu, maxW != 0)) {
                    const uint4 alt4 = *reinterpret_cast<const uint4*>((isEnd ? W.pred : W.succ) + (size_t)maxWord * 4);
                    const uint32_t alt[4] = {alt4.x, alt4.y, alt4.z, alt4.w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                      const uint32_t id = alt[s];
                      if (id == NONE32 || id == maxWord || id == curWord) continue;
                      const uint32_t back     = (lane < SW) ? W.bits[(size_t)id * SW + lane] : 0u;
                      const uint32_t toUpdate = back & ~(maxCW & back);
                      rej2Add |= toUpdate;
                      sup2Remove |= toUpdate;
                    }
                  }
                  // contig.seq = addBase(contig.seq, maxBase, isEnd)
                  if (hi - lo + 1 > P.maxContigLen || lo == 0 || hi + 1 >= ctgBuf) {
                    overflow = true;
                    break;
                  }
                  if (lane == 0) {
                    if (isEnd) seq[hi] = "ACGT"[maxBase];
                    else seq[lo - 1] = "ACGT"[maxBase];
                  }
                  if (isEnd) hi++; else lo--;
                  if ((conservativeEndOffset != 0) || (maxBaseCount < opt.minConservativeCoverage)) conservativeEndOffset += 1;
                  reject |= rej2Add;
                  support |= maxW & ~reject;
                  support &= ~sup2Remove;
                  if (lane == 0) {
                    W.flags[maxWord] &= ~FLAG_UNUSED;
                    W.inContig[maxWord] = (uint8_t)serial;
                  }
                  __syncwarp();
                  curWord = maxWord;
                  PROF_COUNT(18, 1);
                }
                if (mode == 0) consEnd = (int)conservativeEndOffset;
                else consBegin = (int)conservativeEndOffset;
              }
              consEnd = (int)(hi - lo) - consEnd;
            }
            if (lane < MAX_SETW) {
              sets[lane]            = support;
              sets[MAX_SETW + lane] = reject;
            }
            if (lane == 0) {
              CtgMeta m;
              m.lo = lo; m.hi = hi; m.consBegin = consBegin; m.consEnd = consEnd;
              W.ctgMeta[nCtg] = m;
              shU[8]          = isRepeatFound ? 1u : 0u;
              shU[9]          = overflow ? 1u : 0u;
            }
          }
          __syncthreads();
          PROF_MARK(8);
          if (shU[8]) isAssemblySuccess = false;
          if (shU[9]) status = MB200_ERR_ASM_OVERFLOW;
          nCtg++;
          __syncthreads();
          if (status != MB200_OK) break;
        }
        if (status != MB200_OK) break;
        if (isAssemblySuccess) break;

        // ---- pseudo reads for the next word length (:882-910): contigs longer than k + step
        if (tid == 0) {
          uint32_t po = normPackWords, mo = normMaskWords, nr = normalReadCount;
          for (uint32_t c = 0; c < nCtg; ++c) {
            const uint32_t len = W.ctgMeta[c].hi - W.ctgMeta[c].lo;
            if (len > k + opt.wordStepSize) {
              W.readLen[nr]     = len;
              W.readPackOff[nr] = po;
              W.readMaskOff[nr] = mo;
              W.tmpA[nr]        = c;  // source contig
              po += (len + 31) / 32 + 4;
              mo += (len + 63) / 64 + 3;
              nr++;
            }
          }
          shU[10] = nr;
        }
        __syncthreads();
        nReads = shU[10];
        for (uint32_t r = normalReadCount + warp; r < nReads; r += nt / 32) {
          const uint32_t c   = W.tmpA[r];
          const uint8_t* src = W.ctgSeq + (size_t)c * ctgBuf + W.ctgMeta[c].lo;
          const uint32_t len = W.readLen[r];
          uint64_t*      pk  = W.packed + W.readPackOff[r];
          uint64_t*      mk  = W.nmask + W.readMaskOff[r];
          const uint32_t nPk = (len + 31) / 32 + 4, nMk = (len + 63) / 64 + 3;
          packRead(src, len, pk, mk, nPk, nMk, false, lane);
        }
        __syncthreads();
        PROF_MARK(9);
        // NB: the pseudo reads must be packed from the contig buffers BEFORE the next buildContigs overwrites
        // them; the copies used for the output (below) are taken from the packed form.
      }
    }

    // ================= selectContigs (:722-842) + output (warp 0) =================
    PROF_MARK(10);
    if (warp == 0) {
      mb200_asm_locus Lo;
      memset(&Lo, 0, sizeof(Lo));
      Lo.firstContig    = li * opt.maxAssemblyCount;
      Lo.firstPseudo    = li * maxIter;
      Lo.setWords       = SW;
      Lo.lastWordLength = lastK;
      Lo.status         = status;
      if (status == MB200_OK) {
        uint32_t usedReads = 0, usedPseudo = 0;  // lane-wise words of the two std::set<unsigned>
        uint32_t aliveMask = (nCtg >= 32) ? 0xffffffffu : ((1u << nCtg) - 1);  // nCtg <= 2*maxAssemblyCount <= 32
        uint32_t nSel      = 0;
        uint32_t selected[1];
        (void)selected;
        // pseudo-read membership mask of my word: read index >= normalReadCount
        uint32_t pseudoMask;
        {
          const uint32_t lo = lane * 32;
          if (normalReadCount <= lo) pseudoMask = 0xffffffffu;
          else if (normalReadCount >= lo + 32) pseudoMask = 0;
          else pseudoMask = ~((1u << (normalReadCount - lo)) - 1);
        }
        // output bookkeeping gathered in shared memory by lane 0
        uint32_t outBytes = 0;
        while (aliveMask && nSel < opt.maxAssemblyCount) {
          const uint32_t usedNormal   = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(usedReads)) -
                                      __reduce_add_sync(0xffffffffu, (uint32_t)__popc(usedPseudo));
          const uint32_t unusedNormal = normalReadCount - usedNormal;
          if (unusedNormal < opt.minUnusedReads) break;
          int      sel = -1;
          uint32_t maxSupport = 0, maxLength = 0, removeMask = 0;
          for (uint32_t c = 0; c < nCtg; ++c) {
            if (!((aliveMask >> c) & 1)) continue;
            const uint32_t sup    = W.ctgSets[(size_t)c * 2 * MAX_SETW + lane];
            const uint32_t newSup = sup & ~usedReads;
            const uint32_t newNormal = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(newSup & ~pseudoMask));
            if (newNormal < opt.minSupportReads) {
              removeMask |= 1u << c;
              continue;
            }
            const uint32_t curSup = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(newSup));
            const uint32_t curLen = W.ctgMeta[c].hi - W.ctgMeta[c].lo;
            if (curSup > maxSupport || (curSup == maxSupport && curLen > maxLength)) {
              sel        = (int)c;
              maxSupport = curSup;
              maxLength  = curLen;
            }
          }
          if (maxSupport == 0) break;
          // emit the selected contig
          {
            const uint32_t c   = (uint32_t)sel;
            const CtgMeta  m   = W.ctgMeta[c];
            const uint32_t len = m.hi - m.lo;
            unsigned long long off = 0, soff = 0;
            if (lane == 0) {
              off  = atomicAdd(P.contigPoolUsed, (unsigned long long)len);
              soff = atomicAdd(P.setPoolUsed, (unsigned long long)(2 * SW));
            }
            off  = __shfl_sync(0xffffffffu, off, 0);
            soff = __shfl_sync(0xffffffffu, soff, 0);
            const uint8_t* src = W.ctgSeq + (size_t)c * ctgBuf + m.lo;
            if (off + len <= P.contigPoolCap)
              for (uint32_t i = lane; i < len; i += 32) P.contigPool[off + i] = src[i];
            if (soff + 2 * SW <= P.setPoolCap && lane < SW) {
              P.setPool[soff + lane]      = W.ctgSets[(size_t)c * 2 * MAX_SETW + lane];
              P.setPool[soff + SW + lane] = W.ctgSets[(size_t)c * 2 * MAX_SETW + MAX_SETW + lane];
            }
            if (lane == 0) {
              mb200_contig o;
              o.seqOff = off; o.seqLen = len; o.seedReadCount = 0;
              o.conservativeBegin = m.consBegin; o.conservativeEnd = m.consEnd;
              o.supportOff = soff; o.rejectOff = soff + SW;
              P.contigs[(size_t)Lo.firstContig + nSel] = o;
            }
            outBytes += len;
            const uint32_t sup = W.ctgSets[(size_t)c * 2 * MAX_SETW + lane];
            usedReads |= sup;
            usedPseudo |= sup & pseudoMask;
            removeMask |= 1u << c;
          }
          aliveMask &= ~removeMask;
          nSel++;
        }
        (void)outBytes;
        Lo.nContigs = nSel;
        // pseudo reads appended to the caller's read list (from the packed copies)
        Lo.nPseudo = nReads - normalReadCount;
        for (uint32_t r = normalReadCount; r < nReads; ++r) {
          const uint32_t     len = W.readLen[r];
          unsigned long long off = 0;
          if (lane == 0) off = atomicAdd(P.contigPoolUsed, (unsigned long long)len);
          off = __shfl_sync(0xffffffffu, off, 0);
          const uint64_t* pk = W.packed + W.readPackOff[r];
          if (off + len <= P.contigPoolCap)
            for (uint32_t i = lane; i < len; i += 32)
              P.contigPool[off + i] = "ACGT"[(pk[i >> 5] >> (62 - 2 * (i & 31))) & 3];
          if (lane == 0) {
            mb200_seqview v;
            v.off = off; v.len = len; v.reserved = 0;
            P.pseudo[(size_t)Lo.firstPseudo + (r - normalReadCount)] = v;
          }
        }
      } else if (lane == 0) {
        atomicCAS(P.errFlag, 0, status);
      }
      if (lane == 0) P.loci[li] = Lo;
    }
    __syncthreads();
    PROF_MARK(11);
  }
}

// ======================================================================================================
// On-chip assembler (round 2).  Same algorithm and results as assembleKernel above, different residence of the data:
//   * one CTA of CHIP_THREADS threads per locus; every per-word table the hot phases touch lives in SHARED memory under
//     DENSE word ids (16 bit): slot table, read-support bitsets, successor / predecessor ids, count, flags, in-degrees;
//   * words are not stored at all: a word is the (read, offset) of one of its windows (4 bytes) and its key is cut out
//     of the packed reads on demand (the packed normal reads sit in shared memory too when they fit);
//   * count and first read are derived from the bitset (count = normal reads + minCoverage * pseudo reads that contain
//     the word, IterativeAssembler.cpp:524-541), so insertion needs one shared atomicOr per window and nothing else;
//   * global memory keeps only what is streamed (window -> slot map, contig buffers) or rare (the arrays of the exact
//     unordered_map replay, now sized by the word capacity instead of the window count).
// A locus that does not fit (more than Wc distinct words at some word length, more than 32*SWc reads incl. pseudo
// reads, a read longer than 65535) is appended to a retry list and assembled by assembleKernel afterwards.
constexpr int      CHIP_THREADS = 128;
constexpr uint32_t NONE16       = 0xffffu;

struct ChipLayout {
  size_t   packed, nmask, winSlot, hashv, hash2, seqA, arr, pos, bktHead, bktMin, bktCnt, chainNext, gs, orderA, tmpA, idx, low,
           wstack, frame, ctgSeq, ctgSets, ctgMeta, total;
  uint32_t maxReads, nWinMax, packWords, maskWords, bktMax;
  uint32_t Wc, SWc, SL, SPK;  // word capacity, set words per word, slots (power of two, >= Wc + CHIP_THREADS + 1), shared pack words
  uint32_t smemBytes;
};

struct ChipParams {
  AsmParams  A;  // opt, inputs, outputs, counters (work / L unused)
  ChipLayout C;
  uint8_t*   work;
  uint32_t*  retryList;
  uint32_t*  retryCount;
};

struct Chip {  // shared-memory view
  uint32_t *slot, *ref, *bits;
  uint16_t *succ, *pred, *count, *queue;
  uint8_t * flags, *inC, *indeg;
  uint32_t *len, *packOff, *maskOff, *winStart;
  uint64_t*       spk;
  const uint64_t* gPacked;
  uint32_t        nShared;  // reads [0, nShared) have their packed words in spk
  uint32_t        SWc;
};

__device__ __forceinline__ const uint64_t* packOf(const Chip& S, uint32_t r) { return (r < S.nShared ? S.spk : S.gPacked) + S.packOff[r]; }
__device__ __forceinline__ Key keyOfRef(const Chip& S, uint32_t ref, uint32_t k) { return extractKey(packOf(S, ref >> 16), ref & 0xffffu, k); }

__device__ uint32_t findWordChip(const Chip& S, const Key& key, uint32_t k, uint32_t slotMask)
{
  const uint32_t h32 = hashKey(key), tag = h32 >> 27;
  uint32_t       h   = h32 & slotMask;
  for (;;) {
    const uint32_t v = S.slot[h];
    if (v == 0) return NONE16;
    if ((v >> 27) == tag) {
      const uint32_t id = (v & 0x7ffffffu) - 1;
      if (keyEq(keyOfRef(S, S.ref[id], k), key)) return id;
    }
    h = (h + 1) & slotMask;
  }
}

// first read of a word / its count, from the bitset
__device__ __forceinline__ uint32_t firstReadOf(const Chip& S, uint32_t id, uint32_t SW)
{
  for (uint32_t w = 0; w < SW; ++w) {
    const uint32_t b = S.bits[id * S.SWc + w];
    if (b) return w * 32 + (uint32_t)__ffs((int)b) - 1;
  }
  return NONE32;
}

__global__ void __launch_bounds__(CHIP_THREADS) assembleChipKernel(const ChipParams P)
{
  extern __shared__ __align__(16) uint8_t smemRaw[];
  __shared__ uint32_t sh[CHIP_THREADS];
  __shared__ uint32_t shU[20];

  const uint32_t tid  = threadIdx.x;
  const uint32_t nt   = blockDim.x;
  const uint32_t lane = tid & 31;
  const uint32_t warp = tid >> 5;
  const uint32_t Wc = P.C.Wc, SWc = P.C.SWc, SL = P.C.SL, RC = SWc * 32;

  Chip S;
  {
    uint8_t* q = smemRaw;
    S.spk      = reinterpret_cast<uint64_t*>(q); q += 8 * (size_t)P.C.SPK;
    S.slot     = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)SL;
    S.ref      = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)Wc;
    S.bits     = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)Wc * SWc;
    S.len      = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)RC;
    S.packOff  = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)RC;
    S.maskOff  = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)RC;
    S.winStart = reinterpret_cast<uint32_t*>(q); q += 4 * (size_t)(RC + 4);
    S.succ     = reinterpret_cast<uint16_t*>(q); q += 8 * (size_t)Wc;
    S.pred     = reinterpret_cast<uint16_t*>(q); q += 8 * (size_t)Wc;
    S.count    = reinterpret_cast<uint16_t*>(q); q += 2 * (size_t)Wc;
    S.queue    = reinterpret_cast<uint16_t*>(q); q += 2 * (size_t)Wc;
    S.flags    = q; q += Wc;
    S.inC      = q; q += Wc;
    S.indeg    = q; q += Wc;
    S.SWc      = SWc;
  }
  uint8_t*  base    = P.work + (size_t)blockIdx.x * P.C.total;
  uint64_t* gPacked = reinterpret_cast<uint64_t*>(base + P.C.packed);
  uint64_t* gMask   = reinterpret_cast<uint64_t*>(base + P.C.nmask);
  uint16_t* winSlot = reinterpret_cast<uint16_t*>(base + P.C.winSlot);
  S.gPacked         = gPacked;
  Work W;  // the arrays of the unordered_map replay (global, sized by Wc)
  W.hashv     = reinterpret_cast<uint64_t*>(base + P.C.hashv);
  uint64_t* hash2 = reinterpret_cast<uint64_t*>(base + P.C.hash2);
  W.seqA      = reinterpret_cast<uint32_t*>(base + P.C.seqA);
  W.arr       = reinterpret_cast<uint32_t*>(base + P.C.arr);
  W.pos       = reinterpret_cast<uint32_t*>(base + P.C.pos);
  W.bktHead   = reinterpret_cast<uint32_t*>(base + P.C.bktHead);
  W.bktMin    = reinterpret_cast<uint32_t*>(base + P.C.bktMin);
  W.bktCnt    = reinterpret_cast<uint32_t*>(base + P.C.bktCnt);
  W.chainNext = reinterpret_cast<uint32_t*>(base + P.C.chainNext);
  W.gs        = reinterpret_cast<uint32_t*>(base + P.C.gs);
  W.orderA    = reinterpret_cast<uint32_t*>(base + P.C.orderA);
  W.tmpA      = reinterpret_cast<uint32_t*>(base + P.C.tmpA);
  W.idx       = reinterpret_cast<uint32_t*>(base + P.C.idx);
  W.low       = reinterpret_cast<uint32_t*>(base + P.C.low);
  W.wstack    = reinterpret_cast<uint32_t*>(base + P.C.wstack);
  W.frame     = reinterpret_cast<uint32_t*>(base + P.C.frame);
  W.ctgSeq    = base + P.C.ctgSeq;
  W.ctgSets   = reinterpret_cast<uint32_t*>(base + P.C.ctgSets);
  W.ctgMeta   = reinterpret_cast<CtgMeta*>(base + P.C.ctgMeta);

  const mb200_asm_opts opt     = P.A.opt;
  const uint32_t       maxIter = 2 * opt.maxAssemblyCount;
  const uint32_t       ctgBuf  = 2 * P.A.maxContigLen;
  const uint32_t       slotMask = SL - 1;

  for (;;) {
    if (tid == 0) shU[0] = atomicAdd(P.A.lociCounter, 1u);
    __syncthreads();
    const uint32_t li = shU[0];
    __syncthreads();
    if (li >= P.A.nLoci) break;

    PROF_DECL();
    PROF_COUNT(19, 1);
    const uint64_t r0              = P.A.readOff[li];
    const uint32_t normalReadCount = (uint32_t)(P.A.readOff[li + 1] - r0);
    const uint32_t SW              = (normalReadCount + maxIter + 31) / 32;
    int            status          = MB200_OK;
    bool           retry           = (SW > SWc);  // does not fit: legacy kernel

    // ---------------- pack the normal reads
    if (!retry) {
      if (tid == 0) {
        uint32_t po = 0, mo = 0, tooLong = 0;
        for (uint32_t r = 0; r < normalReadCount; ++r) {
          const uint32_t len = P.A.views[r0 + r].len;
          S.len[r]     = len;
          S.packOff[r] = po;
          S.maskOff[r] = mo;
          if (len > 0xffffu) tooLong = 1;
          po += (len + 31) / 32 + 4;
          mo += (len + 63) / 64 + 3;
        }
        shU[1] = po;
        shU[2] = mo;
        shU[11] = tooLong;
      }
      __syncthreads();
      if (shU[11]) retry = true;
    }
    const uint32_t normPackWords = shU[1], normMaskWords = shU[2];
    if (!retry) {
      bool badBase = false;
      for (uint32_t r = warp; r < normalReadCount; r += nt / 32) {
        const uint8_t* src = P.A.pool + P.A.views[r0 + r].off;
        const uint32_t len = S.len[r];
        const uint32_t nPk = (len + 31) / 32 + 4, nMk = (len + 63) / 64 + 3;
        if (packRead(src, len, gPacked + S.packOff[r], gMask + S.maskOff[r], nPk, nMk, true, lane)) badBase = true;
      }
      if (__syncthreads_or(badBase ? 1 : 0)) status = MB200_ERR_BAD_BASE;
      // packed normal reads to shared memory (they are read once per window and per key comparison)
      S.nShared = 0;
      if (normPackWords <= P.C.SPK) {
        for (uint32_t i = tid; i < normPackWords; i += nt) S.spk[i] = gPacked[i];
        S.nShared = normalReadCount;
      }
      __syncthreads();
    }

    uint32_t nReads = normalReadCount;
    uint32_t nCtg   = 0;
    uint32_t lastK  = 0;

    if (!retry && status == MB200_OK) {
      PROF_MARK(0);
      for (uint32_t k = opt.minWordLength; k <= opt.maxWordLength; k += opt.wordStepSize) {
        lastK = k;
        PROF_COUNT(16, 1);
        // ---- windows
        if (tid == 0) {
          uint32_t acc = 0;
          for (uint32_t r = 0; r < nReads; ++r) {
            S.winStart[r] = acc;
            const uint32_t len = S.len[r];
            if (len >= k) acc += len - k + 1;
          }
          S.winStart[nReads] = acc;
          shU[3] = acc;
          shU[4] = 0;   // claimed slots
          shU[12] = 0;  // overflow
          shU[13] = 0;  // dense ids handed out
        }
        for (uint32_t i = tid; i < SL; i += nt) S.slot[i] = 0;
        __syncthreads();
        PROF_MARK(1);
        PROF_COUNT(21, shU[3]);

        // ---- pass A: every N-free window claims or finds the slot of its word
        for (uint32_t r = warp; r < nReads; r += nt / 32) {
          const uint32_t len = S.len[r];
          if (len < k) continue;
          const uint64_t* pk = packOf(S, r);
          const uint64_t* mk = gMask + S.maskOff[r];
          const uint32_t  ws = S.winStart[r];
          for (uint32_t j = lane; j + k <= len; j += 32) {
            bool hasN = getBits(mk, j, k < 64 ? k : 64) != 0;
            if (k > 64) hasN = hasN || (getBits(mk, j + 64, k - 64) != 0);
            uint32_t found = NONE16;
            if (!hasN) {
              const Key      key = extractKey(pk, j, k);
              const uint32_t h32 = hashKey(key), tag = h32 >> 27;
              const uint32_t val = (tag << 27) | (((r << 16) | j) + 1);
              uint32_t       h   = h32 & slotMask;
              for (;;) {
                uint32_t cur = S.slot[h];
                if (cur == 0) {
                  if (*reinterpret_cast<volatile uint32_t*>(&shU[12])) break;  // capacity exceeded: stop claiming
                  cur = atomicCAS(&S.slot[h], 0u, val);
                  if (cur == 0) {
                    if (atomicAdd(&shU[4], 1u) >= Wc) shU[12] = 1;
                    found = h;
                    break;
                  }
                }
                if ((cur >> 27) == tag && keyEq(keyOfRef(S, (cur & 0x7ffffffu) - 1, k), key)) {
                  found = h;
                  break;
                }
                h = (h + 1) & slotMask;
              }
            }
            winSlot[ws + j] = (uint16_t)found;
          }
        }
        __syncthreads();
        if (shU[12]) {  // more distinct words than the shared tables hold
          retry = true;
          break;
        }
        PROF_MARK(2);
        // ---- pass B: dense ids
        for (uint32_t h = tid; h < SL; h += nt) {
          const uint32_t v = S.slot[h];
          if (v == 0) continue;
          const uint32_t id = atomicAdd(&shU[13], 1u);
          S.ref[id]   = (v & 0x7ffffffu) - 1;
          S.slot[h]   = (v & 0xf8000000u) | (id + 1);
          S.flags[id] = 0;
          S.inC[id]   = 0;
          S.indeg[id] = 0;
          for (uint32_t w = 0; w < SW; ++w) S.bits[id * SWc + w] = 0;
        }
        __syncthreads();
        const uint32_t nWords = shU[13];
        // ---- pass C: read-support bits (std::set<std::string> readWords: a word repeated inside a read counts once)
        for (uint32_t r = warp; r < nReads; r += nt / 32) {
          const uint32_t len = S.len[r];
          if (len < k) continue;
          const uint32_t ws  = S.winStart[r];
          const uint32_t bit = 1u << (r & 31);
          for (uint32_t j = lane; j + k <= len; j += 32) {
            const uint32_t hs = winSlot[ws + j];
            if (hs == NONE16) continue;
            const uint32_t id = (S.slot[hs] & 0x7ffffffu) - 1;
            atomicOr(&S.bits[id * SWc + (r >> 5)], bit);
          }
        }
        __syncthreads();
        // ---- counts: 1 per normal read, minCoverage per pseudo read (:541)
        for (uint32_t id = tid; id < nWords; id += nt) {
          uint32_t c = 0;
          for (uint32_t w = 0; w < SW; ++w) {
            const uint32_t b  = S.bits[id * SWc + w];
            const uint32_t lo = w * 32;
            uint32_t       pm;  // pseudo-read mask of this set word
            if (normalReadCount <= lo) pm = 0xffffffffu;
            else if (normalReadCount >= lo + 32) pm = 0;
            else pm = ~((1u << (normalReadCount - lo)) - 1);
            c += (uint32_t)__popc(b & ~pm) + opt.minCoverage * (uint32_t)__popc(b & pm);
          }
          S.count[id] = (uint16_t)(c > 0xffffu ? 0xffffu : c);
        }
        __syncthreads();
        PROF_MARK(3);
        PROF_COUNT(20, nWords);

        // ---- successor / predecessor ids, homopolymer repeats, in-degrees
        for (uint32_t i = tid; i < nWords * 4; i += nt) {
          const uint32_t id  = i >> 2;
          const uint32_t sym = i & 3;
          const Key      key = keyOfRef(S, S.ref[id], k);
          const uint32_t s   = findWordChip(S, succKey(key, sym, k), k, slotMask);
          S.succ[id * 4 + sym] = (uint16_t)s;
          if (s == id) {
            atomicOr(reinterpret_cast<unsigned int*>(S.flags + (id & ~3u)), (unsigned)FLAG_REPEAT << (8 * (id & 3)));
          } else if (s != NONE16) {
            atomicAdd(reinterpret_cast<unsigned int*>(S.indeg + (s & ~3u)), 1u << (8 * (s & 3)));
          }
        }
        __syncthreads();
        PROF_MARK(4);

        // ---- Kahn trimming: any cycle besides homopolymer self loops?
        {
          if (tid == 0) shU[5] = 0;
          __syncthreads();
          for (uint32_t id = tid; id < nWords; id += nt)
            if (S.indeg[id] == 0) S.queue[atomicAdd(&shU[5], 1u)] = (uint16_t)id;
          __syncthreads();
          uint32_t head = 0;
          for (;;) {
            const uint32_t tail = shU[5];
            if (head == tail) break;
            __syncthreads();
            for (uint32_t i = head + tid; i < tail; i += nt) {
              const uint32_t id = S.queue[i];
              const uint2    sc = *reinterpret_cast<const uint2*>(S.succ + id * 4);
              const uint32_t ss[4] = {sc.x & 0xffffu, sc.x >> 16, sc.y & 0xffffu, sc.y >> 16};
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const uint32_t nx = ss[s];
                if (nx == NONE16 || nx == id) continue;
                const uint32_t shf = 8 * (nx & 3);
                const uint32_t old = atomicSub(reinterpret_cast<unsigned int*>(S.indeg + (nx & ~3u)), 1u << shf);
                if (((old >> shf) & 0xffu) == 1u) S.queue[atomicAdd(&shU[5], 1u)] = (uint16_t)nx;
              }
            }
            head = tail;
            __syncthreads();
          }
        }
        const bool hasCycle = (shU[5] != nWords);
        __syncthreads();
        PROF_MARK(5);
        PROF_COUNT(22, hasCycle ? 1 : 0);

        if (hasCycle) {
          // ================= exact replay of getRepeatKmers (hash-order DFS roots), see assembleKernel =================
          uint32_t* bucketStart = W.tmpA;
          uint32_t* members     = W.orderA;
          uint32_t* firstRead   = W.idx;  // until the DFS needs idx
          for (uint32_t r = tid; r <= nReads; r += nt) W.gs[r] = 0;
          __syncthreads();
          for (uint32_t id = tid; id < nWords; id += nt) {
            const uint32_t fr = firstReadOf(S, id, SW);
            firstRead[id]     = fr;
            atomicAdd(&W.gs[fr], 1u);
          }
          __syncthreads();
          if (tid == 0) {
            uint32_t acc = 0;
            for (uint32_t r = 0; r < nReads; ++r) {
              bucketStart[r] = acc;
              acc += W.gs[r];
              W.gs[r] = 0;
            }
            bucketStart[nReads] = acc;
          }
          __syncthreads();
          for (uint32_t id = tid; id < nWords; id += nt) {
            const uint32_t r = firstRead[id];
            members[bucketStart[r] + atomicAdd(&W.gs[r], 1u)] = id;
          }
          __syncthreads();
          for (uint32_t id = tid; id < nWords; id += nt) {
            const uint32_t r   = firstRead[id];
            const Key      key = keyOfRef(S, S.ref[id], k);
            uint32_t       rank = 0;
            for (uint32_t m = bucketStart[r]; m < bucketStart[r + 1]; ++m)
              if (keyLess(keyOfRef(S, S.ref[members[m]], k), key)) rank++;
            const uint32_t p = bucketStart[r] + rank;
            W.seqA[p]        = id;
            W.hashv[p]       = stlHashWord(key, k);
          }
          __syncthreads();
          uint32_t* order1 = W.low;  // order1 and order2 must survive until chainNext is built
          unorderedOrder(W, W.hashv, nWords, order1, sh);
          for (uint32_t i = tid; i < nWords; i += nt) hash2[i] = W.hashv[order1[i]];
          __syncthreads();
          uint32_t* order2 = W.wstack;
          unorderedOrder(W, hash2, nWords, order2, sh);
          for (uint32_t i = tid; i < nWords; i += nt) W.frame[i] = W.seqA[order1[order2[i]]];  // root sequence
          __syncthreads();
          for (uint32_t i = tid; i < nWords; i += nt) W.chainNext[i] = W.frame[i];
          __syncthreads();
          PROF_MARK(12);
          // the DFS state lives in shared memory: idx on the (finished) Kahn queue, low on inC+indeg (contiguous, both
          // dead here), the two stacks on the predecessor table, which is only filled afterwards
          uint16_t* dIdx   = S.queue;
          uint16_t* dLow   = reinterpret_cast<uint16_t*>(S.inC);
          uint16_t* dStack = S.pred;
          uint16_t* dFrame = S.pred + Wc;      // (word << 3) | next symbol needs 14 bits
          for (uint32_t id = tid; id < nWords; id += nt) {
            dIdx[id] = 0;
            dLow[id] = 0;
          }
          __syncthreads();
          if (tid == 0) {
            const uint32_t* roots  = W.chainNext;
            uint32_t        index  = 1;
            uint32_t        nStack = 0;
            for (uint32_t ri = 0; ri < nWords; ++ri) {
              const uint32_t root = roots[ri];
              if (dIdx[root] != 0) continue;
              uint32_t depth = 1;
              dFrame[0]      = (uint16_t)(root << 3);
              dIdx[root] = dLow[root] = (uint16_t)index++;
              dStack[nStack++]        = (uint16_t)root;
              S.flags[root] |= FLAG_ONSTACK;
              while (depth) {
                const uint32_t fr  = dFrame[depth - 1];
                const uint32_t w   = fr >> 3;
                const uint32_t sym = fr & 7;
                if (sym < 4) {
                  dFrame[depth - 1] = (uint16_t)(fr + 1);
                  const uint32_t nw = S.succ[w * 4 + sym];
                  if (nw == w) continue;
                  if (nw == NONE16) continue;
                  if (dIdx[nw] == 0) {
                    dFrame[depth++] = (uint16_t)(nw << 3);
                    dIdx[nw] = dLow[nw] = (uint16_t)index++;
                    dStack[nStack++]    = (uint16_t)nw;
                    S.flags[nw] |= FLAG_ONSTACK;
                  } else if (S.flags[nw] & FLAG_ONSTACK) {
                    if (dIdx[nw] < dLow[w]) dLow[w] = dIdx[nw];
                  }
                } else {
                  if (dLow[w] == dIdx[w]) {
                    const uint32_t lastWord = dStack[nStack - 1];
                    if (lastWord == w) {
                      nStack--;
                      S.flags[w] &= ~FLAG_ONSTACK;
                    } else {
                      const bool isSmallCircle = ((uint32_t)dIdx[lastWord] - (uint32_t)dIdx[w]) <= 50;
                      for (;;) {
                        const uint32_t rw = dStack[--nStack];
                        S.flags[rw] &= ~FLAG_ONSTACK;
                        if (isSmallCircle) S.flags[rw] |= FLAG_REPEAT;
                        if (rw == w) break;
                      }
                    }
                  }
                  depth--;
                  if (depth) {
                    const uint32_t parent = dFrame[depth - 1] >> 3;
                    if (dLow[w] < dLow[parent]) dLow[parent] = dLow[w];
                  }
                }
              }
            }
          }
          __syncthreads();
          for (uint32_t id = tid; id < nWords; id += nt) S.inC[id] = 0;  // was the low-link array
          __syncthreads();
          PROF_MARK(13);
        }
        PROF_MARK(6);

        // ---- predecessor ids (their table doubled as the DFS stacks), unusedWords = {count >= minCoverage}
        for (uint32_t i = tid; i < nWords * 4; i += nt) {
          const uint32_t id  = i >> 2;
          const uint32_t sym = i & 3;
          S.pred[id * 4 + sym] = (uint16_t)findWordChip(S, predKey(keyOfRef(S, S.ref[id], k), sym, k), k, slotMask);
        }
        __syncthreads();
        for (uint32_t id = tid; id < nWords; id += nt) {
          const uint2 sc = *reinterpret_cast<const uint2*>(S.succ + id * 4);
          const uint2 pc = *reinterpret_cast<const uint2*>(S.pred + id * 4);
          const int   nS = ((sc.x & 0xffffu) != NONE16) + ((sc.x >> 16) != NONE16) + ((sc.y & 0xffffu) != NONE16) + ((sc.y >> 16) != NONE16);
          const int   nP = ((pc.x & 0xffffu) != NONE16) + ((pc.x >> 16) != NONE16) + ((pc.y & 0xffffu) != NONE16) + ((pc.y >> 16) != NONE16);
          uint32_t    f  = S.flags[id];
          if (S.count[id] >= opt.minCoverage) f |= FLAG_UNUSED;
          if (nP <= 1) f |= FLAG_ONE_PRED;
          if (nS <= 1) f |= FLAG_ONE_SUCC;
          S.flags[id] = (uint8_t)f;
        }
        __syncthreads();
        PROF_MARK(14);

        // ---- contig loop (:685-716)
        bool isAssemblySuccess = true;
        nCtg                   = 0;
        while (nCtg < maxIter) {
          // seed: max count, ties -> lexicographically smallest word.  Stage 1: the maximum count.
          uint32_t myMax = 0;
          bool     any   = false;
          for (uint32_t id = tid; id < nWords; id += nt) {
            if (!(S.flags[id] & FLAG_UNUSED)) continue;
            const uint32_t c = S.count[id];
            if (!any || c > myMax) myMax = c;
            any = true;
          }
          uint32_t enc = any ? (myMax + 1) : 0;  // 0 = nothing unused
          enc          = __reduce_max_sync(0xffffffffu, enc);
          if (lane == 0) sh[warp] = enc;
          __syncthreads();
          uint32_t encAll = 0;
          for (uint32_t w = 0; w < nt / 32; ++w) encAll = max(encAll, sh[w]);
          __syncthreads();
          if (encAll == 0) {
            PROF_MARK(7);
            break;  // unusedWords.empty()
          }
          const uint32_t cMax = encAll - 1;
          // stage 2: the smallest word among those with that count
          Key      bestKey = {~0ull, ~0ull, ~0ull};
          uint32_t bestId  = NONE32;
          for (uint32_t id = tid; id < nWords; id += nt) {
            if (!(S.flags[id] & FLAG_UNUSED) || S.count[id] != cMax) continue;
            const Key key = keyOfRef(S, S.ref[id], k);
            if (bestId == NONE32 || keyLess(key, bestKey)) {
              bestKey = key;
              bestId  = id;
            }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const uint32_t oid = __shfl_xor_sync(0xffffffffu, bestId, off);
            Key            ok;
            ok.k0 = __shfl_xor_sync(0xffffffffu, bestKey.k0, off);
            ok.k1 = __shfl_xor_sync(0xffffffffu, bestKey.k1, off);
            ok.k2 = __shfl_xor_sync(0xffffffffu, bestKey.k2, off);
            if (oid != NONE32 && (bestId == NONE32 || keyLess(ok, bestKey))) {
              bestKey = ok;
              bestId  = oid;
            }
          }
          if (lane == 0) sh[warp] = bestId;
          __syncthreads();
          if (tid == 0) {
            uint32_t bid = NONE32;
            Key      bk  = {0, 0, 0};
            for (uint32_t w = 0; w < nt / 32; ++w) {
              const uint32_t id = sh[w];
              if (id == NONE32) continue;
              const Key kk = keyOfRef(S, S.ref[id], k);
              if (bid == NONE32 || keyLess(kk, bk)) {
                bid = id;
                bk  = kk;
              }
            }
            shU[7] = bid;
          }
          __syncthreads();
          const uint32_t seed = shU[7];
          __syncthreads();
          PROF_MARK(7);
          PROF_COUNT(17, 1);

          // ---------------- walk(seed)  (warp 0; IterativeAssembler.cpp:149-501)
          if (warp == 0) {
            uint8_t*       seq     = W.ctgSeq + (size_t)nCtg * ctgBuf;
            uint32_t*      sets    = W.ctgSets + (size_t)nCtg * 2 * MAX_SETW;
            const uint32_t serial  = nCtg + 1;
            const Key      seedKey = keyOfRef(S, S.ref[seed], k);
            uint32_t       lo = P.A.maxContigLen - k / 2, hi;
            hi = lo + k;
            for (uint32_t i = lane; i < k; i += 32) seq[lo + i] = "ACGT"[baseAt(seedKey, i, k)];
            uint32_t support = (lane < SW) ? S.bits[seed * SWc + lane] : 0u;
            uint32_t reject  = 0;
            const bool seedIsRepeat = (S.flags[seed] & FLAG_REPEAT) != 0;
            __syncwarp();
            if (lane == 0) S.flags[seed] &= ~FLAG_UNUSED;
            int  consBegin = 0, consEnd = 0;
            bool isRepeatFound = false;
            bool overflow      = false;
            if (seedIsRepeat) {
              consBegin     = 0;
              consEnd       = (int)k;
              isRepeatFound = true;
            } else {
              if (lane == 0) S.inC[seed] = (uint8_t)serial;
              {  // rejecting reads from the seed's sibling words (same k-1 prefix)  :185-210
                const uint32_t lastSym = (uint32_t)seedKey.k0 & 3u;
                Key            trunkK  = seedKey;
                trunkK.k0 &= ~3ull;
                for (uint32_t s = 0; s < 4; ++s) {
                  if (s == lastSym) continue;
                  Key kk = trunkK;
                  kk.k0 |= s;
                  const uint32_t id = findWordChip(S, kk, k, slotMask);
                  if (id == NONE16) continue;
                  if (lane < SW) reject |= S.bits[id * SWc + lane];
                }
              }
              __syncwarp();
              uint32_t curWord = seed;
              for (int mode = 0; mode < 2; ++mode) {
                const bool      isEnd = (mode == 0);
                const uint16_t* fwd   = isEnd ? S.succ : S.pred;
                const uint16_t* bwd   = isEnd ? S.pred : S.succ;
                uint32_t        conservativeEndOffset = 0;
                curWord = seed;
                const uint32_t oneBack = isEnd ? (uint32_t)FLAG_ONE_PRED : (uint32_t)FLAG_ONE_SUCC;
                for (;;) {
                  const uint2    c2 = *reinterpret_cast<const uint2*>(fwd + curWord * 4);
                  const uint32_t cand[4] = {c2.x & 0xffffu, c2.x >> 16, c2.y & 0xffffu, c2.y >> 16};
                  const uint32_t vmask = (cand[0] != NONE16 ? 1u : 0u) | (cand[1] != NONE16 ? 2u : 0u) | (cand[2] != NONE16 ? 4u : 0u) |
                                         (cand[3] != NONE16 ? 8u : 0u);
                  if (vmask == 0) break;  // nothing to extend with: maxBaseCount stays 0 < minCoverage
                  if ((vmask & (vmask - 1)) == 0) {
                    // ---- one candidate (most steps): the general step below with everything about "the other candidates" gone
                    const uint32_t sym = (uint32_t)__ffs((int)vmask) - 1;
                    const uint32_t id  = ((sym & 2u) ? c2.y : c2.x) >> (16 * (sym & 1u)) & 0xffffu;
                    const uint32_t cw1 = (lane < SW) ? S.bits[id * SWc + lane] : 0u;
                    const uint32_t cn1 = S.count[id];
                    const uint32_t in1 = S.inC[id];
                    const uint32_t fl1 = S.flags[id];
                    const uint32_t contig1 = support & cw1;
                    if (__reduce_add_sync(0xffffffffu, (uint32_t)__popc(contig1)) == 0) break;
                    if (cn1 < opt.minCoverage) break;
                    if (in1 == serial) {
                      isRepeatFound = true;
                      break;
                    }
                    uint32_t upd = 0;
                    if (!(fl1 & oneBack)) {  // the chosen word has other neighbours on the side we come from (:375-427)
                      const uint2    a2 = *reinterpret_cast<const uint2*>(bwd + id * 4);
                      const uint32_t alt[4] = {a2.x & 0xffffu, a2.x >> 16, a2.y & 0xffffu, a2.y >> 16};
#pragma unroll
                      for (int t = 0; t < 4; ++t) {
                        const uint32_t o = alt[t];
                        if (o == NONE16 || o == id || o == curWord) continue;
                        const uint32_t back = (lane < SW) ? S.bits[o * SWc + lane] : 0u;
                        upd |= back & ~(contig1 & back);
                      }
                    }
                    if (hi - lo + 1 > P.A.maxContigLen || lo == 0 || hi + 1 >= ctgBuf) {
                      overflow = true;
                      break;
                    }
                    if (lane == 0) seq[isEnd ? hi : lo - 1] = (uint8_t)(0x54474341u >> (8 * sym));  // "ACGT"[sym]
                    if (isEnd) hi++; else lo--;
                    if ((conservativeEndOffset != 0) || (cn1 < opt.minConservativeCoverage)) conservativeEndOffset += 1;
                    reject |= upd;
                    support |= cw1 & ~reject;
                    support &= ~upd;
                    __syncwarp();
                    if (lane == 0) {
                      S.flags[id] = (uint8_t)(fl1 & ~(uint32_t)FLAG_UNUSED);
                      S.inC[id]   = (uint8_t)serial;
                    }
                    __syncwarp();
                    curWord = id;
                    PROF_COUNT(18, 1);
                    continue;
                  }
                  // everything the step can need of the (usually one) candidates, loaded before any of it is used
                  uint32_t cw[4], cnt[4], cc[4], inc[4], flg[4];
                  uint2    bw[4];
#pragma unroll
                  for (int s = 0; s < 4; ++s) {
                    const uint32_t id = cand[s];
                    const bool     v  = (id != NONE16);
                    cw[s]  = (v && lane < SW) ? S.bits[id * SWc + lane] : 0u;
                    cnt[s] = v ? (uint32_t)S.count[id] : 0u;
                    inc[s] = v ? (uint32_t)S.inC[id] : 0u;
                    flg[s] = v ? (uint32_t)S.flags[id] : 0u;
                    bw[s]  = v ? *reinterpret_cast<const uint2*>(bwd + id * 4) : make_uint2(0xffffffffu, 0xffffffffu);
                  }
#pragma unroll
                  for (int s = 0; s < 4; ++s) cc[s] = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(support & cw[s]));
                  uint32_t maxBaseCount = 0, maxCnt = 0, maxBase = 0, maxWord = NONE16, maxInc = 0, maxFlg = 0;
                  uint32_t maxW = 0, maxCW = 0, sup2Remove = 0, rej2Add = 0;
                  uint2    maxBw = make_uint2(0xffffffffu, 0xffffffffu);
#pragma unroll
                  for (int s = 0; s < 4; ++s) {
                    if (cc[s] == 0) continue;  // not a word, or no read of the contig supports it
                    const uint32_t contigWord = support & cw[s];
                    const uint32_t shared     = maxCW & cw[s];
                    if (cc[s] > maxCnt) {
                      sup2Remove |= maxCW & ~shared;
                      rej2Add |= maxW & ~shared;
                      maxW         = cw[s];
                      maxCnt       = cc[s];
                      maxCW        = contigWord;
                      maxBaseCount = cnt[s];
                      maxBase      = (uint32_t)s;
                      maxWord      = cand[s];
                      maxInc       = inc[s];
                      maxFlg       = flg[s];
                      maxBw        = bw[s];
                    } else {
                      sup2Remove |= contigWord & ~shared;
                      rej2Add |= cw[s] & ~shared;
                    }
                  }
                  if (maxBaseCount < opt.minCoverage) break;
                  if (maxInc == serial) {
                    isRepeatFound = true;
                    break;
                  }
                  {  // walk backwards one step at the branching point (:375-427): the other neighbours of the chosen word
                    const uint32_t alt[4] = {maxBw.x & 0xffffu, maxBw.x >> 16, maxBw.y & 0xffffu, maxBw.y >> 16};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                      const uint32_t id = alt[s];
                      if (id == NONE16 || id == maxWord || id == curWord) continue;
                      const uint32_t back     = (lane < SW) ? S.bits[id * SWc + lane] : 0u;
                      const uint32_t toUpdate = back & ~(maxCW & back);
                      rej2Add |= toUpdate;
                      sup2Remove |= toUpdate;
                    }
                  }
                  if (hi - lo + 1 > P.A.maxContigLen || lo == 0 || hi + 1 >= ctgBuf) {
                    overflow = true;
                    break;
                  }
                  if (lane == 0) seq[isEnd ? hi : lo - 1] = (uint8_t)(0x54474341u >> (8 * maxBase));  // "ACGT"[maxBase]
                  if (isEnd) hi++; else lo--;
                  if ((conservativeEndOffset != 0) || (maxBaseCount < opt.minConservativeCoverage)) conservativeEndOffset += 1;
                  reject |= rej2Add;
                  support |= maxW & ~reject;
                  support &= ~sup2Remove;
                  __syncwarp();
                  if (lane == 0) {
                    S.flags[maxWord] = (uint8_t)(maxFlg & ~(uint32_t)FLAG_UNUSED);
                    S.inC[maxWord]   = (uint8_t)serial;
                  }
                  __syncwarp();
                  curWord = maxWord;
                  PROF_COUNT(18, 1);
                }
                if (mode == 0) consEnd = (int)conservativeEndOffset;
                else consBegin = (int)conservativeEndOffset;
              }
              consEnd = (int)(hi - lo) - consEnd;
            }
            if (lane < MAX_SETW) {
              sets[lane]            = support;
              sets[MAX_SETW + lane] = reject;
            }
            if (lane == 0) {
              CtgMeta m;
              m.lo = lo; m.hi = hi; m.consBegin = consBegin; m.consEnd = consEnd;
              W.ctgMeta[nCtg] = m;
              shU[8]          = isRepeatFound ? 1u : 0u;
              shU[9]          = overflow ? 1u : 0u;
            }
          }
          __syncthreads();
          PROF_MARK(8);
          if (shU[8]) isAssemblySuccess = false;
          if (shU[9]) status = MB200_ERR_ASM_OVERFLOW;
          nCtg++;
          __syncthreads();
          if (status != MB200_OK) break;
        }
        if (status != MB200_OK) break;
        if (isAssemblySuccess) break;

        // ---- pseudo reads for the next word length (:882-910)
        if (tid == 0) {
          uint32_t po = normPackWords, mo = normMaskWords, nr = normalReadCount;
          for (uint32_t c = 0; c < nCtg; ++c) {
            const uint32_t len = W.ctgMeta[c].hi - W.ctgMeta[c].lo;
            if (len > k + opt.wordStepSize) {
              S.len[nr]     = len;
              S.packOff[nr] = po;
              S.maskOff[nr] = mo;
              W.tmpA[nr]    = c;
              po += (len + 31) / 32 + 4;
              mo += (len + 63) / 64 + 3;
              nr++;
            }
          }
          shU[10] = nr;
        }
        __syncthreads();
        nReads = shU[10];
        for (uint32_t r = normalReadCount + warp; r < nReads; r += nt / 32) {
          const uint32_t c   = W.tmpA[r];
          const uint8_t* src = W.ctgSeq + (size_t)c * ctgBuf + W.ctgMeta[c].lo;
          const uint32_t len = S.len[r];
          const uint32_t nPk = (len + 31) / 32 + 4, nMk = (len + 63) / 64 + 3;
          packRead(src, len, gPacked + S.packOff[r], gMask + S.maskOff[r], nPk, nMk, false, lane);
        }
        __syncthreads();
        PROF_MARK(9);
      }
    }

    if (retry) {  // uniform across the CTA
      if (tid == 0) P.retryList[atomicAdd(P.retryCount, 1u)] = li;
      __syncthreads();
      continue;
    }

    // ================= selectContigs (:722-842) + output (warp 0) =================
    PROF_MARK(10);
    if (warp == 0) {
      mb200_asm_locus Lo;
      memset(&Lo, 0, sizeof(Lo));
      Lo.firstContig    = li * opt.maxAssemblyCount;
      Lo.firstPseudo    = li * maxIter;
      Lo.setWords       = SW;
      Lo.lastWordLength = lastK;
      Lo.status         = status;
      if (status == MB200_OK) {
        uint32_t usedReads = 0, usedPseudo = 0;
        uint32_t aliveMask = (nCtg >= 32) ? 0xffffffffu : ((1u << nCtg) - 1);
        uint32_t nSel      = 0;
        uint32_t pseudoMask;
        {
          const uint32_t lo = lane * 32;
          if (normalReadCount <= lo) pseudoMask = 0xffffffffu;
          else if (normalReadCount >= lo + 32) pseudoMask = 0;
          else pseudoMask = ~((1u << (normalReadCount - lo)) - 1);
        }
        while (aliveMask && nSel < opt.maxAssemblyCount) {
          const uint32_t usedNormal   = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(usedReads)) -
                                      __reduce_add_sync(0xffffffffu, (uint32_t)__popc(usedPseudo));
          const uint32_t unusedNormal = normalReadCount - usedNormal;
          if (unusedNormal < opt.minUnusedReads) break;
          int      sel = -1;
          uint32_t maxSupport = 0, maxLength = 0, removeMask = 0;
          for (uint32_t c = 0; c < nCtg; ++c) {
            if (!((aliveMask >> c) & 1)) continue;
            const uint32_t sup    = W.ctgSets[(size_t)c * 2 * MAX_SETW + lane];
            const uint32_t newSup = sup & ~usedReads;
            const uint32_t newNormal = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(newSup & ~pseudoMask));
            if (newNormal < opt.minSupportReads) {
              removeMask |= 1u << c;
              continue;
            }
            const uint32_t curSup = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(newSup));
            const uint32_t curLen = W.ctgMeta[c].hi - W.ctgMeta[c].lo;
            if (curSup > maxSupport || (curSup == maxSupport && curLen > maxLength)) {
              sel        = (int)c;
              maxSupport = curSup;
              maxLength  = curLen;
            }
          }
          if (maxSupport == 0) break;
          {
            const uint32_t c   = (uint32_t)sel;
            const CtgMeta  m   = W.ctgMeta[c];
            const uint32_t len = m.hi - m.lo;
            unsigned long long off = 0, soff = 0;
            if (lane == 0) {
              off  = atomicAdd(P.A.contigPoolUsed, (unsigned long long)len);
              soff = atomicAdd(P.A.setPoolUsed, (unsigned long long)(2 * SW));
            }
            off  = __shfl_sync(0xffffffffu, off, 0);
            soff = __shfl_sync(0xffffffffu, soff, 0);
            const uint8_t* src = W.ctgSeq + (size_t)c * ctgBuf + m.lo;
            if (off + len <= P.A.contigPoolCap)
              for (uint32_t i = lane; i < len; i += 32) P.A.contigPool[off + i] = src[i];
            if (soff + 2 * SW <= P.A.setPoolCap && lane < SW) {
              P.A.setPool[soff + lane]      = W.ctgSets[(size_t)c * 2 * MAX_SETW + lane];
              P.A.setPool[soff + SW + lane] = W.ctgSets[(size_t)c * 2 * MAX_SETW + MAX_SETW + lane];
            }
            if (lane == 0) {
              mb200_contig o;
              o.seqOff = off; o.seqLen = len; o.seedReadCount = 0;
              o.conservativeBegin = m.consBegin; o.conservativeEnd = m.consEnd;
              o.supportOff = soff; o.rejectOff = soff + SW;
              P.A.contigs[(size_t)Lo.firstContig + nSel] = o;
            }
            const uint32_t sup = W.ctgSets[(size_t)c * 2 * MAX_SETW + lane];
            usedReads |= sup;
            usedPseudo |= sup & pseudoMask;
            removeMask |= 1u << c;
          }
          aliveMask &= ~removeMask;
          nSel++;
        }
        Lo.nContigs = nSel;
        Lo.nPseudo  = nReads - normalReadCount;
        for (uint32_t r = normalReadCount; r < nReads; ++r) {
          const uint32_t     len = S.len[r];
          unsigned long long off = 0;
          if (lane == 0) off = atomicAdd(P.A.contigPoolUsed, (unsigned long long)len);
          off = __shfl_sync(0xffffffffu, off, 0);
          const uint64_t* pk = gPacked + S.packOff[r];
          if (off + len <= P.A.contigPoolCap)
            for (uint32_t i = lane; i < len; i += 32)
              P.A.contigPool[off + i] = "ACGT"[(pk[i >> 5] >> (62 - 2 * (i & 31))) & 3];
          if (lane == 0) {
            mb200_seqview v;
            v.off = off; v.len = len; v.reserved = 0;
            P.A.pseudo[(size_t)Lo.firstPseudo + (r - normalReadCount)] = v;
          }
        }
      } else if (lane == 0) {
        atomicCAS(P.A.errFlag, 0, status);
      }
      if (lane == 0) P.A.loci[li] = Lo;
    }
    __syncthreads();
    PROF_MARK(11);
  }
}

size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// Host pass over the read views of a batch: validates them and collects what sizes the per-CTA workspace.  It depends
// on the first word length only through the window counts, so a staged batch is scanned once and re-used.
int mb200_asm_scan(mb200_ctx* ctx, uint32_t minWordLength, uint32_t nLoci, const uint64_t* hReadOff, const mb200_seqview* hViews,
                   uint64_t poolBytes, mb200_asm_input_stats* st)
{
  // a few host threads: the scan of a 50 000-locus batch (3.2 M views) sits between two device stages
  const unsigned nThreads = nLoci >= 4096 ? 4u : 1u;
  std::vector<mb200_asm_input_stats> part(nThreads);
  std::vector<int>                   bad(nThreads, 0);
  auto work = [&](unsigned t) {
    mb200_asm_input_stats s;
    memset(&s, 0, sizeof(s));
    const uint32_t lo = (uint32_t)((uint64_t)nLoci * t / nThreads), hi = (uint32_t)((uint64_t)nLoci * (t + 1) / nThreads);
    for (uint32_t li = lo; li < hi; ++li) {
      const uint64_t a = hReadOff[li], b = hReadOff[li + 1];
      if (b < a) { bad[t] = 1; break; }
      uint64_t win = 0, pw = 0, mw = 0;
      for (uint64_t r = a; r < b; ++r) {
        const uint32_t len = hViews[r].len;
        if (hViews[r].off + len > poolBytes) { bad[t] = 2; break; }
        if (len >= minWordLength) win += len - minWordLength + 1;
        pw += (len + 31) / 32 + 4;
        mw += (len + 63) / 64 + 3;
        s.totalBases += len;  // read bytes in (+ contigs/bitsets out, added by the caller)
      }
      if (bad[t]) break;
      s.totalWindows += win;
      s.maxReads     = std::max<uint64_t>(s.maxReads, b - a);
      s.maxWin       = std::max<uint64_t>(s.maxWin, win);
      s.maxPackWords = std::max<uint64_t>(s.maxPackWords, pw);
      s.maxMaskWords = std::max<uint64_t>(s.maxMaskWords, mw);
    }
    part[t] = s;
  };
  if (nThreads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nThreads; ++t) th.emplace_back(work, t);
    for (std::thread& x : th) x.join();
  }
  mb200_asm_input_stats s;
  memset(&s, 0, sizeof(s));
  s.minWordLength = minWordLength;
  for (unsigned t = 0; t < nThreads; ++t) {
    if (bad[t] == 1) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "readOff not monotone");
    if (bad[t] == 2) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "read view outside the pool");
    s.totalBases += part[t].totalBases;
    s.totalWindows += part[t].totalWindows;
    s.maxReads     = std::max(s.maxReads, part[t].maxReads);
    s.maxWin       = std::max(s.maxWin, part[t].maxWin);
    s.maxPackWords = std::max(s.maxPackWords, part[t].maxPackWords);
    s.maxMaskWords = std::max(s.maxMaskWords, part[t].maxMaskWords);
  }
  s.valid = 1;
  *st     = s;
  return MB200_OK;
}

// Device-level run: inputs (pool, readOff, views) and all outputs live on the device; `st` is the host scan of the
// same batch (mb200_asm_scan with opts->minWordLength).
int mb200_asm_run(mb200_ctx* ctx, const mb200_asm_opts* opts, uint32_t nLoci, const mb200_asm_input_stats* st,
                  const uint8_t* dPool, const uint64_t* dReadOff,
                  const mb200_seqview* dViews, mb200_asm_locus* dLoci, mb200_contig* dContigs, mb200_seqview* dPseudo,
                  uint8_t* dContigPool, uint64_t contigPoolCap, unsigned long long* dContigPoolUsed, uint32_t* dSetPool,
                  uint64_t setPoolCap, unsigned long long* dSetPoolUsed, uint32_t* dLociCounter, int* dErrFlag,
                  uint64_t* cellsOut, uint64_t* algoBytesOut)
{
  if (opts->minCoverage < 1 || opts->wordStepSize < 1 || opts->minWordLength < 2 || opts->maxWordLength > 96 ||
      opts->maxAssemblyCount < 1 || opts->maxAssemblyCount > 16)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "assembler options outside the supported range (k in [2,96], maxAssemblyCount in [1,16])");
  if (!st || !st->valid || st->minWordLength != opts->minWordLength) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "stale input scan");
  const uint32_t A            = opts->maxAssemblyCount;
  const uint32_t maxContigLen = ctx->cfg.maxContigLen ? ctx->cfg.maxContigLen : 2048;
  if (st->maxReads + 2 * A > 1024) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "more than 1024 reads (incl. pseudo reads) in one locus");
  const uint32_t maxReads = (uint32_t)st->maxReads, maxPackWords = (uint32_t)st->maxPackWords, maxMaskWords = (uint32_t)st->maxMaskWords;
  const uint64_t maxWin = st->maxWin;
  if (algoBytesOut) *algoBytesOut += st->totalBases;
  if (cellsOut) *cellsOut += st->totalWindows;
  // ---- workspace layout (per CTA)
  WorkLayout L;
  memset(&L, 0, sizeof(L));
  L.maxReads  = maxReads + 2 * A;
  const uint64_t nWinMax64 = maxWin + (uint64_t)2 * A * maxContigLen;
  if (nWinMax64 > 0x3fffffffull) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "locus too large");
  L.nWinMax   = (uint32_t)nWinMax64 + 64;
  L.slotsMax  = 64;
  while (L.slotsMax < 2 * L.nWinMax) L.slotsMax <<= 1;
  L.packWords = maxPackWords + 2 * A * ((maxContigLen + 31) / 32 + 4) + 8;
  L.maskWords = maxMaskWords + 2 * A * ((maxContigLen + 63) / 64 + 3) + 8;
  L.bktMax    = 3 * L.nWinMax + 64;
  size_t off = 0;
  auto   put = [&](size_t& field, size_t bytes) {
    field = off;
    off   = alignUp(off + bytes, 256);
  };
  const size_t NW = L.nWinMax;
  put(L.readLen, 4 * (size_t)(L.maxReads + 1));
  put(L.readPackOff, 4 * (size_t)(L.maxReads + 1));
  put(L.readMaskOff, 4 * (size_t)(L.maxReads + 1));
  put(L.winStart, 4 * (size_t)(L.maxReads + 2));
  put(L.packed, 8 * (size_t)L.packWords);
  put(L.nmask, 8 * (size_t)L.maskWords);
  put(L.key0, 8 * NW);
  put(L.key1, 8 * NW);
  put(L.key2, 8 * NW);
  put(L.count, 4 * NW);
  put(L.firstRead, 4 * NW);
  put(L.bits, 4 * NW * (size_t)((L.maxReads + 31) / 32));
  put(L.slot, 4 * (size_t)L.slotsMax);
  put(L.wordList, 4 * NW);
  put(L.succ, 16 * NW);
  put(L.pred, 16 * NW);
  put(L.indeg, 4 * NW);
  put(L.flags, NW + 8);
  put(L.inContig, NW + 8);
  put(L.hashv, 8 * NW);
  put(L.seqA, 4 * NW);
  put(L.arr, 4 * NW);
  put(L.pos, 4 * NW);
  put(L.bktHead, 4 * (size_t)L.bktMax);
  put(L.bktMin, 4 * (size_t)L.bktMax);
  put(L.bktCnt, 4 * (size_t)L.bktMax);
  put(L.chainNext, 4 * NW);
  put(L.gs, 4 * (NW + L.maxReads + 2));
  put(L.orderA, 4 * NW);
  put(L.tmpA, 4 * (NW + L.maxReads + 2));
  put(L.idx, 4 * NW);
  put(L.low, 4 * NW);
  // wstack and frame are contiguous: together they double as a uint64[nWin] scratch (hash2)
  L.wstack = off;
  off += 4 * NW;
  L.frame = off;
  off     = alignUp(off + 4 * NW, 256);
  put(L.ctgSeq, (size_t)2 * A * 2 * maxContigLen);
  put(L.ctgSets, 4 * (size_t)2 * A * 2 * MAX_SETW);
  put(L.ctgMeta, sizeof(CtgMeta) * (size_t)2 * A);
  L.total = alignUp(off, 256);

  int ctasPerSm = 0;
  MB200_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, assembleKernel, ASM_THREADS, 0));
  if (ctasPerSm < 1) ctasPerSm = 1;
  if (ctasPerSm > 32) ctasPerSm = 32;
  if (ctx->cfg.ctasPerSm && (int)ctx->cfg.ctasPerSm < ctasPerSm) ctasPerSm = (int)ctx->cfg.ctasPerSm;
  uint32_t grid = (uint32_t)(ctx->smCount * ctasPerSm);
  if (grid > nLoci) grid = nLoci;
  const size_t budget = ctx->cfg.scratchBytes ? (size_t)ctx->cfg.scratchBytes : ((size_t)48 << 30);
  while (grid > 1 && (size_t)grid * L.total > budget) grid = (grid + 1) / 2;
  int rc;
  // On-chip kernel: 3 loci per SM at a sixth of the per-locus latency of the global-memory kernel (32 loci per SM).  It
  // wins while the batch is small (a few waves of the on-chip grid: the per-call shim, the reference's per-candidate
  // call pattern), the global-memory kernel wins on throughput for large batches.
  const uint32_t chipMaxLoci = ctx->asmChipMaxLoci ? ctx->asmChipMaxLoci : (uint32_t)ctx->smCount * 3;
  const bool wantChip = !ctx->legacyAssembler && (ctx->forceChipAssembler || nLoci <= chipMaxLoci) &&
                        (uint64_t)opts->minCoverage * (2 * A) + 128 <= 0xffffu;
  if (!wantChip && (rc = mb200_reserve(ctx, ctx->dAsmWork, (size_t)grid * L.total)) != MB200_OK) return rc;

  AsmParams P;
  memset(&P, 0, sizeof(P));
  P.opt            = *opts;
  P.pool           = dPool;
  P.readOff        = dReadOff;
  P.views          = dViews;
  P.nLoci          = nLoci;
  P.lociCounter    = dLociCounter;
  P.errFlag        = dErrFlag;
  P.contigPoolUsed = dContigPoolUsed;
  P.setPoolUsed    = dSetPoolUsed;
  P.work           = reinterpret_cast<uint8_t*>(ctx->dAsmWork.ptr);
  P.L              = L;
  P.maxContigLen   = maxContigLen;
  P.loci           = dLoci;
  P.contigs        = dContigs;
  P.pseudo         = dPseudo;
  P.contigPool     = dContigPool;
  P.contigPoolCap  = contigPoolCap;
  P.setPool        = dSetPool;
  P.setPoolCap     = setPoolCap;
  // ---- on-chip kernel first (every locus that fits), then the kernel above over the retry list
  if (wantChip && nLoci > 0) {
    ChipLayout C;
    memset(&C, 0, sizeof(C));
    int maxSmem = 0, smemPerSm = 0;
    MB200_CUDA_CHECK(ctx, cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    MB200_CUDA_CHECK(ctx, cudaDeviceGetAttribute(&smemPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    uint32_t chipCtas = 4;
    if (const char* e = getenv("MB200_ASM_CHIP_CTAS")) chipCtas = (uint32_t)std::max(1, atoi(e));
    C.SWc = (uint32_t)std::min<uint64_t>(4, std::max<uint64_t>(1, (st->maxReads + 2 * A + 31) / 32));
    const uint32_t RC = C.SWc * 32;
    C.SPK = (uint32_t)std::min<uint64_t>(640, (st->maxPackWords + 7) / 8 * 8);
    C.SL  = 2048;
    const size_t budgetSmem = std::min<size_t>((size_t)maxSmem, (size_t)smemPerSm / chipCtas - 1024) - 64;
    const size_t fixedSmem  = 8 * (size_t)C.SPK + 4 * (size_t)C.SL + 4 * (size_t)RC * 3 + 4 * (size_t)(RC + 4);
    const size_t perWord    = 4 + 4 * (size_t)C.SWc + 8 + 8 + 2 + 2 + 3;
    uint32_t     Wc         = fixedSmem < budgetSmem ? (uint32_t)((budgetSmem - fixedSmem) / perWord) : 0;
    if (const char* e = getenv("MB200_ASM_CHIP_WC")) Wc = (uint32_t)atoi(e);
    Wc &= ~7u;
    if (Wc + CHIP_THREADS + 1 > C.SL) Wc = (C.SL - CHIP_THREADS - 1) & ~7u;
    C.Wc        = Wc;
    C.smemBytes = (uint32_t)(fixedSmem + perWord * Wc);
    if (Wc >= 64 && C.smemBytes <= (uint32_t)maxSmem) {
      C.maxReads  = L.maxReads;
      C.nWinMax   = L.nWinMax;
      C.packWords = L.packWords;
      C.maskWords = L.maskWords;
      C.bktMax    = 64;
      for (int lv = 0; lv < 22; ++lv) {
        static const uint32_t stl[22] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933, 351061, 712697,
                                         1447153, 2938679, 5967347, 12117689, 24607243, 49969847};
        C.bktMax = stl[lv] + 64;
        if (stl[lv] >= Wc) break;
      }
      size_t o2 = 0;
      auto   put2 = [&](size_t& field, size_t bytes) {
        field = o2;
        o2    = alignUp(o2 + bytes, 256);
      };
      const size_t WW = (size_t)Wc + C.maxReads + 8;
      put2(C.packed, 8 * (size_t)C.packWords);
      put2(C.nmask, 8 * (size_t)C.maskWords);
      put2(C.winSlot, 2 * (size_t)C.nWinMax + 64);
      put2(C.hashv, 8 * WW);
      put2(C.hash2, 8 * WW);
      put2(C.seqA, 4 * WW);
      put2(C.arr, 4 * WW);
      put2(C.pos, 4 * WW);
      put2(C.bktHead, 4 * (size_t)C.bktMax);
      put2(C.bktMin, 4 * (size_t)C.bktMax);
      put2(C.bktCnt, 4 * (size_t)C.bktMax);
      put2(C.chainNext, 4 * WW);
      put2(C.gs, 4 * WW);
      put2(C.orderA, 4 * WW);
      put2(C.tmpA, 4 * WW);
      put2(C.idx, 4 * WW);
      put2(C.low, 4 * WW);
      put2(C.wstack, 4 * WW);
      put2(C.frame, 4 * WW);
      put2(C.ctgSeq, (size_t)2 * A * 2 * maxContigLen);
      put2(C.ctgSets, 4 * (size_t)2 * A * 2 * MAX_SETW);
      put2(C.ctgMeta, sizeof(CtgMeta) * (size_t)2 * A);
      C.total = alignUp(o2, 256);

      MB200_CUDA_CHECK(ctx, cudaFuncSetAttribute(assembleChipKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C.smemBytes));
      int chipPerSm = 0;
      MB200_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&chipPerSm, assembleChipKernel, CHIP_THREADS, C.smemBytes));
      if (chipPerSm < 1) chipPerSm = 1;
      uint32_t chipGrid = (uint32_t)(ctx->smCount * chipPerSm);
      if (chipGrid > nLoci) chipGrid = nLoci;
      // the retry pass keeps a small persistent grid (its per-CTA workspace is large)
      uint32_t retryGrid = std::min<uint32_t>(grid, (uint32_t)ctx->smCount * 8);
      const size_t chipWork = (size_t)chipGrid * C.total;
      const size_t listOff  = alignUp(chipWork, 256);
      const size_t cntOff   = listOff + alignUp(4 * (size_t)nLoci, 256);
      if ((rc = mb200_reserve(ctx, ctx->dAsmChip, cntOff + 256)) != MB200_OK) return rc;
      if ((rc = mb200_reserve(ctx, ctx->dAsmWork, (size_t)retryGrid * L.total)) != MB200_OK) return rc;
      uint8_t*  chipBase = reinterpret_cast<uint8_t*>(ctx->dAsmChip.ptr);
      uint32_t* counters = reinterpret_cast<uint32_t*>(chipBase + cntOff);  // [0] retry count, [1] retry pass work counter
      MB200_CUDA_CHECK(ctx, cudaMemsetAsync(counters, 0, 256, ctx->stream));
      ChipParams Q;
      memset(&Q, 0, sizeof(Q));
      Q.A          = P;
      Q.A.work     = nullptr;
      Q.C          = C;
      Q.work       = chipBase;
      Q.retryList  = reinterpret_cast<uint32_t*>(chipBase + listOff);
      Q.retryCount = counters;
      assembleChipKernel<<<chipGrid, CHIP_THREADS, C.smemBytes, ctx->stream>>>(Q);
      MB200_CUDA_CHECK(ctx, cudaGetLastError());
      ctx->timing.launches++;
      P.work          = reinterpret_cast<uint8_t*>(ctx->dAsmWork.ptr);
      P.lociList      = Q.retryList;
      P.lociListCount = counters;
      P.lociCounter   = counters + 1;
      assembleKernel<<<retryGrid, ASM_THREADS, 0, ctx->stream>>>(P);
      MB200_CUDA_CHECK(ctx, cudaGetLastError());
      ctx->timing.launches++;
      if (getenv("MB200_ASM_DEBUG")) {  // diagnostics only: forces a synchronisation
        uint32_t hc[2] = {0, 0};
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[mb200 asm] on-chip kernel: grid %u x %d threads, %u B smem (%d CTAs/SM), Wc %u SWc %u; %u of %u loci retried on the global-memory kernel\n",
                chipGrid, CHIP_THREADS, C.smemBytes, chipPerSm, C.Wc, C.SWc, hc[0], nLoci);
      }
      return MB200_OK;
    }
  }
  assembleKernel<<<grid, ASM_THREADS, 0, ctx->stream>>>(P);
  MB200_CUDA_CHECK(ctx, cudaGetLastError());
  ctx->timing.launches++;
  return MB200_OK;
}

#ifdef MB200_ASM_PROFILE
extern "C" __attribute__((visibility("default"))) int mb200_debug_asm_profile(unsigned long long* out32, int reset)
{
  if (cudaMemcpyFromSymbol(out32, gAsmProf, sizeof(unsigned long long) * 32) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[32] = {0};
    if (cudaMemcpyToSymbol(gAsmProf, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif

extern "C" int mb200_assemble_batch(mb200_ctx* ctx, const mb200_asm_opts* opts, uint32_t nLoci,
                                    const uint64_t* readOff, const mb200_seqview* readViews, const uint8_t* pool,
                                    uint64_t poolBytes, mb200_asm_locus* loci, mb200_contig* contigs,
                                    uint64_t contigCap, mb200_seqview* pseudoReads, uint64_t pseudoCap,
                                    uint8_t* contigPool, uint64_t contigPoolCap, uint32_t* setPool,
                                    uint64_t setPoolCap)
{
  if (!ctx) return MB200_ERR_BAD_ARG;
  if (!opts || (nLoci && (!readOff || !loci || !contigs || !pseudoReads || !contigPool || !setPool)))
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");
  ctx->timing = mb200_timing();
  ctx->candSv.valid = false;  // shares device buffers with the fused path's resident results
  if (nLoci == 0) return MB200_OK;
  const uint32_t A = opts->maxAssemblyCount;
  if (A < 1 || A > 16) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "maxAssemblyCount outside [1,16]");
  if (contigCap < (uint64_t)nLoci * A || pseudoCap < (uint64_t)nLoci * 2 * A)
    return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "contig / pseudo-read record arrays too small");
  MB200_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const uint64_t nReadsTotal = readOff[nLoci];
  if (nReadsTotal && (!readViews || !pool)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "null argument");

  int rc;
  if ((rc = mb200_reserve(ctx, ctx->dPool, poolBytes + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dReadOff, 8 * (size_t)(nLoci + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dViews, sizeof(mb200_seqview) * (size_t)(nReadsTotal + 1))) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut0, sizeof(mb200_asm_locus) * (size_t)nLoci)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut1, sizeof(mb200_contig) * (size_t)nLoci * A)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut2, sizeof(mb200_seqview) * (size_t)nLoci * 2 * A)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut3, contigPoolCap + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dAsmOut4, 4 * (size_t)setPoolCap + 64)) != MB200_OK) return rc;
  if ((rc = mb200_reserve(ctx, ctx->dMisc, 256)) != MB200_OK) return rc;

  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dMisc.ptr, 0, 256, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dPool.ptr, pool, poolBytes, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dReadOff.ptr, readOff, 8 * (size_t)(nLoci + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (nReadsTotal)
    MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dViews.ptr, readViews, sizeof(mb200_seqview) * (size_t)nReadsTotal, cudaMemcpyHostToDevice, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dAsmOut1.ptr, 0, sizeof(mb200_contig) * (size_t)nLoci * A, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dAsmOut2.ptr, 0, sizeof(mb200_seqview) * (size_t)nLoci * 2 * A, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

  uint32_t* misc = reinterpret_cast<uint32_t*>(ctx->dMisc.ptr);
  mb200_asm_input_stats st;
  if ((rc = mb200_asm_scan(ctx, opts->minWordLength, nLoci, readOff, readViews, poolBytes, &st)) != MB200_OK) return rc;
  rc = mb200_asm_run(ctx, opts, nLoci, &st, reinterpret_cast<const uint8_t*>(ctx->dPool.ptr),
                     reinterpret_cast<const uint64_t*>(ctx->dReadOff.ptr), reinterpret_cast<const mb200_seqview*>(ctx->dViews.ptr),
                     reinterpret_cast<mb200_asm_locus*>(ctx->dAsmOut0.ptr), reinterpret_cast<mb200_contig*>(ctx->dAsmOut1.ptr),
                     reinterpret_cast<mb200_seqview*>(ctx->dAsmOut2.ptr), reinterpret_cast<uint8_t*>(ctx->dAsmOut3.ptr),
                     contigPoolCap, reinterpret_cast<unsigned long long*>(misc + 8), reinterpret_cast<uint32_t*>(ctx->dAsmOut4.ptr),
                     setPoolCap, reinterpret_cast<unsigned long long*>(misc + 10), misc + 0, reinterpret_cast<int*>(misc + 4),
                     &ctx->timing.cells, &ctx->timing.algoBytes);
  if (rc != MB200_OK) return rc;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));

  uint32_t hMisc[16];
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(hMisc, ctx->dMisc.ptr, sizeof(hMisc), cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(loci, ctx->dAsmOut0.ptr, sizeof(mb200_asm_locus) * (size_t)nLoci, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(contigs, ctx->dAsmOut1.ptr, sizeof(mb200_contig) * (size_t)nLoci * A, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(pseudoReads, ctx->dAsmOut2.ptr, sizeof(mb200_seqview) * (size_t)nLoci * 2 * A, cudaMemcpyDeviceToHost, ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long usedBytes, usedSets;
  memcpy(&usedBytes, hMisc + 8, 8);
  memcpy(&usedSets, hMisc + 10, 8);
  if (hMisc[4] != 0) return mb200_fail(ctx, (int)hMisc[4], "assembler: a locus failed (bad base / workspace overflow / too many reads)");
  if (usedBytes > contigPoolCap || usedSets > setPoolCap) return mb200_fail(ctx, MB200_ERR_POOL_OVERFLOW, "assembler output pools too small");
  if (usedBytes) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(contigPool, ctx->dAsmOut3.ptr, usedBytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (usedSets) MB200_CUDA_CHECK(ctx, cudaMemcpyAsync(setPool, ctx->dAsmOut4.ptr, 4 * usedSets, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->timing.algoBytes += usedBytes + 4 * usedSets;
  MB200_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  MB200_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->timing.h2dMs, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing.kernelMs, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing.d2hMs, ctx->ev[2], ctx->ev[3]);
  return MB200_OK;
}
