/*
 * orc_assemble.c -- TEST INFRASTRUCTURE ONLY (see manta_oracle.h).
 *
 * CPU restatement of the reference's iterative-k de-Bruijn assembler:
 *
 *   runIterativeAssembler   lib/assembly/IterativeAssembler.cpp:844-931
 *   buildContigs            :644-720
 *   getKmerCounts           :506-550
 *   getRepeatKmers          :627-642   searchRepeats (recursive Tarjan) :555-625
 *   walk                    :149-501
 *   selectContigs           :722-842
 *
 * The single place where the reference depends on a hash-table iteration order --
 * getRepeatKmers iterates a std::unordered_map<std::string,...> that was itself filled while iterating
 * another std::unordered_map -- is reproduced by emulating libstdc++'s _Hashtable: std::hash<std::string>
 * (_Hash_bytes, murmur-style, seed 0xc70f6907), _Prime_rehash_policy growth (13, 29, 59, 127, ...),
 * insert-at-bucket-begin and the node-relinking rehash.  tests/ pin the emulation against a real
 * std::unordered_map through oracle/_ref (ref_unordered_order).
 *
 * Read sets are bitsets; every std::set<unsigned> operation of the reference is the obvious bit operation.
 */
#include <stdlib.h>
#include <string.h>

#include "manta_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* libstdc++ emulation                                                                         */

/* std::_Hash_bytes (libstdc++-v3/libsupc++/hash_bytes.cc, 64-bit size_t) */
static uint64_t stl_hash_bytes(const uint8_t* buf, size_t len)
{
  const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
  const uint64_t seed = 0xc70f6907UL;
  const size_t   alen = len & ~(size_t)7;
  uint64_t       h    = seed ^ (len * mul);
  for (size_t p = 0; p < alen; p += 8) {
    uint64_t d;
    memcpy(&d, buf + p, 8);
    d *= mul;
    d ^= d >> 47;
    d *= mul;
    h ^= d;
    h *= mul;
  }
  if (len & 7) {
    uint64_t d = 0;
    for (int n = (int)(len & 7) - 1; n >= 0; --n) d = (d << 8) + buf[alen + n];
    h ^= d;
    h *= mul;
  }
  h ^= h >> 47;
  h *= mul;
  h ^= h >> 47;
  return h;
}

/* bucket counts std::unordered_map walks through when filled one element at a time
 * (_Prime_rehash_policy::_M_need_rehash / _M_next_bkt, max_load_factor 1): the smallest entry of
 * libstdc++'s __prime_list that is >= 2 * previous.  Measured on this image's libstdc++ (gcc 13.3) by
 * tests/golden/make_golden.py and pinned by tests/test_oracle_vs_ref.py. */
static const uint64_t STL_BUCKETS[] = {13ull,        29ull,        59ull,        127ull,       257ull,      541ull,
                                       1109ull,      2357ull,      5087ull,      10273ull,     20753ull,    42043ull,
                                       85229ull,     172933ull,    351061ull,    712697ull,    1447153ull,  2938679ull,
                                       5967347ull,   12117689ull,  24607243ull,  49969847ull,  101473717ull};
#define STL_NBUCKETS (sizeof(STL_BUCKETS) / sizeof(STL_BUCKETS[0]))

/* Emulated std::unordered_map<std::string,T>: insert n keys (hashes given in insertion order) one by one
 * with operator[], then report the iteration order (as indices into the insertion sequence). */
static void stl_unordered_order(const uint64_t* hash, uint32_t n, uint32_t* order)
{
  enum { NONE = -1, BEFORE_BEGIN = -2 };
  int32_t* next = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
  int32_t  head = NONE; /* _M_before_begin._M_nxt */
  uint64_t nb   = 1;    /* bucket count */
  int32_t* bkt  = (int32_t*)malloc(sizeof(int32_t));
  bkt[0]        = NONE;
  uint64_t nextResize = 0;
  uint32_t growth     = 0;

  for (uint32_t e = 0; e < n; ++e) {
    /* _M_need_rehash(nb, e, 1) */
    if ((uint64_t)e + 1 > nextResize) {
      uint64_t minB = (uint64_t)e + 1;
      if (nextResize == 0 && minB < 11) minB = 11;
      if (minB >= nb) {
        if (growth >= STL_NBUCKETS) abort();
        const uint64_t newNb = STL_BUCKETS[growth++];
        nextResize           = newNb;
        /* _M_rehash_aux(newNb, unique) */
        int32_t* nbkt = (int32_t*)malloc(sizeof(int32_t) * newNb);
        for (uint64_t b = 0; b < newNb; ++b) nbkt[b] = NONE;
        int32_t  p         = head;
        uint64_t bbeginBkt = 0;
        head               = NONE;
        while (p != NONE) {
          const int32_t  pn = next[p];
          const uint64_t b  = hash[p] % newNb;
          if (nbkt[b] == NONE) {
            next[p] = head;
            head    = p;
            nbkt[b] = BEFORE_BEGIN;
            if (next[p] != NONE) nbkt[bbeginBkt] = p;
            bbeginBkt = b;
          } else {
            const int32_t before = nbkt[b];
            if (before == BEFORE_BEGIN) {
              next[p] = head;
              head    = p;
            } else {
              next[p]      = next[before];
              next[before] = p;
            }
          }
          p = pn;
        }
        free(bkt);
        bkt = nbkt;
        nb  = newNb;
      } else {
        nextResize = nb;
      }
    }
    /* _M_insert_bucket_begin */
    const uint64_t b = hash[e] % nb;
    if (bkt[b] != NONE) {
      const int32_t before = bkt[b];
      if (before == BEFORE_BEGIN) {
        next[e] = head;
        head    = (int32_t)e;
      } else {
        next[e]      = next[before];
        next[before] = (int32_t)e;
      }
    } else {
      next[e] = head;
      head    = (int32_t)e;
      if (next[e] != NONE) bkt[hash[next[e]] % nb] = (int32_t)e;
      bkt[b] = BEFORE_BEGIN;
    }
  }
  uint32_t k = 0;
  for (int32_t p = head; p != NONE; p = next[p]) order[k++] = (uint32_t)p;
  free(next);
  free(bkt);
}

/* exported for the emulation test: keys are n strings of keyLen bytes, concatenated */
void orc_unordered_order(uint32_t n, const uint8_t* keys, uint32_t keyLen, uint32_t* order)
{
  uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
  for (uint32_t i = 0; i < n; ++i) h[i] = stl_hash_bytes(keys + (size_t)i * keyLen, keyLen);
  stl_unordered_order(h, n, order);
  free(h);
}

/* ------------------------------------------------------------------------------------------ */
/* bitsets of read indices                                                                     */
typedef uint64_t bw_t;

static inline void bs_zero(bw_t* a, uint32_t W) { memset(a, 0, sizeof(bw_t) * W); }
static inline void bs_copy(bw_t* a, const bw_t* b, uint32_t W) { memcpy(a, b, sizeof(bw_t) * W); }
static inline int  bs_empty(const bw_t* a, uint32_t W)
{
  for (uint32_t i = 0; i < W; ++i)
    if (a[i]) return 0;
  return 1;
}
static inline uint32_t bs_count(const bw_t* a, uint32_t W)
{
  uint32_t c = 0;
  for (uint32_t i = 0; i < W; ++i) c += (uint32_t)__builtin_popcountll(a[i]);
  return c;
}
static inline void bs_set(bw_t* a, uint32_t i) { a[i >> 6] |= ((bw_t)1 << (i & 63)); }
static inline int  bs_test(const bw_t* a, uint32_t i) { return (int)((a[i >> 6] >> (i & 63)) & 1); }

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  uint8_t* seq;
  uint32_t len;
  int      isPseudo;
} read_t;

typedef struct {
  uint8_t* seq; /* malloc'd */
  uint32_t len;
  bw_t*    support;
  bw_t*    reject;
  int32_t  consBegin, consEnd;
} contig_t;

typedef struct {
  uint32_t k, W;
  uint32_t nWords, capWords;
  uint8_t* text;  /* nWords * k bytes: word strings in wordCount insertion order */
  uint32_t* count; /* wordCount */
  bw_t*    reads; /* wordSupportReads, nWords * W */
  uint64_t* hash; /* std::hash<std::string> of each word */
  /* lookup table (open addressing on the same hash) */
  int32_t* slot;
  uint32_t nSlot;
  uint8_t* unused; /* member of unusedWords */
  uint8_t* repeat; /* member of repeatWords */
  uint32_t* inContig; /* generation stamp: member of wordsInContig of the current walk */
  uint32_t gen;
} wordtab_t;

static const uint8_t* wt_word(const wordtab_t* t, uint32_t id) { return t->text + (size_t)id * t->k; }

static int32_t wt_find(const wordtab_t* t, const uint8_t* w)
{
  uint64_t h = stl_hash_bytes(w, t->k);
  for (uint32_t s = (uint32_t)(h % t->nSlot);; s = (s + 1 == t->nSlot) ? 0 : s + 1) {
    const int32_t id = t->slot[s];
    if (id < 0) return -1;
    if (t->hash[id] == h && memcmp(wt_word(t, (uint32_t)id), w, t->k) == 0) return id;
  }
}

static uint32_t wt_insert(wordtab_t* t, const uint8_t* w)
{
  uint64_t h = stl_hash_bytes(w, t->k);
  uint32_t s = (uint32_t)(h % t->nSlot);
  for (;; s = (s + 1 == t->nSlot) ? 0 : s + 1) {
    const int32_t id = t->slot[s];
    if (id < 0) break;
    if (t->hash[id] == h && memcmp(wt_word(t, (uint32_t)id), w, t->k) == 0) return (uint32_t)id;
  }
  const uint32_t id = t->nWords++;
  memcpy(t->text + (size_t)id * t->k, w, t->k);
  t->hash[id] = h;
  t->slot[s]  = (int32_t)id;
  return id;
}

static void wt_init(wordtab_t* t, uint32_t k, uint32_t W, uint32_t maxWords)
{
  memset(t, 0, sizeof(*t));
  t->k        = k;
  t->W        = W;
  t->capWords = maxWords ? maxWords : 1;
  t->text     = (uint8_t*)malloc((size_t)t->capWords * k);
  t->count    = (uint32_t*)calloc(t->capWords, sizeof(uint32_t));
  t->reads    = (bw_t*)calloc((size_t)t->capWords * W, sizeof(bw_t));
  t->hash     = (uint64_t*)malloc(sizeof(uint64_t) * t->capWords);
  t->nSlot    = 2 * t->capWords + 7;
  t->slot     = (int32_t*)malloc(sizeof(int32_t) * t->nSlot);
  for (uint32_t i = 0; i < t->nSlot; ++i) t->slot[i] = -1;
  t->unused   = (uint8_t*)calloc(t->capWords, 1);
  t->repeat   = (uint8_t*)calloc(t->capWords, 1);
  t->inContig = (uint32_t*)calloc(t->capWords, sizeof(uint32_t));
  t->gen      = 0;
}

static void wt_free(wordtab_t* t)
{
  free(t->text);
  free(t->count);
  free(t->reads);
  free(t->hash);
  free(t->slot);
  free(t->unused);
  free(t->repeat);
  free(t->inContig);
}

/* getKmerCounts  IterativeAssembler.cpp:506-550 */
/* comparator context: thread-local, orc_candidates_mt runs the assembler on many threads */
static __thread uint32_t       g_sortK;
static __thread const uint8_t* g_sortSeq;
static int cmp_offsets(const void* a, const void* b)
{
  return memcmp(g_sortSeq + *(const uint32_t*)a, g_sortSeq + *(const uint32_t*)b, g_sortK);
}

static void get_kmer_counts(const mb200_asm_opts* opt, const read_t* reads, uint32_t nReads, wordtab_t* t)
{
  const uint32_t k   = t->k;
  uint32_t*      off = NULL;
  uint32_t       capOff = 0;
  for (uint32_t r = 0; r < nReads; ++r) {
    const read_t* rd = &reads[r];
    if (rd->len < k) continue;
    const uint32_t nWin = rd->len - k + 1;
    if (nWin > capOff) {
      capOff = nWin;
      off    = (uint32_t*)realloc(off, sizeof(uint32_t) * capOff);
    }
    /* windows without 'N' */
    uint32_t n = 0;
    int32_t  lastN = -1;
    for (uint32_t j = 0; j < k - 1 && j < rd->len; ++j)
      if (rd->seq[j] == 'N') lastN = (int32_t)j;
    for (uint32_t j = 0; j < nWin; ++j) {
      if (rd->seq[j + k - 1] == 'N') lastN = (int32_t)(j + k - 1);
      if (lastN >= (int32_t)j) continue;
      off[n++] = j;
    }
    /* std::set<std::string> readWords: distinct words in lexicographic order */
    g_sortK   = k;
    g_sortSeq = rd->seq;
    qsort(off, n, sizeof(uint32_t), cmp_offsets);
    const uint32_t add = rd->isPseudo ? opt->minCoverage : 1;
    for (uint32_t i = 0; i < n; ++i) {
      if (i && memcmp(rd->seq + off[i], rd->seq + off[i - 1], k) == 0) continue;
      const uint32_t id = wt_insert(t, rd->seq + off[i]);
      t->count[id] += add;
      bs_set(t->reads + (size_t)id * t->W, r);
    }
  }
  free(off);
}

/* getRepeatKmers + searchRepeats  IterativeAssembler.cpp:555-642 (recursion unrolled on an explicit stack) */
static void get_repeat_kmers(wordtab_t* t)
{
  const uint32_t n = t->nWords;
  const uint32_t k = t->k;
  if (!n) return;
  /* iteration order of wordCount, then of wordIndices (filled while iterating wordCount) */
  uint32_t* order1 = (uint32_t*)malloc(sizeof(uint32_t) * n);
  uint32_t* order2 = (uint32_t*)malloc(sizeof(uint32_t) * n);
  uint64_t* h2     = (uint64_t*)malloc(sizeof(uint64_t) * n);
  stl_unordered_order(t->hash, n, order1);
  for (uint32_t i = 0; i < n; ++i) h2[i] = t->hash[order1[i]];
  stl_unordered_order(h2, n, order2);
  /* root sequence = order1[order2[i]] */

  uint32_t* idx      = (uint32_t*)calloc(n, sizeof(uint32_t)); /* wordIndices[w].first  */
  uint32_t* low      = (uint32_t*)calloc(n, sizeof(uint32_t)); /* wordIndices[w].second */
  uint8_t*  onStack  = (uint8_t*)calloc(n, 1);
  uint32_t* wstack   = (uint32_t*)malloc(sizeof(uint32_t) * n); /* wordStack */
  uint32_t  nStack   = 0;
  uint32_t* frameW   = (uint32_t*)malloc(sizeof(uint32_t) * n); /* call stack: word */
  uint8_t*  frameS   = (uint8_t*)malloc(n);                     /* call stack: next symbol to try */
  uint8_t*  key      = (uint8_t*)malloc(k);
  static const char alphabet[4] = {'A', 'C', 'G', 'T'};

  uint32_t index = 1;
  for (uint32_t ri = 0; ri < n; ++ri) {
    const uint32_t root = order1[order2[ri]];
    if (idx[root] != 0) continue;
    uint32_t depth = 0;
    frameW[0]      = root;
    frameS[0]      = 0;
    idx[root] = low[root] = index++;
    wstack[nStack++]      = root;
    onStack[root]         = 1;
    depth                 = 1;
    while (depth) {
      const uint32_t w = frameW[depth - 1];
      if (frameS[depth - 1] < 4) {
        const uint8_t sym = (uint8_t)alphabet[frameS[depth - 1]++];
        memcpy(key, wt_word(t, w) + 1, k - 1);
        key[k - 1] = sym;
        if (memcmp(key, wt_word(t, w), k) == 0) { /* homopolymer */
          t->repeat[w] = 1;
          continue;
        }
        const int32_t nw = wt_find(t, key);
        if (nw < 0) continue;
        if (idx[nw] == 0) {
          /* recurse */
          frameW[depth] = (uint32_t)nw;
          frameS[depth] = 0;
          depth++;
          idx[nw] = low[nw] = index++;
          wstack[nStack++]  = (uint32_t)nw;
          onStack[nw]       = 1;
        } else if (onStack[nw]) {
          if (idx[nw] < low[w]) low[w] = idx[nw];
        }
      } else {
        /* all successors done: root test, then return to the caller frame */
        if (low[w] == idx[w]) {
          const uint32_t lastWord = wstack[nStack - 1];
          if (lastWord == w) {
            nStack--;
            onStack[w] = 0;
          } else {
            const int isSmallCircle = ((idx[lastWord] - idx[w]) <= 50);
            for (;;) {
              const uint32_t rw = wstack[--nStack];
              onStack[rw]       = 0;
              if (isSmallCircle) t->repeat[rw] = 1;
              if (rw == w) break;
            }
          }
        }
        depth--;
        if (depth) {
          const uint32_t parent = frameW[depth - 1];
          if (low[w] < low[parent]) low[parent] = low[w];
        }
      }
    }
  }
  free(order1);
  free(order2);
  free(h2);
  free(idx);
  free(low);
  free(onStack);
  free(wstack);
  free(frameW);
  free(frameS);
  free(key);
}

/* walk  IterativeAssembler.cpp:149-501.  The contig grows inside buf[lo,hi). */
typedef struct {
  uint8_t* buf;
  uint32_t lo, hi, cap;
} grow_t;

static int walk(const mb200_asm_opts* opt, uint32_t seed, wordtab_t* t, contig_t* c)
{
  const uint32_t k = t->k, W = t->W;
  static const char alphabet[4] = {'A', 'C', 'G', 'T'};

  grow_t g;
  g.cap = 2 * (t->nWords + k) + 16;
  g.buf = (uint8_t*)malloc(g.cap);
  g.lo  = g.cap / 2 - k / 2;
  g.hi  = g.lo + k;
  memcpy(g.buf + g.lo, wt_word(t, seed), k);

  bs_copy(c->support, t->reads + (size_t)seed * W, W);
  bs_zero(c->reject, W);
  t->unused[seed] = 0;
  c->consBegin = c->consEnd = 0;

  int isRepeatFound = 0;
  if (t->repeat[seed]) {
    c->consBegin  = 0;
    c->consEnd    = (int32_t)k;
    isRepeatFound = 1;
    goto done;
  }

  t->gen++;
  t->inContig[seed] = t->gen;

  uint8_t* key = (uint8_t*)malloc(k);
  bw_t*    maxWordReads       = (bw_t*)malloc(sizeof(bw_t) * W * 6);
  bw_t*    maxContigWordReads = maxWordReads + W;
  bw_t*    sup2Remove         = maxWordReads + 2 * W;
  bw_t*    rej2Add            = maxWordReads + 3 * W;
  bw_t*    contigWordReads    = maxWordReads + 4 * W;
  bw_t*    sharedReads        = maxWordReads + 5 * W;

  /* rejecting reads for the seed from its sibling words (same k-1 prefix)  :185-210 */
  memcpy(key, g.buf + g.lo, k - 1);
  for (int s = 0; s < 4; ++s) {
    if ((uint8_t)alphabet[s] == g.buf[g.lo + k - 1]) continue;
    key[k - 1]       = (uint8_t)alphabet[s];
    const int32_t id = wt_find(t, key);
    if (id < 0) continue;
    const bw_t* rd = t->reads + (size_t)id * W;
    for (uint32_t i = 0; i < W; ++i) c->reject[i] |= rd[i];
  }

  for (unsigned mode = 0; mode < 2; ++mode) {
    const int isEnd               = (mode == 0);
    uint32_t  conservativeEndOffset = 0;
    for (;;) {
      /* previousWord / trunk of the current contig end */
      const uint8_t* prevWord = isEnd ? g.buf + g.hi - k : g.buf + g.lo;
      const uint8_t* trunk    = isEnd ? g.buf + g.hi - (k - 1) : g.buf + g.lo;
      const uint8_t  tmpSymbol = isEnd ? prevWord[0] : prevWord[k - 1];

      uint32_t maxBaseCount = 0, maxContigWordReadCount = 0;
      uint8_t  maxBase = (uint8_t)alphabet[0];
      int32_t  maxWord = -1;
      bs_zero(maxWordReads, W);
      bs_zero(maxContigWordReads, W);
      bs_zero(sup2Remove, W);
      bs_zero(rej2Add, W);

      for (int s = 0; s < 4; ++s) {
        if (isEnd) {
          memcpy(key, trunk, k - 1);
          key[k - 1] = (uint8_t)alphabet[s];
        } else {
          key[0] = (uint8_t)alphabet[s];
          memcpy(key + 1, trunk, k - 1);
        }
        const int32_t id = wt_find(t, key);
        if (id < 0) continue;
        const uint32_t currWordCount = t->count[id];
        const bw_t*    currWordReads = t->reads + (size_t)id * W;
        for (uint32_t i = 0; i < W; ++i) {
          contigWordReads[i] = c->support[i] & currWordReads[i];
          sharedReads[i]     = maxContigWordReads[i] & currWordReads[i];
        }
        if (bs_empty(contigWordReads, W)) continue;
        const uint32_t cnt = bs_count(contigWordReads, W);
        if (cnt > maxContigWordReadCount) {
          for (uint32_t i = 0; i < W; ++i) {
            sup2Remove[i] |= maxContigWordReads[i] & ~sharedReads[i];
            rej2Add[i] |= maxWordReads[i] & ~sharedReads[i];
          }
          bs_copy(maxWordReads, currWordReads, W);
          maxContigWordReadCount = cnt;
          bs_copy(maxContigWordReads, contigWordReads, W);
          maxBaseCount = currWordCount;
          maxBase      = (uint8_t)alphabet[s];
          maxWord      = id;
        } else {
          for (uint32_t i = 0; i < W; ++i) {
            sup2Remove[i] |= contigWordReads[i] & ~sharedReads[i];
            rej2Add[i] |= currWordReads[i] & ~sharedReads[i];
          }
        }
      }

      if (maxBaseCount < opt->minCoverage) break;
      if (t->inContig[maxWord] == t->gen) {
        isRepeatFound = 1;
        break;
      }

      /* walk backwards one step at the branching point (:375-427); previousWordReads is always the
       * empty set at this comparison, so the branch is taken whenever maxWordReads is non-empty.
       * NB: uses the trunk of the contig BEFORE the new base is added. */
      if (!bs_empty(maxWordReads, W)) {
        for (int s = 0; s < 4; ++s) {
          if ((uint8_t)alphabet[s] == tmpSymbol) continue;
          if (isEnd) { /* addBase(trunk, symbol, !isEnd) : prepend */
            key[0] = (uint8_t)alphabet[s];
            memcpy(key + 1, trunk, k - 1);
          } else {
            memcpy(key, trunk, k - 1);
            key[k - 1] = (uint8_t)alphabet[s];
          }
          if (memcmp(key, wt_word(t, (uint32_t)maxWord), k) == 0) continue;
          const int32_t id = wt_find(t, key);
          if (id < 0) continue;
          const bw_t* back = t->reads + (size_t)id * W;
          for (uint32_t i = 0; i < W; ++i) {
            const bw_t toUpdate = back[i] & ~(maxContigWordReads[i] & back[i]);
            rej2Add[i] |= toUpdate;
            sup2Remove[i] |= toUpdate;
          }
        }
      }

      /* contig.seq = addBase(contig.seq, maxBase, isEnd) */
      if (isEnd) {
        g.buf[g.hi++] = maxBase;
      } else {
        g.buf[--g.lo] = maxBase;
      }
      if (g.hi >= g.cap || g.lo == 0) abort();

      if ((conservativeEndOffset != 0) || (maxBaseCount < opt->minConservativeCoverage)) conservativeEndOffset += 1;

      for (uint32_t i = 0; i < W; ++i) {
        c->reject[i] |= rej2Add[i];
        c->support[i] |= maxWordReads[i] & ~c->reject[i];
        c->support[i] &= ~sup2Remove[i];
      }
      t->unused[maxWord]   = 0;
      t->inContig[maxWord] = t->gen;
    }
    if (mode == 0)
      c->consEnd = (int32_t)conservativeEndOffset;
    else
      c->consBegin = (int32_t)conservativeEndOffset;
  }
  c->consEnd = (int32_t)(g.hi - g.lo) - c->consEnd;
  free(key);
  free(maxWordReads);

done:
  c->len = g.hi - g.lo;
  c->seq = (uint8_t*)malloc(c->len ? c->len : 1);
  memcpy(c->seq, g.buf + g.lo, c->len);
  free(g.buf);
  return isRepeatFound;
}

static void contig_free(contig_t* c)
{
  free(c->seq);
  free(c->support);
  free(c->reject);
}

/* buildContigs  IterativeAssembler.cpp:644-720 ; returns isAssemblySuccess */
static int build_contigs(const mb200_asm_opts* opt, const read_t* reads, uint32_t nReads, uint32_t k, uint32_t W,
                         contig_t* contigs, uint32_t* nContigs)
{
  for (uint32_t i = 0; i < *nContigs; ++i) contig_free(&contigs[i]);
  *nContigs = 0;
  int isAssemblySuccess = 1;

  uint32_t maxWords = 0;
  for (uint32_t r = 0; r < nReads; ++r)
    if (reads[r].len >= k) maxWords += reads[r].len - k + 1;

  wordtab_t t;
  wt_init(&t, k, W, maxWords);
  get_kmer_counts(opt, reads, nReads, &t);
  get_repeat_kmers(&t);

  uint32_t nUnused = 0;
  for (uint32_t w = 0; w < t.nWords; ++w) {
    if (t.count[w] >= opt->minCoverage) {
      t.unused[w] = 1;
    }
  }

  for (;;) {
    if (*nContigs >= 2 * opt->maxAssemblyCount) break;
    /* first word of max count in lexicographic order  :686-696 */
    int32_t  maxWord      = -1;
    uint32_t maxWordCount = 0;
    nUnused               = 0;
    for (uint32_t w = 0; w < t.nWords; ++w) {
      if (!t.unused[w]) continue;
      nUnused++;
      const uint32_t cnt = t.count[w];
      if (cnt > maxWordCount ||
          (cnt == maxWordCount && maxWord >= 0 && memcmp(wt_word(&t, w), wt_word(&t, (uint32_t)maxWord), k) < 0)) {
        maxWord      = (int32_t)w;
        maxWordCount = cnt;
      }
    }
    if (!nUnused) break;
    /* (maxWord stays -1 only if every unused word has count 0, impossible with minCoverage >= 1) */
    contig_t* c = &contigs[*nContigs];
    c->support  = (bw_t*)calloc(W, sizeof(bw_t));
    c->reject   = (bw_t*)calloc(W, sizeof(bw_t));
    const int isRepeatFound = walk(opt, (uint32_t)maxWord, &t, c);
    if (isRepeatFound) isAssemblySuccess = 0;
    (*nContigs)++;
  }
  wt_free(&t);
  return isAssemblySuccess;
}

/* selectContigs  IterativeAssembler.cpp:722-842 ; returns indices of the selected candidates in order.
 * "isPseudo" is evaluated as (read index >= normalReadCount): when the last word length fails, the reference
 * rebuilds the pseudo-read list AFTER the candidate contigs were built (:882-910), so their read sets can name
 * pseudo indices beyond the final list; the reference then indexes readInfo out of bounds (:765,:826) and in
 * practice reads the stale isPseudo==true of the erased element.  Index >= normalReadCount is that behaviour
 * made well-defined. */
static uint32_t select_contigs(const mb200_asm_opts* opt, const read_t* reads, uint32_t W, uint32_t normalReadCount,
                               const contig_t* cand, uint32_t nCand, uint32_t* selected)
{
  uint32_t finalCount = 0;
  bw_t*    usedReads       = (bw_t*)calloc(W, sizeof(bw_t));
  bw_t*    usedPseudoReads = (bw_t*)calloc(W, sizeof(bw_t));
  bw_t*    newSup          = (bw_t*)calloc(W, sizeof(bw_t));
  uint8_t* alive           = (uint8_t*)malloc(nCand ? nCand : 1);
  memset(alive, 1, nCand);
  uint32_t nAlive = nCand;

  while (nAlive && finalCount < opt->maxAssemblyCount) {
    const uint32_t usedNormalReads   = bs_count(usedReads, W) - bs_count(usedPseudoReads, W);
    const uint32_t unusedNormalReads = normalReadCount - usedNormalReads;
    if (unusedNormalReads < opt->minUnusedReads) break;

    int32_t  sel = -1;
    uint32_t maxSupport = 0, maxLength = 0;
    for (uint32_t ci = 0; ci < nCand; ++ci) {
      if (!alive[ci]) continue;
      uint32_t newNormalSupport = 0;
      for (uint32_t i = 0; i < W; ++i) newSup[i] = cand[ci].support[i] & ~usedReads[i];
      for (uint32_t i = 0; i < W; ++i) {
        bw_t v = newSup[i];
        while (v) {
          const uint32_t rd = i * 64 + (uint32_t)__builtin_ctzll(v);
          v &= v - 1;
          if (rd < normalReadCount) newNormalSupport++; /* see note on stale pseudo indices below */
        }
      }
      if (newNormalSupport < opt->minSupportReads) {
        /* contigs2Remove: erased after this round (also when the round selects nothing and breaks: the
         * erase only happens after the "maxSupport == 0 -> break", so a break leaves them, harmlessly) */
        alive[ci] = 2;
        continue;
      }
      const uint32_t currNewSupport = bs_count(newSup, W);
      const uint32_t currLen        = cand[ci].len;
      if (currNewSupport > maxSupport || (currNewSupport == maxSupport && currLen > maxLength)) {
        sel        = (int32_t)ci;
        maxSupport = currNewSupport;
        maxLength  = currLen;
      }
    }
    if (maxSupport == 0) break;
    selected[finalCount] = (uint32_t)sel;
    alive[sel]           = 2;
    for (uint32_t ci = 0; ci < nCand; ++ci)
      if (alive[ci] == 2) {
        alive[ci] = 0;
        nAlive--;
      }
    for (uint32_t i = 0; i < W; ++i) {
      bw_t v = cand[sel].support[i];
      usedReads[i] |= v;
      while (v) {
        const uint32_t rd = i * 64 + (uint32_t)__builtin_ctzll(v);
        v &= v - 1;
        if (rd >= normalReadCount) bs_set(usedPseudoReads, rd);
      }
    }
    finalCount++;
  }
  free(usedReads);
  free(usedPseudoReads);
  free(newSup);
  free(alive);
  return finalCount;
}

/* ------------------------------------------------------------------------------------------ */
int orc_assemble_batch(const mb200_asm_opts* opts, uint32_t nLoci, const uint64_t* readOff,
                       const mb200_seqview* readViews, const uint8_t* pool, mb200_asm_locus* loci,
                       mb200_contig* contigs, uint64_t contigCap, mb200_seqview* pseudoReads, uint64_t pseudoCap,
                       uint8_t* contigPool, uint64_t contigPoolCap, uint32_t* setPool, uint64_t setPoolCap)
{
  if (opts->minCoverage < 1 || opts->wordStepSize < 1 || opts->minWordLength < 2) return MB200_ERR_BAD_ARG;
  uint64_t nContigOut = 0, nPseudoOut = 0, poolUsed = 0, setUsed = 0;
  const uint32_t maxIter = 2 * opts->maxAssemblyCount;

  for (uint32_t li = 0; li < nLoci; ++li) {
    const uint32_t normalReadCount = (uint32_t)(readOff[li + 1] - readOff[li]);
    const uint32_t maxReads        = normalReadCount + maxIter;
    const uint32_t W               = (maxReads + 63) / 64;
    read_t*        reads           = (read_t*)calloc(maxReads, sizeof(read_t));
    uint32_t       nReads          = normalReadCount;
    for (uint32_t r = 0; r < normalReadCount; ++r) {
      const mb200_seqview* v = &readViews[readOff[li] + r];
      reads[r].seq           = (uint8_t*)(pool + v->off);
      reads[r].len           = v->len;
      reads[r].isPseudo      = 0;
    }
    contig_t* iter  = (contig_t*)calloc(maxIter ? maxIter : 1, sizeof(contig_t));
    uint32_t  nIter = 0;
    uint32_t  lastK = 0;

    for (uint32_t k = opts->minWordLength; k <= opts->maxWordLength; k += opts->wordStepSize) {
      lastK = k;
      const int ok = build_contigs(opts, reads, nReads, k, W, iter, &nIter);
      if (ok) break;
      /* drop the previous iteration's pseudo reads, add this iteration's long contigs  :882-910 */
      for (uint32_t r = normalReadCount; r < nReads; ++r) free(reads[r].seq);
      nReads = normalReadCount;
      for (uint32_t ci = 0; ci < nIter; ++ci) {
        if (iter[ci].len > k + opts->wordStepSize) {
          reads[nReads].seq = (uint8_t*)malloc(iter[ci].len);
          memcpy(reads[nReads].seq, iter[ci].seq, iter[ci].len);
          reads[nReads].len      = iter[ci].len;
          reads[nReads].isPseudo = 1;
          nReads++;
        }
      }
    }

    uint32_t* selected = (uint32_t*)malloc(sizeof(uint32_t) * (opts->maxAssemblyCount ? opts->maxAssemblyCount : 1));
    const uint32_t nSel = select_contigs(opts, reads, W, normalReadCount, iter, nIter, selected);

    mb200_asm_locus* L = &loci[li];
    memset(L, 0, sizeof(*L));
    L->firstContig    = (uint32_t)nContigOut;
    L->nContigs       = nSel;
    L->firstPseudo    = (uint32_t)nPseudoOut;
    L->nPseudo        = nReads - normalReadCount;
    L->setWords       = (normalReadCount + 2 * opts->maxAssemblyCount + 31) / 32;
    L->lastWordLength = lastK;

    for (uint32_t s = 0; s < nSel; ++s) {
      const contig_t* c = &iter[selected[s]];
      if (nContigOut < contigCap) {
        mb200_contig* o = &contigs[nContigOut];
        memset(o, 0, sizeof(*o));
        o->seqOff = poolUsed;
        o->seqLen = c->len;
        if (poolUsed + c->len <= contigPoolCap) memcpy(contigPool + poolUsed, c->seq, c->len);
        poolUsed += c->len;
        o->conservativeBegin = c->consBegin;
        o->conservativeEnd   = c->consEnd;
        for (int which = 0; which < 2; ++which) {
          const bw_t* src = which ? c->reject : c->support;
          if (which)
            o->rejectOff = setUsed;
          else
            o->supportOff = setUsed;
          if (setUsed + L->setWords <= setPoolCap) {
            for (uint32_t w = 0; w < L->setWords; ++w) {
              const bw_t v        = (w / 2 < W) ? src[w / 2] : 0;
              setPool[setUsed + w] = (uint32_t)((w & 1) ? (v >> 32) : v);
            }
          }
          setUsed += L->setWords;
        }
      }
      nContigOut++;
    }
    for (uint32_t r = normalReadCount; r < nReads; ++r) {
      if (nPseudoOut < pseudoCap) {
        mb200_seqview* v = &pseudoReads[nPseudoOut];
        v->off           = poolUsed;
        v->len           = reads[r].len;
        v->reserved      = 0;
        if (poolUsed + reads[r].len <= contigPoolCap) memcpy(contigPool + poolUsed, reads[r].seq, reads[r].len);
        poolUsed += reads[r].len;
      }
      nPseudoOut++;
      free(reads[r].seq);
    }
    for (uint32_t ci = 0; ci < nIter; ++ci) contig_free(&iter[ci]);
    free(iter);
    free(selected);
    free(reads);
  }
  const int ok = (nContigOut <= contigCap) && (nPseudoOut <= pseudoCap) && (poolUsed <= contigPoolCap) &&
                 (setUsed <= setPoolCap);
  return ok ? MB200_OK : MB200_ERR_POOL_OVERFLOW;
}

const char* orc_version(void) { return "manta_b200 oracle (C restatement of Illumina/manta v1.6.0 hot path)"; }
