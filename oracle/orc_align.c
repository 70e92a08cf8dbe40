/*
 * orc_align.c -- TEST INFRASTRUCTURE ONLY (see manta_oracle.h).
 *
 * CPU restatement of the reference's affine-gap global aligners, full (Q+1)x(R+1) matrix,
 * int32 scores, finite sentinel badVal = -10000, first-maximum-wins back-pointers:
 *
 *   GlobalAligner<int>::align            lib/alignment/GlobalAlignerImpl.hpp:29-181
 *   GlobalLargeIndelAligner<int>::align  lib/alignment/GlobalLargeIndelAlignerImpl.hpp:35-225
 *   single-reference back-trace          lib/alignment/SingleRefAlignerSharedImpl.hpp:75-168
 *   GlobalJumpAligner<int>::align        lib/alignment/GlobalJumpAlignerImpl.hpp:33-333
 *   jump back-trace                      lib/alignment/JumpAlignerBaseImpl.hpp:86-242
 *   max3 / max4 / max5                   lib/alignment/AlignerBase.hpp:46-59, JumpAlignerBase.hpp:93-111,
 *                                        GlobalLargeIndelAligner.hpp:124-151
 *   updateBacktrace                      lib/alignment/AlignerUtil.hpp:53-67
 *   apath_add_seqmatch                   lib/blt_util/align_path_impl.hpp:33-75
 */
#include <stdlib.h>
#include <string.h>

#include "manta_oracle.h"

#define BADVAL (-10000)

/* AlignState::index_t  lib/alignment/Alignment.hpp:47-56 */
enum { ST_MATCH = 0, ST_DELETE = 1, ST_INSERT = 2, ST_JUMP = 3, ST_JUMPINS = 4 };

/* ------------------------------------------------------------------------------------------ */
/* growable alignment path                                                                    */
typedef struct {
  uint32_t* seg; /* (length<<4)|type */
  uint32_t  n, cap;
} path_t;

static void path_push(path_t* p, uint32_t type, uint32_t length)
{
  if (p->n == p->cap) {
    p->cap = p->cap ? 2 * p->cap : 64;
    p->seg = (uint32_t*)realloc(p->seg, sizeof(uint32_t) * p->cap);
  }
  p->seg[p->n++] = MB200_SEG_MAKE(length, type);
}

/* apath_append (align_path.cpp): extend the last segment if it has the same type */
static void path_append(path_t* p, uint32_t type, uint32_t length)
{
  if (p->n && MB200_SEG_TYPE(p->seg[p->n - 1]) == type) {
    p->seg[p->n - 1] += (length << 4);
  } else {
    path_push(p, type, length);
  }
}

static void path_reverse(path_t* p)
{
  for (uint32_t i = 0; i < p->n / 2; ++i) {
    const uint32_t t    = p->seg[i];
    p->seg[i]           = p->seg[p->n - 1 - i];
    p->seg[p->n - 1 - i] = t;
  }
}

static int seg_is_read_len(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_INSERT || t == MB200_SEG_SOFT_CLIP || t == MB200_SEG_SEQ_MATCH ||
         t == MB200_SEG_SEQ_MISMATCH;
}
static int seg_is_ref_len(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_DELETE || t == MB200_SEG_SKIP || t == MB200_SEG_SEQ_MATCH ||
         t == MB200_SEG_SEQ_MISMATCH;
}
static int seg_is_align_match(uint32_t t)
{
  return t == MB200_SEG_MATCH || t == MB200_SEG_SEQ_MATCH || t == MB200_SEG_SEQ_MISMATCH;
}

static uint32_t path_read_length(const path_t* p)
{
  uint32_t v = 0;
  for (uint32_t i = 0; i < p->n; ++i)
    if (seg_is_read_len(MB200_SEG_TYPE(p->seg[i]))) v += MB200_SEG_LEN(p->seg[i]);
  return v;
}
static uint32_t path_ref_length(const path_t* p)
{
  uint32_t v = 0;
  for (uint32_t i = 0; i < p->n; ++i)
    if (seg_is_ref_len(MB200_SEG_TYPE(p->seg[i]))) v += MB200_SEG_LEN(p->seg[i]);
  return v;
}

/* apath_add_seqmatch  align_path_impl.hpp:33-75 ; N on either side is a mismatch */
static void path_add_seqmatch(const uint8_t* q, const uint8_t* qEnd, const uint8_t* r, const uint8_t* rEnd,
                              path_t* p)
{
  path_t out = {0, 0, 0};
  for (uint32_t s = 0; s < p->n; ++s) {
    const uint32_t type = MB200_SEG_TYPE(p->seg[s]);
    const uint32_t len  = MB200_SEG_LEN(p->seg[s]);
    if (seg_is_align_match(type)) {
      for (uint32_t k = 0; k < len; ++k) {
        if (q >= qEnd || r >= rEnd) abort(); /* blt_exception in the reference; unreachable */
        int isMatch = (*q == *r);
        if (*q == 'N' || *r == 'N') isMatch = 0;
        path_append(&out, isMatch ? MB200_SEG_SEQ_MATCH : MB200_SEG_SEQ_MISMATCH, 1);
        ++q;
        ++r;
      }
    } else {
      path_push(&out, type, len);
      if (seg_is_read_len(type)) q += len;
      if (seg_is_ref_len(type)) r += len;
    }
  }
  free(p->seg);
  *p = out;
}

/* AlignerUtil::updatePath  AlignerUtil.hpp:32-38 */
typedef struct {
  uint32_t type, length;
} pseg_t;

static void update_path(path_t* path, pseg_t* ps, uint32_t atype)
{
  if (ps->type == atype) return;
  if (ps->type != MB200_SEG_NONE) path_push(path, ps->type, ps->length);
  ps->type   = atype;
  ps->length = 0;
}

/* BackTrace<int> + updateBacktrace  AlignerUtil.hpp:42-67 */
typedef struct {
  int32_t  max;
  uint32_t state, queryBegin, refBegin;
  int      isInit;
} btrace_t;

static void update_backtrace(int32_t thisMax, uint32_t refIndex, uint32_t queryIndex, btrace_t* bt, uint32_t state)
{
  if (!bt->isInit || thisMax > bt->max) {
    bt->max        = thisMax;
    bt->refBegin   = refIndex;
    bt->queryBegin = queryIndex;
    bt->isInit     = 1;
    bt->state      = state;
  }
}

/* first-maximum-wins scans; the returned index is the stored back-pointer */
static inline uint32_t max3(int32_t* m, int32_t v0, int32_t v1, int32_t v2)
{
  uint32_t p = 0;
  *m         = v0;
  if (v1 > v0) { *m = v1; p = 1; }
  if (v2 > *m) { *m = v2; p = 2; }
  return p;
}
static inline uint32_t max4(int32_t* m, int32_t v0, int32_t v1, int32_t v2, int32_t v3)
{
  uint32_t p = 0;
  *m         = v0;
  if (v1 > v0) { *m = v1; p = 1; }
  if (v2 > *m) { *m = v2; p = 2; }
  if (v3 > *m) { *m = v3; p = 3; }
  return p;
}
static inline uint32_t max5(int32_t* m, int32_t v0, int32_t v1, int32_t v2, int32_t v3, int32_t v4)
{
  uint32_t p = 0;
  *m         = v0;
  if (v1 > *m) { *m = v1; p = 1; }
  if (v2 > *m) { *m = v2; p = 2; }
  if (v3 > *m) { *m = v3; p = 3; }
  if (v4 > *m) { *m = v4; p = 4; }
  return p;
}

typedef struct {
  int32_t m, d, i, jd, ji; /* jd doubles as the jump state of the jump aligner */
} sv_t;

/* back-pointer cell: 3 bits per state */
#define PTR_PACK(pm, pd, pi, pjd, pji) ((uint16_t)((pm) | ((pd) << 3) | ((pi) << 6) | ((pjd) << 9) | ((pji) << 12)))
static inline uint32_t ptr_get(uint16_t cell, uint32_t state)
{
  return (cell >> (3 * state)) & 7u;
}

/* ------------------------------------------------------------------------------------------ */
/* GlobalAligner (nStates==3) and GlobalLargeIndelAligner (nStates==5)                         */
static void single_ref_align(const mb200_scores* sc, int isIndel, int32_t L, const uint8_t* q, uint32_t Q,
                             const uint8_t* ref, uint32_t R, mb200_aln_result* res, path_t* apath)
{
  sv_t*     s1  = (sv_t*)malloc(sizeof(sv_t) * (Q + 1));
  sv_t*     s2  = (sv_t*)malloc(sizeof(sv_t) * (Q + 1));
  uint16_t* ptr = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(Q + 1) * (R + 1));
  sv_t *    thisSV = s1, *prevSV = s2;
#define PTR(qi, ri) ptr[(size_t)(qi) * (R + 1) + (ri)]

  for (uint32_t qi = 0; qi <= Q; ++qi) {
    sv_t* v = &thisSV[qi];
    v->m    = (int32_t)qi * sc->offEdge;
    v->d    = BADVAL;
    uint32_t pi = ST_MATCH;
    if (!sc->isAllowEdgeInsertion) {
      v->i = BADVAL;
    } else {
      pi   = ST_INSERT;
      v->i = sc->open + (int32_t)qi * sc->extend;
    }
    v->jd = v->ji = BADVAL;
    PTR(qi, 0)    = PTR_PACK(0, 0, pi, 0, 0);
  }

  btrace_t bt;
  memset(&bt, 0, sizeof(bt));

  for (uint32_t ri = 0; ri < R; ++ri) {
    sv_t* t = thisSV;
    thisSV  = prevSV;
    prevSV  = t;
    {
      sv_t* v = &thisSV[0];
      v->m    = 0;
      v->d = v->i = v->jd = v->ji = BADVAL;
      PTR(0, ri + 1)               = 0;
    }
    for (uint32_t qi = 0; qi < Q; ++qi) {
      sv_t*    h = &thisSV[qi + 1];
      uint32_t pm, pd, pi, pjd = 0, pji = 0;
      {
        const sv_t* s = &prevSV[qi];
        pm            = isIndel ? max5(&h->m, s->m, s->d, s->i, s->jd, s->ji) : max3(&h->m, s->m, s->d, s->i);
        h->m += (q[qi] == ref[ri]) ? sc->match : sc->mismatch;
      }
      {
        const sv_t* s = &prevSV[qi + 1];
        pd = isIndel ? max5(&h->d, s->m + sc->open, s->d, s->i, BADVAL, s->ji) : max3(&h->d, s->m + sc->open, s->d, s->i);
        h->d += sc->extend;
        if (qi == 0) h->d = BADVAL;
      }
      {
        const sv_t* s = &thisSV[qi];
        pi = isIndel ? max5(&h->i, s->m + sc->open, BADVAL, s->i, BADVAL, BADVAL) : max3(&h->i, s->m + sc->open, BADVAL, s->i);
        h->i += sc->extend;
        if (qi == 0) h->i = BADVAL;
      }
      if (isIndel) {
        {
          const sv_t* s = &prevSV[qi + 1];
          pjd           = max5(&h->jd, s->m + L, BADVAL, s->i + L - sc->open, s->jd, s->ji + L);
          if (qi == 0) h->jd = BADVAL;
        }
        {
          const sv_t* s = &thisSV[qi];
          pji           = max5(&h->ji, s->m + L, BADVAL, BADVAL, BADVAL, s->ji);
          if (qi == 0) h->ji = BADVAL;
        }
      } else {
        h->jd = h->ji = BADVAL;
      }
      PTR(qi + 1, ri + 1) = PTR_PACK(pm, pd, pi, pjd, pji);
    }
    update_backtrace(thisSV[Q].m, ri + 1, Q, &bt, ST_MATCH);
  }

  if (sc->isAllowEdgeInsertion) update_backtrace(thisSV[Q].i, R, Q, &bt, ST_INSERT);

  {
    /* GlobalAligner scans q < Q (GlobalAlignerImpl.hpp:168), the large-indel aligner q <= Q (:211) */
    const uint32_t qEnd = isIndel ? Q + 1 : Q;
    for (uint32_t qi = 0; qi < qEnd; ++qi) {
      update_backtrace(thisSV[qi].m + (int32_t)(Q - qi) * sc->offEdge, R, qi, &bt, ST_MATCH);
    }
  }

  /* back-trace  SingleRefAlignerSharedImpl.hpp:75-168 */
  res->score    = bt.max;
  res->isJumped = 0;
  pseg_t ps     = {MB200_SEG_NONE, 0};
  if (bt.queryBegin < Q) {
    ps.type   = MB200_SEG_SOFT_CLIP;
    ps.length = Q - bt.queryBegin;
  }
  for (;;) {
    const uint32_t next = ptr_get(PTR(bt.queryBegin, bt.refBegin), bt.state);
    if (bt.state == ST_MATCH) {
      if (bt.queryBegin < 1 || bt.refBegin < 1) break;
      update_path(apath, &ps, MB200_SEG_MATCH);
      bt.queryBegin--;
      bt.refBegin--;
    } else if (bt.state == ST_DELETE || bt.state == ST_JUMP) {
      if (bt.refBegin < 1) break;
      update_path(apath, &ps, MB200_SEG_DELETE);
      bt.refBegin--;
    } else {
      if (bt.queryBegin < 1) break;
      update_path(apath, &ps, MB200_SEG_INSERT);
      bt.queryBegin--;
    }
    if (bt.state == ST_JUMP || bt.state == ST_JUMPINS) res->isJumped = 1;
    bt.state = next;
    ps.length++;
  }
  if (ps.type != MB200_SEG_NONE) path_push(apath, ps.type, ps.length);
  if (bt.queryBegin != 0) path_push(apath, MB200_SEG_SOFT_CLIP, bt.queryBegin);
  res->beginPos = (int32_t)bt.refBegin;
  path_reverse(apath);
  path_add_seqmatch(q, q + Q, ref + res->beginPos, ref + R, apath);

#undef PTR
  free(s1);
  free(s2);
  free(ptr);
}

/* ------------------------------------------------------------------------------------------ */
/* GlobalJumpAligner                                                                          */
static void jump_align(const mb200_scores* sc, int32_t jumpScore, const uint8_t* q, uint32_t Q, const uint8_t* ref1,
                       uint32_t R1, const uint8_t* ref2, uint32_t R2, mb200_jump_result* res, path_t* apath1,
                       path_t* apath2)
{
  sv_t*     s1   = (sv_t*)malloc(sizeof(sv_t) * (Q + 1));
  sv_t*     s2   = (sv_t*)malloc(sizeof(sv_t) * (Q + 1));
  uint16_t* ptr1 = (uint16_t*)calloc((size_t)(Q + 1) * (R1 + 1), sizeof(uint16_t));
  uint16_t* ptr2 = (uint16_t*)calloc((size_t)(Q + 1) * (R2 + 1), sizeof(uint16_t));
  sv_t *    thisSV = s1, *prevSV = s2;
#define PTR1(qi, ri) ptr1[(size_t)(qi) * (R1 + 1) + (ri)]
#define PTR2(qi, ri) ptr2[(size_t)(qi) * (R2 + 1) + (ri)]
  /* here state index 3 (JUMP) lives in field jd */

  for (uint32_t qi = 0; qi <= Q; ++qi) {
    sv_t* v = &thisSV[qi];
    v->m    = (int32_t)qi * sc->offEdge;
    v->d = v->i = v->jd = BADVAL;
    v->ji               = BADVAL;
  }

  btrace_t bt;
  memset(&bt, 0, sizeof(bt));

  for (uint32_t ri = 0; ri < R1; ++ri) {
    sv_t* t = thisSV;
    thisSV  = prevSV;
    prevSV  = t;
    {
      sv_t* v = &thisSV[0];
      v->m    = 0;
      v->d = v->i = v->jd = BADVAL;
    }
    for (uint32_t qi = 0; qi < Q; ++qi) {
      sv_t*    h = &thisSV[qi + 1];
      uint32_t pm, pd, pi, pj;
      {
        const sv_t* s = &prevSV[qi];
        pm            = max3(&h->m, s->m, s->d, s->i);
        h->m += (q[qi] == ref1[ri]) ? sc->match : sc->mismatch;
      }
      {
        const sv_t* s = &prevSV[qi + 1];
        pd            = max3(&h->d, s->m + sc->open, s->d, s->i);
        h->d += sc->extend;
        if (qi == 0) h->d = BADVAL;
      }
      {
        const sv_t* s = &thisSV[qi];
        pi            = max3(&h->i, s->m + sc->open, BADVAL, s->i);
        h->i += sc->extend;
        if (qi == 0) h->i = BADVAL;
      }
      {
        const sv_t* s = &prevSV[qi + 1];
        pj            = max4(&h->jd, h->m + jumpScore, BADVAL, h->i + jumpScore, s->jd);
      }
      PTR1(qi + 1, ri + 1) = PTR_PACK(pm, pd, pi, pj, 0);
    }
    update_backtrace(thisSV[Q].m, ri + 1, Q, &bt, ST_MATCH);
  }

  for (uint32_t qi = 0; qi < Q; ++qi) {
    update_backtrace(thisSV[qi].m + (int32_t)(Q - qi) * sc->offEdge, R1, qi, &bt, ST_MATCH);
  }

  for (uint32_t qi = 0; qi <= Q; ++qi) {
    sv_t* v = &thisSV[qi];
    v->m    = (int32_t)qi * sc->offEdge;
    v->d = v->i = BADVAL; /* jump preserved from the last ref1 column */
  }

  for (uint32_t ri = 0; ri < R2; ++ri) {
    sv_t* t = thisSV;
    thisSV  = prevSV;
    prevSV  = t;
    {
      sv_t* v = &thisSV[0];
      v->m    = 0;
      v->d = v->i = v->jd = BADVAL;
    }
    for (uint32_t qi = 0; qi < Q; ++qi) {
      sv_t*    h = &thisSV[qi + 1];
      uint32_t pm, pd, pi;
      {
        const sv_t* s = &prevSV[qi];
        pm            = max4(&h->m, s->m, s->d, s->i, s->jd);
        h->m += (q[qi] == ref2[ri]) ? sc->match : sc->mismatch;
      }
      {
        const sv_t* s = &prevSV[qi + 1];
        pd            = max3(&h->d, s->m + sc->open, s->d, s->i);
        h->d += sc->extend;
      }
      {
        const sv_t* s = &thisSV[qi];
        pi            = max4(&h->i, s->m + sc->open, BADVAL, s->i, s->jd);
        h->i += sc->extend;
      }
      h->jd                = prevSV[qi + 1].jd;
      PTR2(qi + 1, ri + 1) = PTR_PACK(pm, pd, pi, ST_JUMP, 0);
    }
    update_backtrace(thisSV[Q].m, R1 + ri + 1, Q, &bt, ST_MATCH);
  }

  for (uint32_t qi = 0; qi < Q; ++qi) {
    update_backtrace(thisSV[qi].m + (int32_t)(Q - qi) * sc->offEdge, R1 + R2, qi, &bt, ST_MATCH);
  }

  /* back-trace  JumpAlignerBaseImpl.hpp:101-242 */
  res->score          = bt.max;
  res->jumpInsertSize = 0;
  res->jumpRange      = 0;
  res->beginPos1      = 0;
  res->beginPos2      = 0;

  pseg_t ps = {MB200_SEG_NONE, 0};
  if (bt.queryBegin < Q) {
    ps.type   = MB200_SEG_SOFT_CLIP;
    ps.length = Q - bt.queryBegin;
  }
  int isRef2End = 0;
  while (bt.queryBegin > 0 && bt.refBegin > 0) {
    if (isRef2End) break;
    const int      isRef1    = (bt.refBegin <= R1);
    path_t*        apath     = isRef1 ? apath1 : apath2;
    const uint32_t refXBegin = bt.refBegin - (isRef1 ? 0 : R1);
    const uint32_t next =
        ptr_get(isRef1 ? PTR1(bt.queryBegin, refXBegin) : PTR2(bt.queryBegin, refXBegin), bt.state);

    if (bt.state == ST_MATCH) {
      if (!isRef1 && refXBegin == 1 && next == ST_MATCH) isRef2End = 1;
      update_path(apath, &ps, MB200_SEG_MATCH);
      bt.queryBegin--;
      bt.refBegin--;
    } else if (bt.state == ST_DELETE) {
      update_path(apath, &ps, MB200_SEG_DELETE);
      bt.refBegin--;
    } else if (bt.state == ST_INSERT) {
      update_path(apath, &ps, MB200_SEG_INSERT);
      bt.queryBegin--;
    } else { /* ST_JUMP (SPLICE never occurs in this aligner) */
      if (ps.type != MB200_SEG_NONE) {
        res->beginPos2 = (int32_t)(bt.refBegin - R1);
        if (ps.type == MB200_SEG_INSERT) {
          res->jumpInsertSize += ps.length;
          ps.type   = MB200_SEG_NONE;
          ps.length = 0;
        } else {
          update_path(apath2, &ps, MB200_SEG_NONE);
        }
      } else {
        if (next == ST_JUMP) bt.refBegin--;
      }
    }
    bt.state = next;
    ps.length++;
  }

  {
    const int isRef1 = (bt.refBegin < R1);
    path_t*   apath  = isRef1 ? apath1 : apath2;
    if (ps.type != MB200_SEG_NONE) path_push(apath, ps.type, ps.length);
    if (bt.queryBegin != 0) path_push(apath, MB200_SEG_SOFT_CLIP, bt.queryBegin);
    if (isRef1) {
      res->beginPos1 = (int32_t)bt.refBegin;
    } else {
      res->beginPos2 = (int32_t)(bt.refBegin - R1);
    }
  }
  path_reverse(apath1);
  path_reverse(apath2);

  if (apath1->n && apath2->n) {
    uint32_t r1i = (uint32_t)res->beginPos1 + path_ref_length(apath1);
    uint32_t r2i = (uint32_t)res->beginPos2;
    uint32_t qi  = path_read_length(apath1);
    uint32_t jumpInsertCount = res->jumpInsertSize;
    for (;;) {
      if (r1i == R1) break;
      if (jumpInsertCount > 0) {
        if (qi == Q) break;
        if (ref1[r1i] != q[qi]) break;
      } else {
        if (r2i == R2) break;
        if (ref1[r1i] != ref2[r2i]) break;
      }
      res->jumpRange++;
      r1i++;
      if (jumpInsertCount > 0) {
        jumpInsertCount--;
        qi++;
      } else {
        r2i++;
      }
    }
  }

  path_add_seqmatch(q, q + Q, ref1 + res->beginPos1, ref1 + R1, apath1);
  {
    const uint32_t queryOffset = path_read_length(apath1) + res->jumpInsertSize;
    path_add_seqmatch(q + queryOffset, q + Q, ref2 + res->beginPos2, ref2 + R2, apath2);
  }

#undef PTR1
#undef PTR2
  free(s1);
  free(s2);
  free(ptr1);
  free(ptr2);
}

/* ------------------------------------------------------------------------------------------ */
static int emit(const path_t* p, uint32_t* pool, uint64_t cap, uint64_t* used, uint64_t* off, uint32_t* len)
{
  *off = *used;
  *len = p->n;
  for (uint32_t i = 0; i < p->n; ++i) {
    if (*used < cap) pool[*used] = p->seg[i];
    (*used)++;
  }
  return *used <= cap;
}

int orc_jump_align_batch(const mb200_scores* scores, int32_t jumpScore, uint32_t n, const uint8_t* pool,
                         const mb200_seqview* query, const mb200_seqview* ref1, const mb200_seqview* ref2,
                         mb200_jump_result* results, uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed)
{
  uint64_t used = 0;
  int      ok   = 1;
  for (uint32_t i = 0; i < n; ++i) {
    if (query[i].len == 0) return MB200_ERR_EMPTY_QUERY;
    if (ref1[i].len == 0) return MB200_ERR_EMPTY_REF1;
    if (ref2[i].len == 0) return MB200_ERR_EMPTY_REF2;
  }
  for (uint32_t i = 0; i < n; ++i) {
    path_t             a1 = {0, 0, 0}, a2 = {0, 0, 0};
    mb200_jump_result* o = &results[i];
    memset(o, 0, sizeof(*o));
    jump_align(scores, jumpScore, pool + query[i].off, query[i].len, pool + ref1[i].off, ref1[i].len,
               pool + ref2[i].off, ref2[i].len, o, &a1, &a2);
    ok &= emit(&a1, cigarPool, cigarCap, &used, &o->cigarOff1, &o->cigarLen1);
    ok &= emit(&a2, cigarPool, cigarCap, &used, &o->cigarOff2, &o->cigarLen2);
    free(a1.seg);
    free(a2.seg);
  }
  *cigarUsed = used;
  return ok ? MB200_OK : MB200_ERR_POOL_OVERFLOW;
}

static int single_batch(const mb200_scores* scores, int isIndel, int32_t L, uint32_t n, const uint8_t* pool,
                        const mb200_seqview* query, const mb200_seqview* ref, mb200_aln_result* results,
                        uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed)
{
  uint64_t used = 0;
  int      ok   = 1;
  for (uint32_t i = 0; i < n; ++i) {
    if (query[i].len == 0) return MB200_ERR_EMPTY_QUERY;
    if (ref[i].len == 0) return MB200_ERR_EMPTY_REF;
  }
  for (uint32_t i = 0; i < n; ++i) {
    path_t            a = {0, 0, 0};
    mb200_aln_result* o = &results[i];
    memset(o, 0, sizeof(*o));
    single_ref_align(scores, isIndel, L, pool + query[i].off, query[i].len, pool + ref[i].off, ref[i].len, o, &a);
    ok &= emit(&a, cigarPool, cigarCap, &used, &o->cigarOff, &o->cigarLen);
    free(a.seg);
  }
  *cigarUsed = used;
  return ok ? MB200_OK : MB200_ERR_POOL_OVERFLOW;
}

int orc_indel_align_batch(const mb200_scores* scores, int32_t largeIndelScore, uint32_t n, const uint8_t* pool,
                          const mb200_seqview* query, const mb200_seqview* ref, mb200_aln_result* results,
                          uint32_t* cigarPool, uint64_t cigarCap, uint64_t* cigarUsed)
{
  return single_batch(scores, 1, largeIndelScore, n, pool, query, ref, results, cigarPool, cigarCap, cigarUsed);
}

int orc_global_align_batch(const mb200_scores* scores, uint32_t n, const uint8_t* pool, const mb200_seqview* query,
                           const mb200_seqview* ref, mb200_aln_result* results, uint32_t* cigarPool,
                           uint64_t cigarCap, uint64_t* cigarUsed)
{
  return single_batch(scores, 0, 0, n, pool, query, ref, results, cigarPool, cigarCap, cigarUsed);
}
