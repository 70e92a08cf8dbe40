/* prune_probe.c -- analysis only: how much of the large-indel DP matrix can hold a cell of
 * an optimal path?  A cell (q,c) is "live" if best_state_score(q,c) + match*(Q-q) >= OPT (an admissible bound: every
 * remaining query base can gain at most `match`).  Recurrences as in oracle/orc_align.c (GlobalLargeIndelAlignerImpl.hpp:35-225),
 * scores only.  Reports live cells and live tiles for a few tile shapes.
 *   gcc -O2 -shared -fPIC tools/prune_probe.c -o /tmp/libprune.so */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#define BADVAL (-10000)
static int max2(int a, int b) { return a > b ? a : b; }
typedef struct { int m, d, i, jd, ji; } sv_t;

/* out[0]=OPT, out[1]=live cells, out[2]=cells, out[3+2k]=live tiles, out[4+2k]=tiles for tile rows {32,64,128,256} x 32 cols */
void prune_probe(const uint8_t* q, uint32_t Q, const uint8_t* ref, uint32_t R, int match, int mismatch, int open, int extend,
                 int offEdge, int L, int slack, int64_t* out)
{
  int* best = (int*)malloc(sizeof(int) * (size_t)(Q + 1) * (R + 1));
#define B(qi, ri) best[(size_t)(qi) * (R + 1) + (ri)]
  sv_t* a = (sv_t*)malloc(sizeof(sv_t) * (Q + 1));
  sv_t* b = (sv_t*)malloc(sizeof(sv_t) * (Q + 1));
  sv_t *cur = a, *prev = b;
  for (uint32_t qi = 0; qi <= Q; ++qi) {
    cur[qi].m = (int)qi * offEdge; cur[qi].d = cur[qi].i = cur[qi].jd = cur[qi].ji = BADVAL;
    B(qi, 0) = cur[qi].m;
  }
  int opt = -1000000;
  for (uint32_t ri = 0; ri < R; ++ri) {
    sv_t* t = cur; cur = prev; prev = t;
    cur[0].m = 0; cur[0].d = cur[0].i = cur[0].jd = cur[0].ji = BADVAL;
    B(0, ri + 1) = 0;
    for (uint32_t qi = 0; qi < Q; ++qi) {
      sv_t* h = &cur[qi + 1];
      const sv_t* s = &prev[qi];
      h->m = max2(max2(max2(s->m, s->d), max2(s->i, s->jd)), s->ji) + ((q[qi] == ref[ri]) ? match : mismatch);
      s = &prev[qi + 1];
      h->d = max2(max2(s->m + open, s->d), max2(s->i, s->ji)) + extend;
      h->jd = max2(max2(s->m + L, s->i + L - open), max2(s->jd, s->ji + L));
      s = &cur[qi];
      h->i = max2(s->m + open, s->i) + extend;
      h->ji = max2(s->m + L, s->ji);
      if (qi == 0) h->d = h->i = h->jd = h->ji = BADVAL;
      B(qi + 1, ri + 1) = max2(max2(max2(h->m, h->d), max2(h->i, h->jd)), h->ji);
    }
    if (cur[Q].m > opt) opt = cur[Q].m;
  }
  for (uint32_t qi = 0; qi <= Q; ++qi) opt = max2(opt, cur[qi].m + (int)(Q - qi) * offEdge);
  const int lb = opt - slack;
  int64_t live = 0;
  static const int TR[4] = {32, 64, 128, 256};
  int64_t liveT[4] = {0, 0, 0, 0}, totT[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) {
    const uint32_t nr = (Q + TR[k] - 1) / TR[k], nc = (R + 31) / 32;
    uint8_t* flag = (uint8_t*)calloc((size_t)nr * nc, 1);
    for (uint32_t qi = 1; qi <= Q; ++qi)
      for (uint32_t ri = 1; ri <= R; ++ri)
        if (B(qi, ri) + match * (int)(Q - qi) >= lb) {
          flag[(size_t)((qi - 1) / TR[k]) * nc + (ri - 1) / 32] = 1;
          if (k == 0) live++;
        }
    for (size_t i = 0; i < (size_t)nr * nc; ++i) liveT[k] += flag[i];
    totT[k] = (int64_t)nr * nc;
    free(flag);
  }
  out[0] = opt; out[1] = live; out[2] = (int64_t)Q * R;
  for (int k = 0; k < 4; ++k) { out[3 + 2 * k] = liveT[k] * TR[k] * 32; out[4 + 2 * k] = totT[k] * TR[k] * 32; }
  free(best); free(a); free(b);
}
