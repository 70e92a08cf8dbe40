/*
 * orc_candidates.c -- TEST INFRASTRUCTURE ONLY (see manta_oracle.h).
 *
 * Multi-threaded CPU driver of the per-candidate composition (assemble, then align every contig) on top of the C
 * restatement; used as the CPU baseline where the reference-built library (oracle/_ref) is unavailable, and as the
 * checksum provider for full-size GPU runs.  Loci are handed out dynamically to `threads` pthreads.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "manta_oracle.h"

typedef struct {
  const mb200_candidate_opts* opts;
  uint32_t                    nLoci;
  const uint64_t*             readOff;
  const mb200_seqview*        views;
  const uint8_t*              isJump;
  const mb200_seqview *       ref1, *ref2;
  const uint8_t*              pool;
  uint32_t                    next;
  uint64_t                    nContigs, cells;
  int64_t                     scoreSum;
  pthread_mutex_t             mu;
} job_t;

static void* worker(void* arg)
{
  job_t*         J = (job_t*)arg;
  const uint32_t A = J->opts->asmOpts.maxAssemblyCount;
  uint64_t       myContigs = 0, myCells = 0;
  int64_t        mySum = 0;
  size_t         poolCap = 1 << 16;
  uint8_t*       cpool = (uint8_t*)malloc(poolCap);
  size_t         setCap = 4096;
  uint32_t*      spool = (uint32_t*)malloc(sizeof(uint32_t) * setCap);
  mb200_contig*  contigs = (mb200_contig*)malloc(sizeof(mb200_contig) * A);
  mb200_seqview* pseudo  = (mb200_seqview*)malloc(sizeof(mb200_seqview) * 2 * A);
  uint32_t       cigar[1];
  for (;;) {
    const uint32_t li = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
    if (li >= J->nLoci) break;
    const uint64_t  off2[2] = {0, J->readOff[li + 1] - J->readOff[li]};
    mb200_asm_locus L;
    int             st;
    for (;;) {
      st = orc_assemble_batch(&J->opts->asmOpts, 1, off2, J->views + J->readOff[li], J->pool, &L, contigs, A, pseudo,
                              2 * A, cpool, poolCap, spool, setCap);
      if (st != MB200_ERR_POOL_OVERFLOW) break;
      poolCap *= 4;
      setCap *= 2;
      cpool = (uint8_t*)realloc(cpool, poolCap);
      spool = (uint32_t*)realloc(spool, sizeof(uint32_t) * setCap);
    }
    if (st != MB200_OK) continue;
    for (uint32_t k = 0; k < L.nContigs; ++k) {
      /* one private pool per alignment: [contig | ref1 | ref2] is avoided by pointing views at the two pools through
       * a common base: the aligners only need pool + off, so use the contig pool as base and pointer differences */
      uint64_t      used = 0;
      mb200_seqview q, r1, r2;
      q.off = (uint64_t)((cpool + contigs[k].seqOff) - J->pool); /* wraps modulo 2^64: pool + off is still exact */
      q.len = contigs[k].seqLen;
      q.reserved = 0;
      r1 = J->ref1[li];
      r2 = J->ref2[li];
      if (J->isJump[li]) {
        mb200_jump_result res;
        orc_jump_align_batch(&J->opts->jumpScores, J->opts->jumpScore, 1, J->pool, &q, &r1, &r2, &res, cigar, 0, &used);
        mySum += res.score;
        myCells += (uint64_t)q.len * (r1.len + r2.len);
      } else {
        mb200_aln_result res;
        orc_indel_align_batch(&J->opts->indelScores, J->opts->largeIndelScore, 1, J->pool, &q, &r1, &res, cigar, 0, &used);
        mySum += res.score;
        myCells += (uint64_t)q.len * r1.len;
      }
      myContigs++;
    }
  }
  free(cpool);
  free(spool);
  free(contigs);
  free(pseudo);
  pthread_mutex_lock(&J->mu);
  J->nContigs += myContigs;
  J->cells += myCells;
  J->scoreSum += mySum;
  pthread_mutex_unlock(&J->mu);
  return NULL;
}

int orc_candidates_mt(const mb200_candidate_opts* opts, uint32_t nLoci, const uint64_t* readOff,
                      const mb200_seqview* readViews, const uint8_t* isJump, const mb200_seqview* ref1,
                      const mb200_seqview* ref2, const uint8_t* pool, int threads, uint64_t* nContigsOut,
                      uint64_t* cellsOut, int64_t* scoreSumOut)
{
  job_t J;
  memset(&J, 0, sizeof(J));
  J.opts = opts; J.nLoci = nLoci; J.readOff = readOff; J.views = readViews; J.isJump = isJump;
  J.ref1 = ref1; J.ref2 = ref2; J.pool = pool;
  pthread_mutex_init(&J.mu, NULL);
  if (threads < 1) threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, worker, &J);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&J.mu);
  *nContigsOut = J.nContigs;
  *cellsOut    = J.cells;
  *scoreSumOut = J.scoreSum;
  return MB200_OK;
}
