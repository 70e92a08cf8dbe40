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
  const mb200_ref_cuts*       cuts;
  int                         doSelect;
  uint32_t                    next;
  uint64_t                    nContigs, cells;
  int64_t                     scoreSum, qcSum;
  pthread_mutex_t             mu;
} job_t;

static void* worker(void* arg)
{
  job_t*             J = (job_t*)arg;
  const uint32_t     A = J->opts->asmOpts.maxAssemblyCount;
  uint64_t           myContigs = 0, myCells = 0;
  int64_t            mySum = 0, myQc = 0;
  size_t             poolCap = 1 << 16;
  uint8_t*           cpool = (uint8_t*)malloc(poolCap);
  size_t             setCap = 4096;
  uint32_t*          spool = (uint32_t*)malloc(sizeof(uint32_t) * setCap);
  mb200_contig*      contigs = (mb200_contig*)malloc(sizeof(mb200_contig) * A);
  mb200_seqview*     pseudo  = (mb200_seqview*)malloc(sizeof(mb200_seqview) * 2 * A);
  mb200_aln_result*  aln  = (mb200_aln_result*)calloc(A, sizeof(mb200_aln_result));
  mb200_jump_result* jmp  = (mb200_jump_result*)calloc(A, sizeof(mb200_jump_result));
  mb200_contig_qc*   qc   = (mb200_contig_qc*)calloc(A, sizeof(mb200_contig_qc));
  mb200_seqview*     cseq = (mb200_seqview*)calloc(A, sizeof(mb200_seqview));
  uint32_t*          supp = (uint32_t*)calloc(A, sizeof(uint32_t));
  size_t             cigCap = J->doSelect ? (1 << 14) : 0;
  uint32_t*          cigar  = (uint32_t*)malloc(sizeof(uint32_t) * (cigCap ? cigCap : 1));
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
    uint64_t cigUsed = 0;
    for (uint32_t k = 0; k < L.nContigs; ++k) {
      /* the aligners only need pool + off: address the contig (which lives in this worker's private pool) through a
       * pointer difference from the shared pool -- it wraps modulo 2^64 and pool + off is still exact */
      mb200_seqview q, r1, r2;
      q.off = (uint64_t)((cpool + contigs[k].seqOff) - J->pool);
      q.len = contigs[k].seqLen;
      q.reserved = 0;
      cseq[k] = q;
      r1 = J->ref1[li];
      r2 = J->ref2[li];
      supp[k] = 0;
      for (uint32_t w = 0; w < L.setWords; ++w) supp[k] += (uint32_t)__builtin_popcount(spool[contigs[k].supportOff + w]);
      mb200_trim_result tr = {0, 0};
      if (J->cuts && !J->isJump[li]) {
        orc_ref_trim_batch(1, J->pool, &q, &r1, &J->cuts[li], &tr);
        myQc += ORC_QC_TERM_TRIM(tr);
        r1.off += (uint64_t)tr.adjLeadingCut;
        r1.len -= (uint32_t)(tr.adjLeadingCut + tr.adjTrailingCut);
      }
      for (;;) {
        uint64_t used = 0;
        int      rc;
        if (J->isJump[li]) {
          rc = orc_jump_align_batch(&J->opts->jumpScores, J->opts->jumpScore, 1, J->pool, &q, &r1, &r2, &jmp[k], cigar + cigUsed,
                                    cigCap ? cigCap - cigUsed : 0, &used);
          jmp[k].cigarOff1 += cigUsed;
          jmp[k].cigarOff2 += cigUsed;
        } else {
          rc = orc_indel_align_batch(&J->opts->indelScores, J->opts->largeIndelScore, 1, J->pool, &q, &r1, &aln[k],
                                     cigar + cigUsed, cigCap ? cigCap - cigUsed : 0, &used);
          aln[k].cigarOff += cigUsed;
          aln[k].beginPos += tr.adjLeadingCut; /* :2038 */
        }
        if (cigCap && rc == MB200_ERR_POOL_OVERFLOW) {
          cigCap = 2 * (cigCap + used);
          cigar  = (uint32_t*)realloc(cigar, sizeof(uint32_t) * cigCap);
          continue;
        }
        cigUsed += cigCap ? used : 0;
        break;
      }
      if (J->isJump[li]) {
        mySum += jmp[k].score;
        myCells += (uint64_t)q.len * (r1.len + r2.len);
      } else {
        mySum += aln[k].score;
        myCells += (uint64_t)q.len * r1.len;
      }
      myContigs++;
    }
    if (J->doSelect) {
      const uint32_t zero = 0;
      int32_t        best = -1;
      /* ref1 view of the QC is the UNtrimmed window (beginPos was shifted back above) */
      orc_select_contigs_batch(&J->opts->qcOpts, 1, &J->isJump[li], &zero, &L.nContigs, L.nContigs, aln, jmp, cigar, J->pool,
                               cseq, &J->ref1[li], supp, qc, &best);
      for (uint32_t k = 0; k < L.nContigs; ++k) myQc += ORC_QC_TERM_CONTIG(qc[k]);
      myQc += ORC_QC_TERM_LOCUS(best);
    }
  }
  free(cpool); free(spool); free(contigs); free(pseudo); free(aln); free(jmp); free(qc); free(cseq); free(supp); free(cigar);
  pthread_mutex_lock(&J->mu);
  J->nContigs += myContigs;
  J->cells += myCells;
  J->scoreSum += mySum;
  J->qcSum += myQc;
  pthread_mutex_unlock(&J->mu);
  return NULL;
}

int orc_candidates_mt(const mb200_candidate_opts* opts, uint32_t nLoci, const uint64_t* readOff,
                      const mb200_seqview* readViews, const uint8_t* isJump, const mb200_seqview* ref1,
                      const mb200_seqview* ref2, const uint8_t* pool, const mb200_ref_cuts* cuts, int doSelect,
                      int threads, uint64_t* nContigsOut, uint64_t* cellsOut, int64_t* scoreSumOut, int64_t* qcSumOut)
{
  job_t J;
  memset(&J, 0, sizeof(J));
  J.opts = opts; J.nLoci = nLoci; J.readOff = readOff; J.views = readViews; J.isJump = isJump;
  J.ref1 = ref1; J.ref2 = ref2; J.pool = pool; J.cuts = cuts; J.doSelect = doSelect;
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
  if (qcSumOut) *qcSumOut = J.qcSum;
  return MB200_OK;
}
