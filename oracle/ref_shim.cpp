// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" adapter around the UNMODIFIED Illumina/manta reference sources, compiled where
// they lie under /root/reference by oracle/build_ref.sh into oracle/_ref/libmanta_ref.so.
// It flattens the reference's result objects into the structs of include/manta_b200.h so that
// the reference, the C restatement (oracle/manta_oracle.c) and the CUDA library can be compared
// byte for byte.  Only tests/, bench.py's cpu_baseline / --impl reference leg and
// __graft_entry__.smoke() may load the resulting library.
//
// Reference entry points driven here:
//   GlobalJumpAligner<int>::align        lib/alignment/GlobalJumpAligner.hpp:45-53
//   GlobalLargeIndelAligner<int>::align  lib/alignment/GlobalLargeIndelAligner.hpp:48-54
//   GlobalAligner<int>::align            lib/alignment/GlobalAligner.hpp
//   splitReadAligner                     GSC/SplitReadAlignment.hpp:63-70
//   runIterativeAssembler                lib/assembly/IterativeAssembler.hpp:43-47
//   qscore_snp                           lib/blt_util/qscore_snp.hpp:36-62

#include <cstdint>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "alignment/GlobalAligner.hpp"
#include "alignment/GlobalJumpAligner.hpp"
#include "alignment/GlobalLargeIndelAligner.hpp"
#include "applications/GenerateSVCandidates/SplitReadAlignment.hpp"
#include "assembly/IterativeAssembler.hpp"
#include "blt_util/qscore_snp.hpp"
#include "common/Exceptions.hpp"

#include "manta_b200.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

// append a reference path as (length<<4)|type words; returns false on pool overflow
bool emitPath(const ALIGNPATH::path_t& apath, uint32_t* pool, uint64_t cap, uint64_t& used, uint64_t& off,
              uint32_t& len)
{
  off = used;
  len = static_cast<uint32_t>(apath.size());
  for (const ALIGNPATH::path_segment& ps : apath) {
    if (used < cap) pool[used] = MB200_SEG_MAKE(ps.length, static_cast<uint32_t>(ps.type));
    used++;
  }
  return used <= cap;
}

AlignmentScores<int> toScores(const mb200_scores* s)
{
  return AlignmentScores<int>(s->match, s->mismatch, s->open, s->extend, s->offEdge, s->isAllowEdgeInsertion != 0);
}

int classifyException(const std::exception& e)
{
  const std::string what(e.what());
  if (what.find("empty query") != std::string::npos) return MB200_ERR_EMPTY_QUERY;
  if (what.find("empty reference1") != std::string::npos) return MB200_ERR_EMPTY_REF1;
  if (what.find("empty reference2") != std::string::npos) return MB200_ERR_EMPTY_REF2;
  if (what.find("empty reference") != std::string::npos) return MB200_ERR_EMPTY_REF;
  return MB200_ERR_BAD_ARG;
}

}  // namespace

REF_API int ref_jump_align_batch(const mb200_scores* scores, int32_t jumpScore, uint32_t n, const uint8_t* pool,
                                 const mb200_seqview* query, const mb200_seqview* ref1,
                                 const mb200_seqview* ref2, mb200_jump_result* results, uint32_t* cigarPool,
                                 uint64_t cigarCap, uint64_t* cigarUsed)
{
  const GlobalJumpAligner<int> aligner(toScores(scores), jumpScore);
  JumpAlignmentResult<int>     res;
  uint64_t                     used(0);
  bool                         ok(true);
  for (uint32_t i(0); i < n; ++i) {
    const char* q(reinterpret_cast<const char*>(pool + query[i].off));
    const char* r1(reinterpret_cast<const char*>(pool + ref1[i].off));
    const char* r2(reinterpret_cast<const char*>(pool + ref2[i].off));
    try {
      aligner.align(q, q + query[i].len, r1, r1 + ref1[i].len, r2, r2 + ref2[i].len, res);
    } catch (const std::exception& e) {
      return classifyException(e);
    }
    mb200_jump_result& o(results[i]);
    std::memset(&o, 0, sizeof(o));
    o.score          = res.score;
    o.jumpInsertSize = res.jumpInsertSize;
    o.jumpRange      = res.jumpRange;
    o.beginPos1      = res.align1.beginPos;
    o.beginPos2      = res.align2.beginPos;
    ok &= emitPath(res.align1.apath, cigarPool, cigarCap, used, o.cigarOff1, o.cigarLen1);
    ok &= emitPath(res.align2.apath, cigarPool, cigarCap, used, o.cigarOff2, o.cigarLen2);
  }
  *cigarUsed = used;
  return ok ? MB200_OK : MB200_ERR_POOL_OVERFLOW;
}

template <typename Aligner>
static int singleRefBatch(const Aligner& aligner, uint32_t n, const uint8_t* pool, const mb200_seqview* query,
                          const mb200_seqview* ref, mb200_aln_result* results, uint32_t* cigarPool,
                          uint64_t cigarCap, uint64_t* cigarUsed)
{
  AlignmentResult<int> res;
  uint64_t             used(0);
  bool                 ok(true);
  for (uint32_t i(0); i < n; ++i) {
    const char* q(reinterpret_cast<const char*>(pool + query[i].off));
    const char* r(reinterpret_cast<const char*>(pool + ref[i].off));
    // GlobalLargeIndelAligner only asserts on empty input (GlobalLargeIndelAlignerImpl.hpp:50-51)
    if (query[i].len == 0) return MB200_ERR_EMPTY_QUERY;
    if (ref[i].len == 0) return MB200_ERR_EMPTY_REF;
    try {
      aligner.align(q, q + query[i].len, r, r + ref[i].len, res);
    } catch (const std::exception& e) {
      return classifyException(e);
    }
    mb200_aln_result& o(results[i]);
    std::memset(&o, 0, sizeof(o));
    o.score    = res.score;
    o.beginPos = res.align.beginPos;
    o.isJumped = res.isJumped ? 1 : 0;
    ok &= emitPath(res.align.apath, cigarPool, cigarCap, used, o.cigarOff, o.cigarLen);
  }
  *cigarUsed = used;
  return ok ? MB200_OK : MB200_ERR_POOL_OVERFLOW;
}

REF_API int ref_indel_align_batch(const mb200_scores* scores, int32_t largeIndelScore, uint32_t n,
                                  const uint8_t* pool, const mb200_seqview* query, const mb200_seqview* ref,
                                  mb200_aln_result* results, uint32_t* cigarPool, uint64_t cigarCap,
                                  uint64_t* cigarUsed)
{
  const GlobalLargeIndelAligner<int> aligner(toScores(scores), largeIndelScore);
  return singleRefBatch(aligner, n, pool, query, ref, results, cigarPool, cigarCap, cigarUsed);
}

REF_API int ref_global_align_batch(const mb200_scores* scores, uint32_t n, const uint8_t* pool,
                                   const mb200_seqview* query, const mb200_seqview* ref,
                                   mb200_aln_result* results, uint32_t* cigarPool, uint64_t cigarCap,
                                   uint64_t* cigarUsed)
{
  const GlobalAligner<int> aligner(toScores(scores));
  return singleRefBatch(aligner, n, pool, query, ref, results, cigarPool, cigarCap, cigarUsed);
}

/// fill the two table pairs the scorer uses: index 0 = qscore_snp(snpPrior0), 1 = qscore_snp(snpPrior1)
REF_API void ref_qscore_tables(double snpPrior0, double snpPrior1, mb200_qscore_tables* out)
{
  const double priors[2] = {snpPrior0, snpPrior1};
  for (int t(0); t < 2; ++t) {
    const qscore_snp qs(priors[t]);
    for (int q(0); q <= MB200_MAX_QSCORE; ++q) {
      out->lnComp[t][q] = qs.qphred_to_ln_comp_error_prob(q);
      out->lnErr[t][q]  = qs.qphred_to_ln_error_prob(q);
    }
  }
}

REF_API int ref_split_align_batch(double snpPrior0, double snpPrior1, uint32_t flankScoreSize, uint32_t n,
                                  const uint8_t* pool, const mb200_sr_job* jobs, mb200_sr_result* results)
{
  const qscore_snp qs0(snpPrior0);
  const qscore_snp qs1(snpPrior1);
  for (uint32_t i(0); i < n; ++i) {
    const mb200_sr_job& job(jobs[i]);
    const std::string   q(reinterpret_cast<const char*>(pool + job.query.off), job.query.len);
    const std::string   t(reinterpret_cast<const char*>(pool + job.target.off), job.target.len);
    const known_pos_range2 range(job.bpBegin, job.bpEnd);
    SRAlignmentInfo        info;
    mb200_sr_result&       o(results[i]);
    std::memset(&o, 0, sizeof(o));
    try {
      splitReadAligner(flankScoreSize, q, (job.table ? qs1 : qs0), pool + job.qualOff, t, range, info);
    } catch (const std::exception& e) {
      const std::string what(e.what());
      if (what.find("scanEnd < scanStart") != std::string::npos)
        o.status = MB200_ERR_SPLIT_SCAN;
      else if (what.find("alignment input") != std::string::npos)
        o.status = MB200_ERR_SPLIT_INPUT;
      else if (what.find("alignment outcome") != std::string::npos)
        o.status = MB200_ERR_SPLIT_OUTCOME;
      else
        o.status = MB200_ERR_BAD_QUAL;
      continue;
    }
    o.alignPos        = info.alignPos;
    o.leftSize        = info.leftSize;
    o.homSize         = info.homSize;
    o.rightSize       = info.rightSize;
    o.leftMismatches  = info.leftMismatches;
    o.homMismatches   = info.homMismatches;
    o.rightMismatches = info.rightMismatches;
    o.alignScore      = info.alignScore;
    o.alignLnLhood    = info.alignLnLhood;
    o.isEvidence      = info.isEvidence ? 1 : 0;
    o.isTier2Evidence = info.isTier2Evidence ? 1 : 0;
    o.evidence        = info.evidence;
  }
  return MB200_OK;
}

/// Run runIterativeAssembler per locus.  Any input byte is accepted (the reference works on std::string).
REF_API int ref_assemble_batch(const mb200_asm_opts* opts, uint32_t nLoci, const uint64_t* readOff,
                               const mb200_seqview* readViews, const uint8_t* pool, mb200_asm_locus* loci,
                               mb200_contig* contigs, uint64_t contigCap, mb200_seqview* pseudoReads,
                               uint64_t pseudoCap, uint8_t* contigPool, uint64_t contigPoolCap,
                               uint32_t* setPool, uint64_t setPoolCap)
{
  IterativeAssemblerOptions opt;
  opt.minWordLength           = opts->minWordLength;
  opt.maxWordLength           = opts->maxWordLength;
  opt.wordStepSize            = opts->wordStepSize;
  opt.minCoverage             = opts->minCoverage;
  opt.minConservativeCoverage = opts->minConservativeCoverage;
  opt.maxAssemblyCount        = opts->maxAssemblyCount;
  opt.minUnusedReads          = opts->minUnusedReads;
  opt.minSupportReads         = opts->minSupportReads;

  uint64_t nContig(0), nPseudo(0), poolUsed(0), setUsed(0);
  bool     ok(true);
  for (uint32_t li(0); li < nLoci; ++li) {
    AssemblyReadInput reads;
    for (uint64_t r(readOff[li]); r < readOff[li + 1]; ++r) {
      reads.emplace_back(reinterpret_cast<const char*>(pool + readViews[r].off), readViews[r].len);
    }
    const unsigned     normalReadCount(reads.size());
    AssemblyReadOutput readInfo;
    Assembly           asmContigs;
    runIterativeAssembler(opt, reads, readInfo, asmContigs);

    mb200_asm_locus& L(loci[li]);
    std::memset(&L, 0, sizeof(L));
    L.firstContig = nContig;
    L.nContigs    = asmContigs.size();
    L.firstPseudo = nPseudo;
    L.nPseudo     = reads.size() - normalReadCount;
    L.setWords    = (normalReadCount + 2 * opts->maxAssemblyCount + 31) / 32;

    auto putBytes = [&](const std::string& s, uint64_t& off) {
      off = poolUsed;
      if (poolUsed + s.size() <= contigPoolCap) std::memcpy(contigPool + poolUsed, s.data(), s.size());
      poolUsed += s.size();
    };
    auto putSet = [&](const std::set<unsigned>& s, uint64_t& off) {
      off = setUsed;
      if (setUsed + L.setWords <= setPoolCap) {
        std::memset(setPool + setUsed, 0, sizeof(uint32_t) * L.setWords);
        for (const unsigned rd : s) setPool[setUsed + rd / 32] |= (1u << (rd % 32));
      }
      setUsed += L.setWords;
    };

    for (const AssembledContig& c : asmContigs) {
      if (nContig < contigCap) {
        mb200_contig& o(contigs[nContig]);
        std::memset(&o, 0, sizeof(o));
        putBytes(c.seq, o.seqOff);
        o.seqLen            = c.seq.size();
        o.seedReadCount     = c.seedReadCount;
        o.conservativeBegin = c.conservativeRange.begin_pos();
        o.conservativeEnd   = c.conservativeRange.end_pos();
        putSet(c.supportReads, o.supportOff);
        putSet(c.rejectReads, o.rejectOff);
      }
      nContig++;
    }
    for (unsigned r(normalReadCount); r < reads.size(); ++r) {
      if (nPseudo < pseudoCap) {
        mb200_seqview& v(pseudoReads[nPseudo]);
        v.reserved = 0;
        putBytes(reads[r], v.off);
        v.len = reads[r].size();
      }
      nPseudo++;
    }
  }
  ok = (nContig <= contigCap) && (nPseudo <= pseudoCap) && (poolUsed <= contigPoolCap) && (setUsed <= setPoolCap);
  return ok ? MB200_OK : MB200_ERR_POOL_OVERFLOW;
}

/// iteration order of a real std::unordered_map<std::string,unsigned> filled with operator[] in the given
/// key order (pins the libstdc++ emulation of oracle/orc_assemble.c)
REF_API void ref_unordered_order(uint32_t n, const uint8_t* keys, uint32_t keyLen, uint32_t* order)
{
  std::unordered_map<std::string, unsigned> m;
  for (uint32_t i(0); i < n; ++i) {
    m[std::string(reinterpret_cast<const char*>(keys + static_cast<size_t>(i) * keyLen), keyLen)] = i;
  }
  uint32_t k(0);
  for (const auto& kv : m) order[k++] = kv.second;
}

REF_API const char* ref_version()
{
  return "Illumina/manta reference sources (v1.6.0) via oracle/ref_shim.cpp";
}
