// staged_driver.cpp -- the batched call sequence of INTEGRATION.md section 3 on one synthetic complex candidate:
//   upload -> assemble + trim + align + QC/select (fused, on the device) -> refined small SVs -> SVAlignmentInfo
//   -> split-read realignment of one supporting read against the alt contig and the reference.
// Build:  g++ -std=c++17 -Iinclude examples/staged_driver.cpp manta_b200/libmanta_b200.so -Wl,-rpath,$PWD/manta_b200 -o staged_driver
// (needs an sm_100 device to run; there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "manta_b200_shim.hpp"

int main()
{
  mb200::Context ctx(0);

  // ---- one candidate: a 1.5 kb window, a 120-base deletion at 700, 48 alt-haplotype reads tiling the breakpoint
  std::string window;
  unsigned    x(2024);
  for (int i = 0; i < 1500; ++i) { x = x * 1664525u + 1013904223u; window.push_back("ACGT"[(x >> 24) & 3]); }
  const std::string alt(window.substr(0, 700) + window.substr(820));
  std::vector<uint8_t>       pool;
  std::vector<mb200_seqview> reads;
  for (int r = 0; r < 48; ++r) {
    const std::string read(alt.substr(560 + 2 * r, 150));
    reads.push_back(mb200::detail::append(pool, read.begin(), read.end()));
  }
  const mb200_seqview ref1(mb200::detail::append(pool, window.begin(), window.end()));
  const uint64_t      readOff[2] = {0, reads.size()};
  const uint8_t       isJump(0);
  const mb200_ref_cuts cuts{20, 20, 250, 250};
  ctx.check(mb200_candidates_upload(ctx.get(), 1, readOff, reads.data(), &isJump, &ref1, &ref1, pool.data(), pool.size(), &cuts));

  // ---- fused run: reference defaults (IterativeAssemblerOptions, SVRefinerOptions)
  mb200_candidate_opts opts = {};
  opts.asmOpts         = mb200_asm_opts{41, 76, 5, 1, 2, 10, 3, 2};
  opts.indelScores     = mb200_scores{2, -8, -24, -1, -1, 0};
  opts.largeIndelScore = -100;
  opts.jumpScores      = mb200_scores{2, -8, -12, -1, -1, 0};
  opts.jumpScore       = -100;
  opts.qcOpts          = mb200_qc_opts{mb200_scores{2, -8, -18, 0, -1, 0}, 10, 0};
  const uint32_t A(opts.asmOpts.maxAssemblyCount);
  mb200_asm_locus                locus;
  std::vector<mb200_contig>      contigs(A);
  std::vector<mb200_seqview>     pseudo(2 * A);
  std::vector<uint8_t>           contigPool(1 << 16);
  std::vector<uint32_t>          setPool(1 << 12), cigar(1 << 14);
  std::vector<mb200_aln_result>  aln(A);
  std::vector<mb200_jump_result> jump(A);
  std::vector<mb200_trim_result> trim(A);
  std::vector<mb200_contig_qc>   qc(A);
  int32_t                        best(-1);
  uint64_t                       usedC(0), usedS(0), usedG(0);
  ctx.check(mb200_candidates_run(ctx.get(), &opts, &locus, contigs.data(), pseudo.data(), contigPool.data(), contigPool.size(), &usedC,
                                 setPool.data(), setPool.size(), &usedS, aln.data(), jump.data(), cigar.data(), cigar.size(), &usedG,
                                 trim.data(), qc.data(), &best));
  std::printf("contigs: %u, bestAlignmentIndex: %d\n", locus.nContigs, best);
  if (best < 0) return 1;

  // ---- refined small SVs of the selected contig (setSmallCandSV) and the split-read targets (SVAlignmentInfo)
  uint32_t                    svFirst[2];
  std::vector<mb200_small_sv> svs(8);
  uint64_t                    nSv(0);
  ctx.check(mb200_candidates_small_svs(ctx.get(), svFirst, svs.data(), svs.size(), &nSv));
  if (nSv == 0) return 1;
  const mb200_contig&         bc(contigs[best]);
  const std::string           contig(reinterpret_cast<const char*>(contigPool.data() + bc.seqOff), bc.seqLen);
  mb200::AlignmentResult<int> bestAlign;
  bestAlign.score          = aln[best].score;
  bestAlign.align.beginPos = aln[best].beginPos;
  bestAlign.align.apath    = mb200::detail::toPath(cigar.data(), aln[best].cigarOff, aln[best].cigarLen);
  std::printf("best contig %u bases, alignment %s at %d, deletion ref[%d,%d) cipos [0,%d]\n", bc.seqLen,
              mb200::apath_to_cigar(bestAlign.align.apath).c_str(), aln[best].beginPos, svs[0].refBegin, svs[0].refEnd, svs[0].ciposEnd);
  const mb200::SVAlignmentInfo info(svs[0], bestAlign, contig, window);

  // ---- one supporting read against the alt contig at bp1 (qscore_snp(1e-3) tables like lib/blt_util/qscore_snp.cpp:29-41)
  double lnComp[71], lnErr[71];
  for (int i = 0; i <= 70; ++i) {
    const double qerr = std::pow(10., -static_cast<double>(i) / 10.);
    const double q2p  = (qerr * (1. - 1e-3 / 3.)) + ((1 - qerr) * 1e-3);
    lnComp[i]         = (std::fabs(-q2p) < 0.01) ? std::log1p(-q2p) : std::log(1 - q2p);
    lnErr[i]          = std::log(q2p);
  }
  const std::string    read(alt.substr(630, 150));
  std::vector<uint8_t> qual(read.size(), 30);
  mb200::SRAlignmentInfo sr;
  mb200::splitReadAligner(ctx, 50, read, lnComp, lnErr, qual.data(), info.bp1ContigSeq(), info.bp1ContigOffset.begin_pos(),
                          info.bp1ContigOffset.end_pos(), sr);
  std::printf("split read on alt contig: leftSize %u rightSize %u mismatches %u/%u evidence %.3f\n", sr.leftSize, sr.rightSize,
              sr.leftMismatches, sr.rightMismatches, sr.evidence);
  return 0;
}
