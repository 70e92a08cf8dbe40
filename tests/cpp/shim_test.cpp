// Replays known-answer vectors of the reference's own unit tests through the C++ shim (reference call shapes):
//   lib/alignment/test/GlobalJumpAlignerTest.cpp:68-129, GlobalLargeIndelAlignerTest.cpp:42-76,
//   lib/assembly/test/IterativeAssemblerTest.cpp:132-170, GSC/test/SplitReadAlignmentTest.cpp:386-428
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "manta_b200_shim.hpp"

#define REQUIRE(x)                                                     \
  do {                                                                 \
    if (!(x)) {                                                        \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); \
      std::exit(1);                                                    \
    }                                                                  \
  } while (0)

int main()
{
  mb200::Context ctx(0);
  {
    typedef int                       score_t;
    mb200::AlignmentScores<score_t>   scores(2, -4, -5, -1, -1);
    mb200::GlobalJumpAligner<score_t> aligner(ctx, scores, -3);
    mb200::JumpAlignmentResult<score_t> result;
    const std::string seq("ABABAABABACDCDCDyCDCDC"), ref1("xABABABABABAx"), ref2("xCDCDCDCDCDCDCx");
    aligner.align(seq.begin(), seq.end(), ref1.begin(), ref1.end(), ref2.begin(), ref2.end(), result);
    REQUIRE(mb200::apath_to_cigar(result.align1.apath) == "5=1D5=");
    REQUIRE(result.align1.beginPos == 1);
    REQUIRE(mb200::apath_to_cigar(result.align2.apath) == "6=1I5=");
    REQUIRE(result.align2.beginPos == 1);
    bool threw(false);
    try {
      const std::string empty;
      aligner.align(empty.begin(), empty.end(), ref1.begin(), ref1.end(), ref2.begin(), ref2.end(), result);
    } catch (const mb200::GeneralException& e) {
      threw = (std::string(e.what()) == "Unexpected empty query sequence");
    }
    REQUIRE(threw);
  }
  {
    mb200::AlignmentScores<int>         scores(2, -4, -5, -1, -4);
    mb200::GlobalLargeIndelAligner<int> aligner(ctx, scores, -10);
    mb200::AlignmentResult<int>         result;
    const std::string seq("BCDEFHIKLM"), ref("ABCDEFGHIKLMN");
    aligner.align(seq.begin(), seq.end(), ref.begin(), ref.end(), result);
    REQUIRE(mb200::apath_to_cigar(result.align.apath) == "5=1D5=");
    REQUIRE(result.align.beginPos == 1);
  }
  {
    mb200::IterativeAssemblerOptions opt;
    opt.minWordLength = 6; opt.maxWordLength = 6; opt.minCoverage = 1; opt.minSupportReads = 1; opt.minUnusedReads = 1;
    mb200::AssemblyReadInput reads{"ATATAGACGATG", "ACGATGTCTATCTT", "ACGATGTTGGCCTT"};
    mb200::AssemblyReadOutput readInfo;
    mb200::Assembly           contigs;
    mb200::runIterativeAssembler(ctx, opt, reads, readInfo, contigs);
    REQUIRE(contigs.size() == 2u);
    REQUIRE(contigs[0].seq == "ATATAGACGATGTCTATCTT");
    REQUIRE(contigs[1].seq == "ATATAGACGATGTTGGCCTT");
    REQUIRE(readInfo[0].isUsed && readInfo[0].contigIds.size() == 2 && readInfo[0].contigIds[1] == 1u);
    REQUIRE(readInfo[2].contigIds[0] == 1u);
  }
  {
    // qscore_snp(1e-3) tables computed like lib/blt_util/qscore_snp.cpp:29-41
    double lnComp[71], lnErr[71];
    for (int i = 0; i <= 70; ++i) {
      const double qerr = std::pow(10., -static_cast<double>(i) / 10.);
      const double q2p  = (qerr * (1. - 1e-3 / 3.)) + ((1 - qerr) * 1e-3);
      lnComp[i]         = (std::fabs(-q2p) < 0.01) ? std::log1p(-q2p) : std::log(1 - q2p);
      lnErr[i]          = std::log(q2p);
    }
    const std::string q("TCTATCACCCATCGTACCACTCACGGGAGCTCTCC");
    const std::string t("GATCACAGGTCTATCACCCTATTAACCACTCACGGGAGCTCTCCATGCATTTGGTATTTTCGTCTGGGGGGTGTGCACGCGATAGCATTGCGAGACGCTGGA");
    uint8_t qual[35];
    for (int i = 0; i < 35; ++i) qual[i] = 30;
    mb200::SRAlignmentInfo info;
    mb200::splitReadAligner(ctx, 2, q, lnComp, lnErr, qual, t, 8, 50, info);
    REQUIRE(info.leftSize == 0 && info.rightSize == 34 && info.rightMismatches == 27 && info.homSize == 1 && info.alignScore == 7);
    bool threw(false);
    try {
      mb200::splitReadAligner(ctx, 2, q, lnComp, lnErr, qual, t, 1000, 1050, info);
    } catch (const mb200::GeneralException&) {
      threw = true;
    }
    REQUIRE(threw);
  }
  {
    // refiner rows a5 / a9 on a planted 60-base deletion: trim the window, align, QC + select
    std::string refBuild;
    unsigned    x(12345);
    for (int i = 0; i < 1200; ++i) {
      x = x * 1664525u + 1013904223u;
      refBuild.push_back("ACGT"[(x >> 24) & 3]);
    }
    const std::string ref(refBuild);
    const std::string contig(ref.substr(400, 150) + ref.substr(610, 150));
    int               lead(0), trail(0);
    mb200::trimReferenceWindow(ctx, contig, ref, 100, 100, 800, 800, lead, trail);
    REQUIRE(lead == 400 && trail == 1200 - 760);
    mb200::AlignmentScores<int>         scores(2, -8, -24, -1, -1);
    mb200::GlobalLargeIndelAligner<int> aligner(ctx, scores, -100);
    std::vector<mb200::AlignmentResult<int>> alignments(2);
    aligner.align(contig.begin(), contig.end(), ref.begin() + lead, ref.end() - trail, alignments[0]);
    alignments[0].align.beginPos += lead;
    REQUIRE(mb200::apath_to_cigar(alignments[0].align.apath) == "150=60D150=");
    const std::string flat(ref.substr(300, 200));  // second contig: no variant
    aligner.align(flat.begin(), flat.end(), ref.begin(), ref.end(), alignments[1]);
    std::vector<mb200::CandidateSegmentSetType> segments;
    unsigned                                    bestIndex(99);
    const mb200::AlignmentScores<int>           filterScores(2, -8, -18, 0, -1);
    REQUIRE(mb200::selectSmallSVContig(ctx, filterScores, 10, {flat, contig}, {7, 9}, {alignments[1], alignments[0]}, ref, segments, bestIndex));
    REQUIRE(bestIndex == 1u && segments[0].empty() && segments[1].size() == 1 && segments[1][0].first == 1u && segments[1][0].second == 1u);
    // refined SV of the selected contig (setSmallCandSV): the planted deletion, no homology at the breakpoint expected
    // beyond what the random sequence gives; bp2 - bp1 == deletion length + 1
    const std::vector<mb200_small_sv> svs(mb200::refineSmallSVs(ctx, 10, contig, alignments[0], segments[1], bestIndex, ref));
    REQUIRE(svs.size() == 1 && svs[0].status == 0 && svs[0].contig == 1u && svs[0].segment == 0u);
    REQUIRE(svs[0].refEnd - svs[0].refBegin == 60 && svs[0].readBegin == svs[0].readEnd && svs[0].ciposBegin == 0);
    REQUIRE(svs[0].bp2Begin - svs[0].bp1Begin == 61 && svs[0].bp1End - svs[0].bp1Begin == svs[0].ciposEnd + 1);
    // spanning: two clean 120-base arms pass, a 10-base arm fails the QC
    mb200::JumpAlignmentResult<int> good, bad;
    good.score = 400;
    good.align1.apath = {{MB200_SEG_SEQ_MATCH, 120}};
    good.align2.apath = {{MB200_SEG_SEQ_MATCH, 120}};
    bad = good;
    bad.score = 500;
    bad.align2.apath = {{MB200_SEG_SEQ_MATCH, 10}};
    unsigned best(99);
    REQUIRE(mb200::selectJumpContigDNA(ctx, filterScores, {bad, good}, best) && best == 1u);
    REQUIRE(!mb200::selectJumpContigDNA(ctx, filterScores, {bad}, best));
  }
  {
    // GSC/test/SVCandidateAssemblyRefinerTest.cpp:272-358 (test_getLargeIndelSegments) through the shim's run finder
    auto P = [](std::initializer_list<mb200::path_segment> l) { return mb200::path_t(l); };
    const unsigned M(MB200_SEG_MATCH), I(MB200_SEG_INSERT), D(MB200_SEG_DELETE);
    const mb200::path_t path1(P({{M, 35}, {I, 5}, {M, 30}, {D, 6}, {M, 10}, {I, 3}}));
    REQUIRE(mb200::detail::indelRuns(path1, 10).empty());
    mb200::CandidateSegmentSetType r(mb200::detail::indelRuns(path1, 6));
    REQUIRE(r.size() == 1 && r[0].first == 3 && r[0].second == 3);
    r = mb200::detail::indelRuns(path1, 3);
    REQUIRE(r.size() == 3 && r[0].first == 1 && r[1].first == 3 && r[2].first == 5 && r[2].second == 5);
    r = mb200::detail::indelRuns(P({{M, 35}, {I, 5}, {D, 6}, {M, 40}, {I, 3}}), 4);
    REQUIRE(r.size() == 1 && r[0].first == 1 && r[0].second == 2);
    r = mb200::detail::indelRuns(P({{M, 35}, {I, 5}, {D, 3}, {M, 40}, {I, 3}}), 4);
    REQUIRE(r.size() == 1 && r[0].first == 1 && r[0].second == 2);
    r = mb200::detail::indelRuns(P({{M, 35}, {I, 3}, {D, 5}, {M, 40}, {I, 3}}), 4);
    REQUIRE(r.size() == 1 && r[0].first == 1 && r[0].second == 2);
    REQUIRE(mb200::detail::indelRuns(P({{M, 35}, {I, 3}, {D, 3}, {M, 40}, {I, 3}}), 4).empty());
  }
  std::printf("shim_test ok\n");
  return 0;
}
