#!/usr/bin/env bash
# Drive the reference's executables (oracle/_ref/full/bin, built by build_ref_full.sh) over the reference's demo inputs
# the way src/python/lib/mantaWorkflow.py composes the commands for
#   configManta.py --normalBam N --tumorBam T --referenceFasta F --region=8:107652000-107655000
#                  --region=11:94974000-94989000 --exome           (src/demo/runMantaWorkflowDemo.py:64-71)
# (pyflow / python2 are not in this image).  usage: run_demo.sh <outdir> [GenerateSVCandidates binary]
# MIN_EDGE_OBS / MIN_SPANNING / MIN_SCORED override the workflow defaults (3 / 3 / 50) for the "relaxed" run of
# tests/test_vcf_demo.py, which pushes many more (noise) candidates through assembly, alignment and scoring.
#
# TEST INFRASTRUCTURE ONLY.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
FULL="$HERE/_ref/full"
BIN="$FULL/bin"
D="$FULL/demo"
OUT="${1:?output directory}"
GSC="${2:-$BIN/GenerateSVCandidates}"
PHASE1="${PHASE1_DIR:-$OUT}"    # where alignmentStats.xml / svLocusGraph.bin live (reused by the B200 run)
mkdir -p "$OUT"
N="$D/HCC1954.NORMAL.30x.compare.COST16011_region.bam"
T="$D/G15512.HCC1954.1.COST16011_region.bam"
F="$D/Homo_sapiens_assembly19.COST16011_region.fa"
[ -f "$F" ] || tar -xjf "$F.tar.bz2" -C "$D"

if [ ! -f "$PHASE1/svLocusGraph.bin" ]; then
  mkdir -p "$PHASE1"
  # mantaWorkflow.py runStats (:178-215)
  "$BIN/GetAlignmentStats" --ref "$F" --output-file "$PHASE1/alignmentStats.000.xml" --align-file "$N" >"$PHASE1/stats.log" 2>&1
  "$BIN/GetAlignmentStats" --ref "$F" --output-file "$PHASE1/alignmentStats.001.xml" --align-file "$T" >>"$PHASE1/stats.log" 2>&1
  "$BIN/MergeAlignmentStats" --output-file "$PHASE1/alignmentStats.xml" \
      --align-stats-file "$PHASE1/alignmentStats.000.xml" --align-stats-file "$PHASE1/alignmentStats.001.xml" >>"$PHASE1/stats.log" 2>&1
  # runLocusGraph (:233-300); --exome turns the high-depth filter off (:802), so no --chrom-depth
  i=0
  : > "$PHASE1/graphList.txt"
  for region in 8:107652000-107655000 11:94974000-94989000; do
    g="$PHASE1/svLocusGraph.$i.bin"
    "$BIN/EstimateSVLoci" --output-file "$g" --align-stats "$PHASE1/alignmentStats.xml" --region "$region" \
        --min-candidate-sv-size 8 --min-edge-observations "${MIN_EDGE_OBS:-3}" --ref "$F" --align-file "$N" --tumor-align-file "$T" >>"$PHASE1/graph.log" 2>&1
    echo "$g" >> "$PHASE1/graphList.txt"
    i=$((i+1))
  done
  "$BIN/MergeSVLoci" --output-file "$PHASE1/svLocusGraph.bin" --graph-file-list "$PHASE1/graphList.txt" >>"$PHASE1/graph.log" 2>&1
  "$BIN/CheckSVLoci" --graph-file "$PHASE1/svLocusGraph.bin" >>"$PHASE1/graph.log" 2>&1
fi

# runHyGen (:463-577): one bin, tumor/normal mode, cancer calling => no remote read retrieval
"$GSC" --threads "${THREADS:-1}" --align-stats "$PHASE1/alignmentStats.xml" --graph-file "$PHASE1/svLocusGraph.bin" \
    --bin-index 0 --bin-count 1 --max-edge-count 10 --min-candidate-sv-size 8 --min-candidate-spanning-count "${MIN_SPANNING:-3}" \
    --min-scored-sv-size "${MIN_SCORED:-50}" --ref "$F" --candidate-output-file "$OUT/candidateSV.vcf" \
    --diploid-output-file "$OUT/diploidSV.vcf" --min-qual-score 10 --min-pass-qual-score 20 --min-pass-gt-score 15 \
    --somatic-output-file "$OUT/somaticSV.vcf" --min-somatic-score 10 --min-pass-somatic-score 30 \
    --edge-runtime-log "$OUT/edgeRuntimeLog.txt" --edge-stats-log "$OUT/svCandidateGenerationStats.xml" \
    --edge-stats-report "$OUT/svCandidateGenerationStats.tsv" \
    --align-file "$N" --tumor-align-file "$T" >"$OUT/gsc.log" 2>&1
