#!/usr/bin/env python
"""Golden vectors on REAL reads, produced by the unmodified reference algorithms.

oracle/build_ref_full.sh --trace builds the reference's GenerateSVCandidates with a call recorder (oracle/ref_trace.hpp,
oracle/trace_calls.patch: one line of JSON after every runIterativeAssembler / GlobalJumpAligner::align /
GlobalLargeIndelAligner::align / GlobalAligner::align / splitReadAligner call); this script runs it on the reference's demo
(oracle/run_demo.sh), checks that the recorder did not change the VCF, and writes tests/golden/demo_calls.json.gz:
every assembler and aligner call, and every SPLIT_STRIDE-th of the 6 448 split-read alignments.

    python tests/golden/make_demo_calls.py          (where /root/reference exists)
"""
import collections
import gzip
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FULL = os.path.join(ROOT, "oracle", "_ref", "full")
SPLIT_STRIDE = 7

subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref_full.sh"), "--trace"])
with tempfile.TemporaryDirectory() as td:
    stock, traced, log = os.path.join(td, "stock"), os.path.join(td, "traced"), os.path.join(td, "calls.jsonl")
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "run_demo.sh"), stock])
    env = dict(os.environ, MANTA_TRACE_CALLS=log, PHASE1_DIR=stock)
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "run_demo.sh"), traced, os.path.join(FULL, "bin", "GenerateSVCandidates_trace")], env=env)
    for v in ("candidateSV", "diploidSV", "somaticSV"):
        a = [l for l in open(os.path.join(stock, v + ".vcf")) if not l.startswith("##")]
        b = [l for l in open(os.path.join(traced, v + ".vcf")) if not l.startswith("##")]
        assert a == b, v
    calls = [json.loads(l) for l in open(log)]
count = collections.Counter(c["kind"] for c in calls)
keep, nsplit = [], 0
for c in calls:
    if c["kind"] == "split":
        nsplit += 1
        if (nsplit - 1) % SPLIT_STRIDE:
            continue
    keep.append(c)
out = dict(source="src/demo (tumor/normal BAMs, --exome, two regions) through the reference's own GenerateSVCandidates with "
                  "oracle/trace_calls.patch; calls recorded: %s; kept: all but the split-read alignments, of which every %d-th; "
                  "split tables: table 0 = qscore_snp(CallOptionsShared::snpPrior), a FLOAT 1e-3f (options/CallOptionsShared.hpp:30), i.e. "
                  "0.0010000000474974513 as the double qscore_snp takes -- not 0.001; table 1 = qscore_snp(0) (alt allele), SVScorer.hpp:47"
                  % (dict(count), SPLIT_STRIDE),
           snpPriors=[float(np.float32(1e-3)), 0.0], cases=keep)
path = os.path.join(ROOT, "tests", "golden", "demo_calls.json.gz")
with gzip.open(path, "wt", compresslevel=9) as f:
    json.dump(out, f, separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes;", collections.Counter(c["kind"] for c in keep))
