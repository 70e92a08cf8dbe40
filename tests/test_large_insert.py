"""Large insertions with an incompletely assembled insert (rest of row a9: GSC/SVCandidateAssemblyRefiner.cpp:2079-2112 and
processLargeInsertion :833-1007).  CPU: the oracle's restatement against the reference's own statics (oracle/_ref,
ref_large_insertion_batch) on assembled loci and on perturbed inputs that drive every exit of processLargeInsertion.
GPU: mb200_large_insertion_batch against the oracle, every field and every CIGAR."""
import numpy as np
import pytest

import qc_cases
from manta_b200 import _cabi as cb
from manta_b200 import synth

EDGE = (2, -8, -18, -1, -1, 0)        # largeInsertEdgeAlignScores      lib/options/SVRefinerOptions.hpp:41
COMPLETE = (2, -8, -100, 0, -1, 0)    # largeInsertCompleteAlignScores  :42


class LiCase:
    """a candidate batch pushed through the checker's assembler / trim / large-indel aligner / QC / small-SV steps, laid out
    for the large-insertion call"""

    def __init__(self, orc, batch):
        opts = cb.default_candidate_opts()
        self.cuts = cb.default_ref_cuts(batch.nLoci)
        case, _ = qc_cases.pipeline_case(orc, batch, opts, self.cuts)
        qc, best = case.run(orc)
        self.svFirst, self.svs = case.small_svs(orc, qc, best)
        loci, contigs, _, _, _ = case.asm
        self.cons = np.zeros((max(1, len(case.contigSeq)), 2), dtype=np.int32)
        for li in range(len(case.nContigs)):
            for k in range(int(case.nContigs[li])):
                c = contigs[int(loci["firstContig"][li]) + k]
                self.cons[int(case.slotFirst[li]) + k] = (int(c["conservativeBegin"]), int(c["conservativeEnd"]))
        self.case = case

    def run(self, binding, **over):
        c = self.case
        a = dict(contigSeq=c.contigSeq, cons=self.cons, cuts=self.cuts, svFirst=self.svFirst, svs=self.svs)
        a.update(over)
        return binding.large_insertion_raw(EDGE, COMPLETE, c.isJump, c.slotFirst, c.nContigs, c.aln, c.cigar, c.pool,
                                           a["contigSeq"], a["cons"], c.ref1, a["cuts"], a["svFirst"], a["svs"])


def _equal(a, b, merge_qc_exits=False):
    """(info, records, cigar pool) of two implementations: every field, CIGARs by content"""
    ia, oa, ca = a
    ib, ob, cb_ = b
    assert ia.tobytes() == ib.tobytes()
    sa, sb = oa["status"].copy(), ob["status"].copy()
    if merge_qc_exits:  # the reference's two QC exits (:955, :966-973) leave the same state behind
        sa[sa == cb.LI_TRIM_QC] = cb.LI_FLANK_QC
    assert sa.tolist() == sb.tolist()
    for i in range(len(oa)):
        x, y = oa[i], ob[i]
        for f in oa.dtype.names:
            if f in ("status", "cigarOff"):
                continue
            if merge_qc_exits and x["status"] in (cb.LI_FLANK_QC, cb.LI_TRIM_QC) and f.startswith("insertTrim"):
                continue
            assert x[f] == y[f], (i, f, int(x["status"]))
        assert ca[int(x["cigarOff"]):int(x["cigarOff"]) + int(x["cigarLen"])].tolist() == \
            cb_[int(y["cigarOff"]):int(y["cigarOff"]) + int(y["cigarLen"])].tolist(), i
    return oa["status"]


def _variants(lc, rng):
    """perturbations that push processLargeInsertion through its other exits"""
    c = lc.case
    yield {}
    # right contig replaced by an unrelated one: the joined contig no longer aligns as one clean insertion
    shuffled = c.contigSeq.copy()
    perm = rng.permutation(len(shuffled))
    shuffled["off"], shuffled["len"] = c.contigSeq["off"][perm], c.contigSeq["len"][perm]
    yield dict(contigSeq=shuffled)
    # contigs cut short on the right: short flanks / short partial insert sequences
    short = c.contigSeq.copy()
    short["len"] = np.maximum(60, short["len"] - rng.integers(0, 120, len(short))).astype(np.uint32)
    yield dict(contigSeq=short)
    # narrow conservative ranges
    cons = lc.cons.copy()
    cons[:, 0] += rng.integers(0, 60, len(cons))
    cons[:, 1] = np.maximum(cons[:, 0] + 1, cons[:, 1] - rng.integers(0, 60, len(cons)))
    yield dict(cons=cons)
    yield dict(cuts=None, svFirst=None, svs=None)


def _with_duplicates(lc, accepted_records):
    """small-SV records claiming a fully assembled insertion at the start of every accepted large insertion"""
    n = lc.case.isJump.size
    svFirst = np.zeros(n + 1, dtype=np.uint32)
    recs = []
    for li in range(n):
        svFirst[li] = len(recs)
        r = accepted_records[li]
        if r["status"] == cb.LI_ACCEPTED and li % 2 == 0:
            v = np.zeros(1, dtype=cb.SMALL_SV_DT)
            v["locus"], v["bp1Begin"], v["bp2Begin"], v["readBegin"], v["readEnd"] = li, r["bp1Begin"], r["bp1Begin"] + 1, 10, 80
            recs.append(v[0])
    svFirst[n] = len(recs)
    return svFirst, (np.array(recs, dtype=cb.SMALL_SV_DT) if recs else np.zeros(0, dtype=cb.SMALL_SV_DT))


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_vs_reference_statics(orc, ref, seed):
    lc = LiCase(orc, synth.make_large_insertion_batch(40, seed=seed))
    rng = np.random.default_rng(seed)
    seen = set()
    base = None
    for over in _variants(lc, rng):
        st = _equal(lc.run(orc, **over), lc.run(ref, **over), merge_qc_exits=True)
        seen |= set(int(x) for x in st)
        if base is None:
            base = lc.run(orc)[1]
    svFirst, svs = _with_duplicates(lc, base)
    st = _equal(lc.run(orc, svFirst=svFirst, svs=svs), lc.run(ref, svFirst=svFirst, svs=svs), merge_qc_exits=True)
    seen |= set(int(x) for x in st)
    assert (base["status"] == cb.LI_ACCEPTED).sum() >= 10
    assert {cb.LI_NONE, cb.LI_DUPLICATE, cb.LI_ACCEPTED} <= seen and len(seen) >= 4, seen


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gpu_vs_oracle(gpu, orc, seed):
    lc = LiCase(orc, synth.make_large_insertion_batch(48, seed=seed))
    rng = np.random.default_rng(100 + seed)
    seen = set()
    base = None
    for over in _variants(lc, rng):
        st = _equal(lc.run(gpu, **over), lc.run(orc, **over))
        seen |= set(int(x) for x in st)
        if base is None:
            base = lc.run(orc)[1]
    svFirst, svs = _with_duplicates(lc, base)
    seen |= set(int(x) for x in _equal(lc.run(gpu, svFirst=svFirst, svs=svs), lc.run(orc, svFirst=svFirst, svs=svs)))
    assert (base["status"] == cb.LI_ACCEPTED).sum() >= 10 and len(seen) >= 4


@pytest.mark.gpu
def test_gpu_config_batch_and_bad_args(gpu, orc):
    """the bench's generator (mixed complex / spanning candidates, mostly no large insertion) and argument checks"""
    lc = LiCase(orc, synth.make_candidate_batch(40, seed=3))
    _equal(lc.run(gpu), lc.run(orc))
    bad = lc.cons.copy()
    bad[0, 1] = bad[0, 0] - 1
    with pytest.raises(cb.Mb200Error):
        lc.run(gpu, cons=bad)
