"""Row a11: getReadSplitScore / incrementSplitReadEvidence (GSC/SVScorerSplit.cpp:42-267).  CPU: the oracle's restatement
against the reference's own statics (oracle/_ref ref_split_score_batch: the unmodified incrementSplitReadEvidence on the
reference's splitReadAligner results and evidence objects).  GPU: mb200_split_score_batch against the oracle, byte for byte
(float fields by bit pattern): the four SRAlignmentInfo per read, the per-read evidence entries, the per-candidate counts."""
import numpy as np
import pytest

import manta_b200
from manta_b200 import _cabi as cb
from manta_b200 import synth


def _tables():
    return manta_b200.qscore_tables(manta_b200.qscore_snp(1e-3), manta_b200.qscore_snp(0.0))


def _same(a, b):
    for x, y, name in zip(a, b, ("sr", "evidence", "counts")):
        assert x.tobytes() == y.tobytes(), name


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_vs_reference_statics(orc, ref, seed):
    pool, cands, readFirst, reads = synth.make_split_candidates(40, seed=seed)
    a = orc.split_score_raw(_tables(), 50, 15, 5, cands, readFirst, reads, pool)
    _same(a, ref.split_score_raw(_tables(), 50, 15, 5, cands, readFirst, reads, pool))
    ev = a[1]
    assert (ev["status"] == cb.MB200_SPLIT_DUPLICATE).sum() == 40 * 5 and ev["countedAlt"].sum() > 100 and ev["countedRef"].sum() > 100
    assert ev["isTier2SplitSupport"].sum() > ev["isSplitSupport"].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,flank,mapq", [(1, 50, (15, 5)), (2, 50, (30, 10)), (3, 20, (0, 0)), (4, 100, (61, 61))])
def test_gpu_vs_oracle(gpu, orc, seed, flank, mapq):
    pool, cands, readFirst, reads = synth.make_split_candidates(64, seed=seed, reads_per_cand=24 + 8 * seed)
    _same(gpu.split_score_raw(_tables(), flank, mapq[0], mapq[1], cands, readFirst, reads, pool),
          orc.split_score_raw(_tables(), flank, mapq[0], mapq[1], cands, readFirst, reads, pool))


@pytest.mark.gpu
def test_gpu_ragged_and_failing_reads(gpu, orc):
    """candidates without reads, a read longer than a target (the aligner's input error), a read with an out-of-range
    quality, all-N reads"""
    pool, cands, readFirst, reads = synth.make_split_candidates(12, seed=9, reads_per_cand=16)
    pool = pool.copy()
    readFirst = readFirst.copy()
    readFirst[5:] = readFirst[5]                      # candidates 4.. have no reads (5.. share the same empty range)
    n = int(readFirst[-1])
    reads = reads[:n].copy()
    cands["altBp2"]["len"][1] = 100                    # shorter than the reads: MB200_ERR_SPLIT_INPUT for that target
    pool[int(reads["qualOff"][3]) + 7] = 90           # quality outside the table
    pool[int(reads["off"][6]):int(reads["off"][6]) + int(reads["len"][6])] = ord("N")
    g = gpu.split_score_raw(_tables(), 50, 15, 5, cands, readFirst, reads, pool)
    o = orc.split_score_raw(_tables(), 50, 15, 5, cands, readFirst, reads, pool)
    _same(g, o)
    assert (o[1]["status"] > 1).sum() >= 2 and int(o[2]["alt"]["splitReadCount"][6:].sum()) == 0
