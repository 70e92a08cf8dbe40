"""CUDA split-read scan (through the C ABI) against the golden vectors and the oracle (bit-exact floats)."""
import ctypes as C

import numpy as np
import pytest

import manta_b200
from manta_b200 import _cabi as cb
from manta_b200 import synth
from test_oracle_golden import check_split, golden_tables, load, split_jobs

pytestmark = pytest.mark.gpu


def test_splitread_golden_gpu(gpu):
    g = load("splitread.json")
    check_split(gpu, golden_tables(g), g)


def test_split_vs_oracle(gpu, orc):
    cases = synth.random_split_cases(seed=21, n=4000)
    pool, jobs = split_jobs(cases)
    t = np.zeros(1, dtype=cb.QSCORE_TABLES_DT)
    orc.lib.orc_qscore_tables(C.c_double(1e-3), C.c_double(0.0), t.ctypes.data_as(C.c_void_p))
    for flank in (50, 2, 0, 200):
        a = gpu.split_align_raw(t, (1e-3, 0.0), flank, pool, jobs)
        b = orc.split_align_raw(t, (1e-3, 0.0), flank, pool, jobs)
        assert a.tobytes() == b.tobytes(), flank


def test_host_tables_match_oracle(orc):
    """the python qscore_snp mirror (host libm) reproduces the reference tables bit for bit"""
    t = np.zeros(1, dtype=cb.QSCORE_TABLES_DT)
    orc.lib.orc_qscore_tables(C.c_double(1e-3), C.c_double(0.0), t.ctypes.data_as(C.c_void_p))
    mine = manta_b200.qscore_tables(manta_b200.qscore_snp(1e-3), manta_b200.qscore_snp(0.0))
    assert mine.tobytes() == t.tobytes()


def test_split_reference_api(gpu):
    """GSC/test/SplitReadAlignmentTest.cpp:386-428 through the reference-shaped front end"""
    q = "TCTATCACCCATCGTACCACTCACGGGAGCTCTCC"
    target = ("GATCACAGGTCTATCACCCTATTAACCACTCACGGGAGCTCTCCATGCATTTGGT"
              "ATTTTCGTCTGGGGGGTGTGCACGCGATAGCATTGCGAGACGCTGGA")
    qs = manta_b200.qscore_snp(1e-3)
    with pytest.raises(cb.Mb200Error):
        manta_b200.splitReadAligner(2, q, qs, [30] * 35, "GATCACAGGTCTAT", (8, 50), ctx=gpu)
    with pytest.raises(cb.Mb200Error):
        manta_b200.splitReadAligner(2, q, qs, [30] * 35, target, (1000, 1050), ctx=gpu)
    r = manta_b200.splitReadAligner(2, q, qs, [30] * 35, target, (8, 50), ctx=gpu)
    assert (r.leftSize, r.leftMismatches, r.rightSize, r.rightMismatches, r.homSize, r.homMismatches, r.alignScore) == \
        (0, 0, 34, 27, 1, 1, 7)


def test_split_microbench_shape(gpu, orc):
    """BASELINE configs[4] shape (reads x 4 targets of 1200 bp) from the bench generator"""
    pool, jobs = synth.make_split_batch(5000)
    t = manta_b200.qscore_tables(manta_b200.qscore_snp(1e-3), manta_b200.qscore_snp(0.0))
    a = gpu.split_align_raw(t, (1e-3, 0.0), 50, pool, jobs)
    b = orc.split_align_raw(t, (1e-3, 0.0), 50, pool, jobs)
    assert a.tobytes() == b.tobytes()
    assert int((a["status"] == 0).sum()) == len(jobs) and a["isEvidence"].mean() > 0.1
