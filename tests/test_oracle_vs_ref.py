"""The C restatement against the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh) on seeded
random inputs.  Skipped where the reference library has not been built."""
import ctypes as C
import random

import numpy as np
import pytest

from manta_b200 import _cabi as cb
from manta_b200 import synth
from test_oracle_golden import split_jobs

SCORE_SETS = [((2, -4, -5, -1, -1, 0), -3), ((2, -8, -12, -1, -1, 0), -100), ((2, -8, -24, -1, -1, 0), -100),
              ((2, -8, -18, 0, -1, 1), -50), ((1, -4, -5, -1, -2, 1), -10)]


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_aligners_vs_ref(orc, ref, seed):
    for si, (sc, extra) in enumerate(SCORE_SETS):
        S = cb.Scores(*sc)
        if not sc[5]:
            cases = synth.random_alignment_cases(seed * 100 + si, 25, "jump")
            assert orc.jump_align(S, extra, cases) == ref.jump_align(S, extra, cases)
        cases = synth.random_alignment_cases(seed * 100 + si + 50, 25, "indel")
        assert orc.single_align("indel", S, extra, cases) == ref.single_align("indel", S, extra, cases)
        assert orc.single_align("global", S, 0, cases) == ref.single_align("global", S, 0, cases)


def test_long_query_sentinel_regime(orc, ref):
    """q > ~1250: scores fall below badVal=-10000 and the sentinel can win a max"""
    rng = random.Random(5)
    q = "".join(rng.choice("ACGT") for _ in range(1600))
    r = "".join(rng.choice("ACGT") for _ in range(300))
    S = cb.Scores(2, -8, -24, -1, -1, 0)
    assert orc.single_align("indel", S, -100, [(q, r)]) == ref.single_align("indel", S, -100, [(q, r)])
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    assert orc.jump_align(S, -100, [(q, r, r[::-1])]) == ref.jump_align(S, -100, [(q, r, r[::-1])])


def test_empty_input_errors(orc, ref):
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    for b in (orc, ref):
        for triple, st in [(("", "A", "A"), cb.MB200_ERR_EMPTY_QUERY), (("A", "", "A"), cb.MB200_ERR_EMPTY_REF1),
                           (("A", "A", ""), cb.MB200_ERR_EMPTY_REF2)]:
            with pytest.raises(cb.Mb200Error) as e:
                b.jump_align(S, -100, [triple])
            assert e.value.status == st
        with pytest.raises(cb.Mb200Error) as e:
            b.single_align("global", S, 0, [("A", "")])
        assert e.value.status == cb.MB200_ERR_EMPTY_REF


def test_split_vs_ref(orc, ref):
    cases = synth.random_split_cases(seed=21, n=1500)
    pool, jobs = split_jobs(cases)
    t = np.zeros(1, dtype=cb.QSCORE_TABLES_DT)
    ref.lib.ref_qscore_tables(C.c_double(1e-3), C.c_double(0.0), t.ctypes.data_as(C.c_void_p))
    a = ref.split_align_raw(t, (1e-3, 0.0), 50, pool, jobs)
    b = orc.split_align_raw(t, (1e-3, 0.0), 50, pool, jobs)
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("seed", [31, 32])
def test_assembler_vs_ref(orc, ref, seed):
    for opts, reads in synth.random_assembly_cases(seed, 40):
        o = cb.default_asm_opts(**opts)
        assert orc.assemble(o, [reads]) == ref.assemble(o, [reads])


def test_unordered_order_vs_ref(orc, ref):
    rng = random.Random(7)
    for n in [0, 1, 13, 14, 29, 30, 59, 60, 127, 128, 1000, 5087, 5088, 30000]:
        k = rng.choice([5, 8, 41, 46, 76]) if n <= 128 else rng.choice([12, 41, 46, 76])
        keys = set()
        while len(keys) < n:
            keys.add("".join(rng.choice("ACGT") for _ in range(k)))
        keys = sorted(keys)
        rng.shuffle(keys)
        buf = np.frombuffer("".join(keys).encode() + b"\0", dtype=np.uint8).copy()
        o1 = np.zeros(max(n, 1), dtype=np.uint32)
        o2 = np.zeros(max(n, 1), dtype=np.uint32)
        ref.lib.ref_unordered_order(C.c_uint32(n), buf.ctypes.data_as(C.c_void_p), C.c_uint32(k), o1.ctypes.data_as(C.c_void_p))
        orc.lib.orc_unordered_order(C.c_uint32(n), buf.ctypes.data_as(C.c_void_p), C.c_uint32(k), o2.ctypes.data_as(C.c_void_p))
        assert (o1 == o2).all(), n


# ---- rows a5 / a9: reference-window trim, contig QC and bestAlignmentIndex selection -------------------------------
def _same_qc(a, b):
    (qa, ba), (qb, bb) = a, b
    assert ba.tolist() == bb.tolist()
    assert qa.tobytes() == qb.tobytes()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_select_contigs_fuzz_vs_ref(orc, ref, seed):
    import qc_cases
    case = qc_cases.fuzz_case(seed, 250)
    got, want = case.run(orc), case.run(ref)
    _same_qc(got, want)
    qc, best = want
    # the fuzz must actually reach the interesting outcomes
    cplx = np.repeat(case.isJump == 0, case.nContigs)
    assert qc["isCandidate"][cplx].sum() > 5 and (qc["nSegments"] > 1).sum() > 0
    assert (best >= 0).sum() > 5 and (best < 0).sum() > 5 and (best > 0).sum() > 0


def test_trim_fuzz_vs_ref(orc, ref):
    import qc_cases
    for seed in (5, 6):
        pool, contig, refv, cuts = qc_cases.trim_fuzz(seed, 400)
        a = orc.ref_trim_raw(pool, contig, refv, cuts)
        b = ref.ref_trim_raw(pool, contig, refv, cuts)
        assert a.tobytes() == b.tobytes()
        assert len(np.unique(b["adjLeadingCut"])) > 20


def test_pipeline_trim_select_vs_ref(orc, ref):
    """assemble -> trim -> align -> QC -> select on a synthetic batch: the restatement and the reference agree on
    every intermediate product, and the threaded compositions agree on the checksums"""
    import qc_cases
    batch = synth.make_candidate_batch(40, seed=31, window=1500)
    opts = cb.default_candidate_opts()
    cuts = cb.default_ref_cuts(batch.nLoci)
    cuts["leadingCut"], cuts["trailingCut"], cuts["maxLeadingCut"], cuts["maxTrailingCut"] = 20, 20, 250, 250
    co, to = qc_cases.pipeline_case(orc, batch, opts, cuts)
    cr, tr = qc_cases.pipeline_case(ref, batch, opts, cuts)
    assert to.tobytes() == tr.tobytes() and co.aln.tobytes() == cr.aln.tobytes() and co.jump.tobytes() == cr.jump.tobytes()
    assert co.cigar.tobytes() == cr.cigar.tobytes()
    got, want = co.run(orc), cr.run(ref)
    _same_qc(got, want)
    qc, best = want
    assert (best >= 0).sum() >= 10
    for with_cuts in (None, cuts):
        for sel in (False, True):
            assert orc.candidates_mt(opts, batch, batch.nLoci, 4, with_cuts, sel) == \
                ref.candidates_mt(opts, batch, batch.nLoci, 4, with_cuts, sel)
    # the flat checksum helper reproduces the threaded drivers' qcSum
    A = int(co.nContigs.max()) if len(co.nContigs) else 1
    slot = np.concatenate([np.arange(n) + li * A for li, n in enumerate(co.nContigs)]).astype(np.int64)
    qcA = np.zeros(batch.nLoci * A, dtype=cb.CONTIG_QC_DT)
    trA = np.zeros(batch.nLoci * A, dtype=cb.TRIM_RESULT_DT)
    qcA[slot], trA[slot] = qc, tr
    assert cb.qc_checksum(batch.isJump, co.nContigs, A, trA, qcA, best, True) == \
        ref.candidates_mt(opts, batch, batch.nLoci, 4, cuts, True)[3]


def _same_svs(a, b):
    (fa, sa), (fb, sb) = a, b
    assert fa.tolist() == fb.tolist() and len(sa) == len(sb)
    ok = sb["status"] == 0
    for k in sa.dtype.names:
        if k == "ciposBegin":  # where the reference throws, its left scan may run off the front of a string
            assert sa[k][ok].tolist() == sb[k][ok].tolist()
        else:
            assert sa[k].tolist() == sb[k].tolist(), k


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_small_sv_fuzz_vs_ref(orc, ref, seed):
    import qc_cases
    case = qc_cases.fuzz_case(seed, 250)
    qc, best = case.run(ref)
    got, want = case.small_svs(orc, qc, best), case.small_svs(ref, qc, best)
    _same_svs(got, want)
    svs = want[1]
    assert len(svs) > 20 and (svs["status"] == 0).sum() > 10 and (svs["segment"] > 0).sum() > 0
    assert (svs["ciposEnd"] > 0).sum() > 0


def test_small_sv_pipeline_vs_ref(orc, ref):
    import qc_cases
    batch = synth.make_candidate_batch(60, seed=41, window=1500)
    opts = cb.default_candidate_opts()
    case, _ = qc_cases.pipeline_case(ref, batch, opts, None)
    qc, best = case.run(ref)
    got, want = case.small_svs(orc, qc, best), case.small_svs(ref, qc, best)
    _same_svs(got, want)
    assert (want[1]["status"] == 0).sum() >= 10


@pytest.mark.parametrize("seed", [1, 2])
def test_jump_sv_fuzz_vs_ref(orc, ref, seed):
    """adjustAssembledBreakend over every combination of breakend state, alignment order and strand"""
    import qc_cases
    case = qc_cases.fuzz_case(seed, 400)
    ref1, ref2, orient = qc_cases.jump_sv_inputs(case, seed)
    qc, best = case.run(ref)
    a = orc.jump_sv_raw(case.isJump, case.slotFirst, len(case.jump), best, case.jump, case.cigar, ref1, ref2, orient)
    b = ref.jump_sv_raw(case.isJump, case.slotFirst, len(case.jump), best, case.jump, case.cigar, ref1, ref2, orient)
    assert a.tobytes() == b.tobytes()
    v = b[b["valid"] == 1]
    assert len(v) > 30
    combos = {(int(o["isBp2AlignedFirst"]), int(o["isBp1Reversed"]), int(o["isBp2Reversed"]), int(o["bp1RightOpen"]), int(o["bp2RightOpen"]))
              for o in orient[b["valid"] == 1]}
    assert len(combos) >= 20


def test_jump_qc_exotic_segments_vs_ref(orc, ref):
    """hard clips and reference skips in the jump arms (never emitted by the aligners, handled by the path helpers)"""
    import qc_cases
    case = qc_cases.fuzz_case(9, 300, exotic=True)
    _same_qc(case.run(orc), case.run(ref))
