"""Reference-window trim (row a5) and contig QC / bestAlignmentIndex selection (row a9) on the GPU against the
oracle: standalone entry points on fuzzed inputs, and the fused candidate path with both stages switched on."""
import numpy as np
import pytest

from manta_b200 import _cabi as cb
from manta_b200 import synth
import qc_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_select_contigs_fuzz(gpu, orc, seed):
    case = qc_cases.fuzz_case(seed, 400)
    (qg, bg), (qo, bo) = case.run(gpu), case.run(orc)
    assert bg.tolist() == bo.tolist()
    assert qg.tobytes() == qo.tobytes()


def test_select_contigs_other_options(gpu, orc):
    """non-default filter scores / minimum variant size"""
    case = qc_cases.fuzz_case(77, 300)
    for scores, minSize in [((1, -4, -6, -1, -1, 0), 5), ((2, -8, -18, 0, -1, 0), 30), ((3, -2, -2, -1, -1, 0), 1)]:
        o = cb.default_qc_opts()
        o["filterScores"] = scores
        o["minCandidateVariantSize"] = minSize
        (qg, bg), (qo, bo) = case.run(gpu, o), case.run(orc, o)
        assert bg.tolist() == bo.tolist() and qg.tobytes() == qo.tobytes()


def test_select_contigs_edge_cases(gpu, orc):
    """loci without contigs, a batch without slots, long paths (> the shared-memory staging size)"""
    case = qc_cases.fuzz_case(9, 50)
    case.nContigs[:] = 0
    qg, bg = case.run(gpu)
    assert bg.tolist() == [-1] * 50
    # one complex contig with > 600 path segments
    rng = np.random.default_rng(3)
    path = []
    for i in range(300):
        path += [(qc_cases.EQ, 40), (qc_cases.X, 1)]
        if i % 25 == 7:
            path += [(qc_cases.D if i % 2 else qc_cases.I, 12)]
    path += [(qc_cases.EQ, 60)]
    contig, ref = qc_cases._apply_path(rng, path, 17)
    pool = np.concatenate([contig, ref])
    aln = np.zeros(1, dtype=cb.ALN_RESULT_DT)
    aln[0] = (500, 17, 0, len(path), 0)
    cig = np.asarray([(ln << 4) | t for t, ln in path], dtype=np.uint32)
    c = qc_cases.QcCase(np.zeros(1, np.uint8), np.zeros(1, np.uint32), np.ones(1, np.uint32), aln,
                        np.zeros(1, dtype=cb.JUMP_RESULT_DT), cig, pool,
                        np.asarray([(0, len(contig), 0)], dtype=cb.SEQVIEW_DT),
                        np.asarray([(len(contig), len(ref), 0)], dtype=cb.SEQVIEW_DT), np.asarray([5], np.uint32))
    (qg, bg), (qo, bo) = c.run(gpu), c.run(orc)
    assert qg.tobytes() == qo.tobytes() and bg.tolist() == bo.tolist() and qo["nSegments"][0] >= 2


def test_select_contigs_bad_args(gpu):
    case = qc_cases.fuzz_case(5, 20)
    case.aln["cigarOff"][:] = 10 ** 9
    if (case.isJump == 0).any() and case.nContigs[case.isJump == 0].sum() > 0:
        with pytest.raises(cb.Mb200Error) as e:
            case.run(gpu)
        assert e.value.status == cb.MB200_ERR_BAD_ARG


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_ref_trim_fuzz(gpu, orc, seed):
    pool, contig, refv, cuts = qc_cases.trim_fuzz(seed, 600)
    assert gpu.ref_trim_raw(pool, contig, refv, cuts).tobytes() == orc.ref_trim_raw(pool, contig, refv, cuts).tobytes()


def test_ref_trim_long_contig_and_errors(gpu, orc):
    rng = np.random.default_rng(1)
    ref = qc_cases.ACGT[rng.integers(0, 4, 9000)]
    contig = ref[2000:7000].copy()   # longer than the shared-memory table: per-warp global table
    pool = np.concatenate([contig, ref])
    cv = np.asarray([(0, len(contig), 0)], dtype=cb.SEQVIEW_DT)
    rv = np.asarray([(len(contig), len(ref), 0)], dtype=cb.SEQVIEW_DT)
    cuts = np.zeros(1, dtype=cb.REF_CUTS_DT)
    cuts[0] = (1990, 1990, 3000, 3000)   # start close to the contig: no chance 10-mer hit before it
    got = gpu.ref_trim_raw(pool, cv, rv, cuts)
    assert got.tobytes() == orc.ref_trim_raw(pool, cv, rv, cuts).tobytes()
    assert (int(got["adjLeadingCut"][0]), int(got["adjTrailingCut"][0])) == (2000, 2000)
    cv["len"][0] = 9
    with pytest.raises(cb.Mb200Error) as e:
        gpu.ref_trim_raw(pool, cv, rv, cuts)
    assert e.value.status == cb.MB200_ERR_BAD_ARG


def _fused(gpu, batch, opts, cuts, select):
    gpu.candidates_upload(batch, cuts)
    return gpu.candidates_run(opts, select=select)


def test_fused_trim_select_vs_oracle(gpu, orc):
    """assemble -> trim -> align -> QC -> select on the device == the same composition of oracle calls"""
    batch = synth.make_candidate_batch(64, seed=31, window=1500)
    opts = cb.default_candidate_opts()
    cuts = cb.default_ref_cuts(batch.nLoci)
    cuts["leadingCut"], cuts["trailingCut"], cuts["maxLeadingCut"], cuts["maxTrailingCut"] = 20, 20, 250, 250
    A = int(opts["asmOpts"]["maxAssemblyCount"][0])
    case, trims = qc_cases.pipeline_case(orc, batch, opts, cuts)
    qo, bo = case.run(orc)
    out = _fused(gpu, batch, opts, cuts, True)
    assert out["loci"]["nContigs"].tolist() == case.nContigs.tolist()
    slot = np.concatenate([np.arange(n) + li * A for li, n in enumerate(case.nContigs)]).astype(np.int64)
    assert out["best"].tolist() == bo.tolist()
    assert out["qc"][slot].tobytes() == qo.tobytes()
    assert out["trim"][slot].tobytes() == trims.tobytes()
    locusOf = slot // A
    cplx = batch.isJump[locusOf] == 0
    for k in ("score", "beginPos", "isJumped", "cigarLen"):
        assert out["aln"][k][slot][cplx].tolist() == case.aln[k][cplx].tolist(), k
    # the same CIGARs
    for s, a in zip(slot[cplx], case.aln[cplx]):
        g = out["aln"][s]
        assert out["cigar"][int(g["cigarOff"]):int(g["cigarOff"]) + int(g["cigarLen"])].tolist() == \
            case.cigar[int(a["cigarOff"]):int(a["cigarOff"]) + int(a["cigarLen"])].tolist()
    assert (bo >= 0).sum() > 10
    t = gpu.candidates_timing()
    assert t["trimMs"] > 0 and t["qcMs"] > 0
    # selection without the trim, and the trim without selection, are independent switches
    out2 = _fused(gpu, batch, opts, None, True)
    case2, _ = qc_cases.pipeline_case(orc, batch, opts, None)
    q2, b2 = case2.run(orc)
    assert out2["best"].tolist() == b2.tolist() and out2["qc"][slot].tobytes() == q2.tobytes()
    assert not out2["trim"]["adjLeadingCut"].any()
    out3 = _fused(gpu, batch, opts, cuts, False)
    assert out3["aln"]["beginPos"][slot][cplx].tolist() == case.aln["beginPos"][cplx].tolist()


def test_fused_trim_select_checksum_full_window(gpu, orc):
    """BASELINE config-2 sized loci with the refiner's default cuts: alignment-score and QC checksums of the whole
    batch equal the multi-threaded CPU composition"""
    batch = synth.make_candidate_batch(160, seed=99)
    opts = cb.default_candidate_opts()
    cuts = cb.default_ref_cuts(batch.nLoci)
    out = _fused(gpu, batch, opts, cuts, True)
    A = int(opts["asmOpts"]["maxAssemblyCount"][0])
    nC = out["loci"]["nContigs"]
    live = np.arange(A)[None, :] < nC[:, None]
    isJ = (batch.isJump != 0)[:, None]
    score = np.where(isJ, out["jump"]["score"].reshape(-1, A), out["aln"]["score"].reshape(-1, A))
    nc, cells, ssum, qsum = orc.candidates_mt(opts, batch, batch.nLoci, 8, cuts, True)
    assert (int(live.sum()), int(score[live].sum())) == (nc, ssum)
    assert cb.qc_checksum(batch.isJump, nC, A, out["trim"], out["qc"], out["best"], True) == qsum
    t = gpu.candidates_timing()
    assert t["indelCells"] + t["jumpCells"] == cells
    assert (out["best"] >= 0).sum() > 40


def test_refiner_qc_golden(gpu):
    """the reference's own unit-test answers and the committed reference outputs, through the CUDA path"""
    from test_oracle_golden import check_refiner_golden
    check_refiner_golden(gpu)


def _svs_equal(a, b):
    (fa, sa), (fb, sb) = a, b
    assert fa.tolist() == fb.tolist()
    assert sa.tobytes() == sb.tobytes()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_small_sv_fuzz(gpu, orc, seed):
    case = qc_cases.fuzz_case(seed, 400)
    qc, best = case.run(orc)
    _svs_equal(case.small_svs(gpu, qc, best), case.small_svs(orc, qc, best))
    # another minimum size: different candidate runs, same bookkeeping
    o = cb.default_qc_opts()
    o["minCandidateVariantSize"] = 25
    qc, best = case.run(orc, o)
    _svs_equal(case.small_svs(gpu, qc, best, 25), case.small_svs(orc, qc, best, 25))


def test_small_sv_golden(gpu):
    from test_oracle_golden import check_small_sv_golden
    check_small_sv_golden(gpu)


def test_fused_small_svs(gpu, orc):
    """refined small SVs straight from the resident results of a fused run == the oracle on the same pipeline"""
    batch = synth.make_candidate_batch(64, seed=77, window=1500)
    opts = cb.default_candidate_opts()
    cuts = cb.default_ref_cuts(batch.nLoci)
    cuts["leadingCut"], cuts["trailingCut"], cuts["maxLeadingCut"], cuts["maxTrailingCut"] = 20, 20, 250, 250
    case, _ = qc_cases.pipeline_case(orc, batch, opts, cuts)
    qo, bo = case.run(orc)
    first_o, svs_o = case.small_svs(orc, qo, bo)
    _fused(gpu, batch, opts, cuts, True)
    first_g, svs_g = gpu.candidates_small_svs()
    assert first_g.tolist() == first_o.tolist() and svs_g.tobytes() == svs_o.tobytes()
    assert (svs_o["status"] == 0).sum() >= 10
    # a run without selection leaves nothing to refine
    _fused(gpu, batch, opts, cuts, False)
    with pytest.raises(cb.Mb200Error):
        gpu.candidates_small_svs()


def test_fused_select_cigar_overflow_and_empty_loci(gpu, orc):
    """a CIGAR pool that is too small: the QC kernels must stand down (the pool is incomplete), the call reports the
    overflow with the size needed, and the retry is exact; loci without reads produce no contigs and best == -1"""
    import ctypes as C
    batch = synth.make_candidate_batch(24, seed=13, window=1500)
    # empty the reads of two loci (one complex, one spanning if present)
    ro = batch.readOff.copy()
    for li in (3, 4):
        n = int(ro[li + 1] - ro[li])
        ro[li + 1:] -= n
    keep = np.ones(len(batch.readViews), dtype=bool)
    for li in (3, 4):
        keep[int(batch.readOff[li]):int(batch.readOff[li + 1])] = False
    batch.readViews = batch.readViews[keep].copy()
    batch.readOff = ro
    opts = cb.default_candidate_opts()
    gpu.candidates_upload(batch)
    out = gpu.candidates_run(opts, select=True)           # sizes everything
    assert out["best"][3] == -1 and out["best"][4] == -1 and out["loci"]["nContigs"][3] == 0
    want_best, want_qc = out["best"].copy(), out["qc"].copy()
    small = dict(out)
    small["cigar"] = np.zeros(8, dtype=np.uint32)          # far too small
    used = [C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    st = gpu._call("candidates_run", p(opts), p(small["loci"]), p(small["contigs"]), p(small["pseudo"]), p(small["contigPool"]),
                   C.c_uint64(small["contigPool"].size), C.byref(used[0]), p(small["setPool"]), C.c_uint64(small["setPool"].size),
                   C.byref(used[1]), p(small["aln"]), p(small["jump"]), p(small["cigar"]), C.c_uint64(8), C.byref(used[2]),
                   p(small["trim"]), p(small["qc"]), p(small["best"]))
    assert st == cb.MB200_ERR_POOL_OVERFLOW and used[2].value > 8
    out2 = gpu.candidates_run(opts, select=True)
    assert out2["best"].tolist() == want_best.tolist() and out2["qc"].tobytes() == want_qc.tobytes()
    case, _ = qc_cases.pipeline_case(orc, batch, opts, None)
    assert case.run(orc)[1].tolist() == want_best.tolist()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_jump_sv_fuzz(gpu, orc, seed):
    case = qc_cases.fuzz_case(seed, 500)
    ref1, ref2, orient = qc_cases.jump_sv_inputs(case, seed)
    qc, best = case.run(orc)
    args = (case.isJump, case.slotFirst, len(case.jump), best, case.jump, case.cigar, ref1, ref2, orient)
    assert gpu.jump_sv_raw(*args).tobytes() == orc.jump_sv_raw(*args).tobytes()


def test_jump_sv_golden(gpu):
    from test_oracle_golden import check_jump_sv_golden
    check_jump_sv_golden(gpu)


def test_fused_jump_svs(gpu, orc):
    batch = synth.make_candidate_batch(64, seed=78, window=1500)
    opts = cb.default_candidate_opts()
    case, _ = qc_cases.pipeline_case(orc, batch, opts, None)
    qo, bo = case.run(orc)
    rng = np.random.default_rng(4)
    orient = np.zeros(batch.nLoci, dtype=cb.JUMP_ORIENT_DT)
    for k in ("isBp2AlignedFirst", "isBp1Reversed", "isBp2Reversed", "bp1RightOpen", "bp2RightOpen"):
        orient[k] = rng.integers(0, 2, batch.nLoci)
    want = orc.jump_sv_raw(case.isJump, case.slotFirst, len(case.jump), bo, case.jump, case.cigar, batch.ref1, batch.ref2, orient)
    _fused(gpu, batch, opts, None, True)
    got = gpu.candidates_jump_svs(orient)
    assert got.tobytes() == want.tobytes() and int(want["valid"].sum()) >= 5
