"""The C restatement (oracle/) against the committed golden vectors (tests/golden/*.json):
values asserted by the reference's own unit tests plus full results recorded from the reference build."""
import ctypes as C
import json
import os
from dataclasses import asdict

import numpy as np
import pytest

from manta_b200 import _cabi as cb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def run_alignment(b, case):
    S = cb.Scores(*case["scores"])
    if case["kind"] == "jump":
        return asdict(b.jump_align(S, case["extra"], [tuple(case["seqs"])])[0])
    return asdict(b.single_align(case["kind"], S, case["extra"], [tuple(case["seqs"])])[0])


def lookup(d, dotted):
    for part in dotted.split("."):
        d = d[part]
    return int(d) if isinstance(d, bool) else d


def test_alignment_golden(orc):
    cases = load("alignment.json")["cases"]
    assert len([c for c in cases if c["expected"]]) >= 40
    for c in cases:
        got = run_alignment(orc, c)
        for k, v in c["expected"].items():
            assert lookup(got, k) == v, (c["test"], k)
        assert got == c["reference"], c["test"]


def asm_to_json(res):
    contigs, pseudo, info = res
    return dict(contigs=[json.loads(json.dumps(asdict(c))) for c in contigs], pseudoReads=pseudo,
                readInfo=[dict(isUsed=i.isUsed, isPseudo=i.isPseudo, contigIds=i.contigIds) for i in info])


def test_assembly_golden(orc):
    for c in load("assembly.json")["cases"]:
        got = asm_to_json(orc.assemble(cb.default_asm_opts(**c["opts"]), [c["reads"]])[0])
        if c["expected"]:
            assert [x["seq"] for x in got["contigs"]] == c["expected"]["contigs"], c["test"]
        assert got == c["reference"], c["test"]


def test_qscore_tables_golden(orc):
    g = load("splitread.json")
    t = np.zeros(1, dtype=cb.QSCORE_TABLES_DT)
    orc.lib.orc_qscore_tables(C.c_double(g["snpPriors"][0]), C.c_double(g["snpPriors"][1]), t.ctypes.data_as(C.c_void_p))
    for name in ("lnComp", "lnErr"):
        want = np.array([[float.fromhex(x) for x in row] for row in g[name]])
        assert t[name][0].tobytes() == want.tobytes(), name


def split_jobs(cases):
    sp = cb.SeqPool()
    jobs = np.zeros(len(cases), dtype=cb.SR_JOB_DT)
    for i, c in enumerate(cases):
        jobs["q_off"][i], jobs["q_len"][i] = sp.add(c["query"])
        jobs["qualOff"][i], _ = sp.add(bytes(c["qual"]))
        jobs["t_off"][i], jobs["t_len"][i] = sp.add(c["target"])
        jobs["bpBegin"][i], jobs["bpEnd"][i] = c["bp"]
        jobs["table"][i] = c["table"]
    return sp.finish(), jobs


def check_split(binding, tables, g):
    cases = g["cases"]
    pool, jobs = split_jobs(cases)
    res = {f: binding.split_align_raw(tables, tuple(g["snpPriors"]), f, pool, jobs) for f in (2, 50)}
    for i, c in enumerate(cases):
        r = res[c["flank"]][i]
        want = c["reference"]
        for k in r.dtype.names:
            if k == "reserved":
                continue
            if k in ("alignLnLhood", "evidence"):
                assert int(np.float32(r[k]).view(np.uint32)) == want[k + "_bits"], (c["test"], k)
            else:
                assert int(r[k]) == want[k], (c["test"], k)


def golden_tables(g):
    t = np.zeros(1, dtype=cb.QSCORE_TABLES_DT)
    for name in ("lnComp", "lnErr"):
        t[name][0] = np.array([[float.fromhex(x) for x in row] for row in g[name]])
    return t


def test_splitread_golden(orc):
    g = load("splitread.json")
    check_split(orc, golden_tables(g), g)


def test_unordered_order_golden(orc):
    """libstdc++ std::unordered_map<std::string,T> iteration order"""
    for c in load("unordered_order.json")["cases"]:
        n = len(c["keys"])
        buf = np.frombuffer("".join(c["keys"]).encode() + b"\0", dtype=np.uint8).copy()
        order = np.zeros(max(n, 1), dtype=np.uint32)
        orc.lib.orc_unordered_order(C.c_uint32(n), buf.ctypes.data_as(C.c_void_p), C.c_uint32(c["k"]),
                                    order.ctypes.data_as(C.c_void_p))
        assert [int(x) for x in order[:n]] == c["order"]


def check_refiner_golden(binding):
    import qc_cases
    g = load("refiner_qc.json")
    for t in g["unit"]:
        case, o = qc_cases.unit_case(t, g["scores"])
        qc, best = case.run(binding, o)
        for k in ("isCandidate", "isLowQuality"):
            if k in t["expected"]:
                assert qc[k].tolist() == t["expected"][k], (t["src"], k)
        if "best" in t["expected"]:
            assert int(best[0]) == t["expected"]["best"], t["src"]
        assert [[int(x) for x in row] for row in qc.tolist()] == t["reference"]["qc"], t["src"]
        assert int(best[0]) == t["reference"]["best"], t["src"]
    for f in g["fuzz"]:
        qc, best = qc_cases.fuzz_case(f["seed"], f["n_loci"]).run(binding)
        assert [[int(x) for x in row] for row in qc.tolist()] == f["qc"] and best.tolist() == f["best"]
    for f in g["trim"]:
        pool, contig, refv, cuts = qc_cases.trim_fuzz(f["seed"], f["n"])
        tr = binding.ref_trim_raw(pool, contig, refv, cuts)
        assert [[int(a), int(b)] for a, b in tr.tolist()] == f["trim"]


def check_small_sv_golden(binding):
    import qc_cases
    from manta_b200 import _cabi as cb
    g = load("refiner_qc.json")
    for t in g["small_sv_unit"]:
        u = dict(kind="complex", minSize=t["minSize"], contig=t["contig"], ref=t["ref"],
                 contigs=[dict(cigar=t["cigar"], beginPos=t["beginPos"], score=1, isJumped=0, support=1)])
        case, _ = qc_cases.unit_case(u, g["scores"])
        qc = np.zeros(1, dtype=cb.CONTIG_QC_DT)
        qc[0]["isCandidate"], qc[0]["nSegments"], qc[0]["pathFirst"], qc[0]["pathLast"] = 1, 1, 1, 1
        first, svs = case.small_svs(binding, qc, np.zeros(1, dtype=np.int32), t["minSize"])
        assert first.tolist() == [0, 1] and [int(x) for x in svs[0].tolist()] == t["reference"], t["src"]
        exp = t["expected"]
        if "cipos" in exp:
            assert [int(svs[0]["ciposBegin"]), int(svs[0]["ciposEnd"])] == exp["cipos"], t["src"]
        if "bp1" in exp:
            assert [int(svs[0]["bp1Begin"]), int(svs[0]["bp1End"])] == exp["bp1"], t["src"]
            assert [int(svs[0]["bp2Begin"]), int(svs[0]["bp2End"])] == exp["bp2"], t["src"]
    for f in g["small_sv_fuzz"]:
        case = qc_cases.fuzz_case(f["seed"], f["n_loci"])
        qc, best = case.run(binding)
        first, svs = case.small_svs(binding, qc, best)
        assert [int(x) for x in first] == f["svFirst"]
        for got, want in zip(svs.tolist(), f["svs"]):
            assert all(w is None or int(a) == w for a, w in zip(got, want))


def check_jump_sv_golden(binding):
    import re
    import qc_cases
    from manta_b200 import _cabi as cb
    g = load("refiner_qc.json")
    code = {c: i for i, c in enumerate(cb.SEG_CODES)}
    segs = lambda cg: [(int(n) << 4) | code[c] for n, c in re.findall(r"(\d+)([MIDNSHP=X])", cg)]  # noqa: E731
    for t in g["jump_sv_unit"]:
        s1, s2 = segs(t["cigar1"]), segs(t["cigar2"])
        jump = np.zeros(1, dtype=cb.JUMP_RESULT_DT)
        jump[0] = (100, 0, t["jumpRange"], t["begin1"], t["begin2"], len(s1), len(s2), 0, 0, len(s1))
        orient = np.zeros(1, dtype=cb.JUMP_ORIENT_DT)
        for k, v in t["orient"].items():
            orient[k] = v
        got = binding.jump_sv_raw(np.ones(1, np.uint8), np.zeros(1, np.uint32), 1, np.zeros(1, np.int32), jump,
                                  np.asarray(s1 + s2, dtype=np.uint32), np.asarray([(0, t["ref1Len"], 0)], dtype=cb.SEQVIEW_DT),
                                  np.asarray([(0, t["ref2Len"], 0)], dtype=cb.SEQVIEW_DT), orient)[0]
        assert [int(got["bp1Begin"]), int(got["bp1End"])] == t["expected"]["bp1"], t["src"]
        assert [int(got["bp2Begin"]), int(got["bp2End"])] == t["expected"]["bp2"], t["src"]
        assert [int(x) for x in got.tolist()] == t["reference"], t["src"]
    for f in g["jump_sv_fuzz"]:
        case = qc_cases.fuzz_case(f["seed"], f["n_loci"])
        ref1, ref2, orient = qc_cases.jump_sv_inputs(case, f["seed"])
        qc, best = case.run(binding)
        got = binding.jump_sv_raw(case.isJump, case.slotFirst, len(case.jump), best, case.jump, case.cigar, ref1, ref2, orient)
        assert [[int(x) for x in row] for row in got.tolist()] == f["svs"]


def test_jump_sv_golden(orc):
    """test_generateRefinedSVCandidateFromJumpAlignment known answers + reference outputs on seeded inputs"""
    check_jump_sv_golden(orc)


def test_small_sv_golden(orc):
    """test_GetVariantRange / test_setSmallCandSV known answers + reference outputs on seeded inputs"""
    check_small_sv_golden(orc)


def test_refiner_qc_golden(orc):
    """GSC/test/SVCandidateAssemblyRefinerTest.cpp known answers + reference outputs on seeded inputs"""
    check_refiner_golden(orc)
