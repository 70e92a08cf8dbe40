"""Golden vectors on real reads: every assembler / aligner call and a seventh of the split-read alignments that the
reference's own GenerateSVCandidates makes on its demo (tests/golden/demo_calls.json.gz, recorded by
tests/golden/make_demo_calls.py through oracle/trace_calls.patch -- inputs and the unmodified reference's results).
CPU: the oracle port reproduces them; GPU: so does the CUDA path through the C ABI (assembler: both kernels)."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

from manta_b200 import _cabi as cb
from test_oracle_golden import GOLD, asm_to_json, check_split, run_alignment


@pytest.fixture(scope="module")
def calls():
    with gzip.open(os.path.join(GOLD, "demo_calls.json.gz"), "rt") as f:
        g = json.load(f)
    by = {}
    for c in g["cases"]:
        by.setdefault(c["kind"], []).append(c)
    assert len(by["asm"]) == 5 and len(by["jump"]) == 3 and len(by["indel"]) == 3 and len(by["split"]) > 900
    return g, by


def check_assembly(binding, by):
    for i, c in enumerate(by["asm"]):
        got = asm_to_json(binding.assemble(cb.default_asm_opts(**c["opts"]), [c["reads"]])[0])
        assert got == c["reference"], i


def check_alignments(binding, by):
    for kind in ("jump", "indel"):
        for i, c in enumerate(by[kind]):
            assert run_alignment(binding, c) == c["reference"], (kind, i)


def check_splits(binding, g, by, orc):
    # the reference's snp prior is a FLOAT (1e-3f, options/CallOptionsShared.hpp:30): qscore_snp sees 0.0010000000474974513.
    # Tables from the oracle's qscore_snp restatement (pinned to the reference's in test_qscore_tables_golden).
    assert g["snpPriors"] == [float(np.float32(1e-3)), 0.0]
    tables = np.zeros(1, dtype=cb.QSCORE_TABLES_DT)
    orc.lib.orc_qscore_tables(C.c_double(g["snpPriors"][0]), C.c_double(g["snpPriors"][1]), tables.ctypes.data_as(C.c_void_p))
    cases = [dict(c, test="demo split %d" % i) for i, c in enumerate(by["split"])]
    flanks = sorted({c["flank"] for c in cases})
    assert flanks == [50]
    check_split(binding, tables, dict(cases=cases, snpPriors=g["snpPriors"]))


def test_oracle_assembly_on_demo_reads(orc, calls):
    check_assembly(orc, calls[1])


def test_oracle_alignments_on_demo_contigs(orc, calls):
    check_alignments(orc, calls[1])


def test_oracle_split_alignments_on_demo_reads(orc, calls):
    check_splits(orc, *calls, orc)


@pytest.mark.gpu
def test_gpu_assembly_on_demo_reads(gpuasm, calls):
    check_assembly(gpuasm, calls[1])


@pytest.mark.gpu
def test_gpu_alignments_on_demo_contigs(gpu, gpu32, gpusel, calls):
    for ctx in (gpu, gpu32, gpusel):
        check_alignments(ctx, calls[1])


@pytest.mark.gpu
def test_gpu_split_alignments_on_demo_reads(gpu, orc, calls):
    check_splits(gpu, *calls, orc)


def test_reference_inputs_stay_inside_the_library_limits(calls):
    """DESIGN.md 3.3 lists what the library rejects loudly although the reference would accept it (bases outside ACGTN,
    contigs beyond cfg.maxContigLen = 2048, qualities above 70).  On BAM-decoded data the reference never produces such
    input -- checked here on everything its own executable fed to the hot path on the demo."""
    _, by = calls
    for c in by["asm"]:
        for r in c["reads"]:
            assert set(r) <= set("ACGTN")
        for ctg in c["reference"]["contigs"]:
            assert len(ctg["seq"]) <= 2048 and set(ctg["seq"]) <= set("ACGT")
    for kind in ("jump", "indel"):
        for c in by[kind]:
            for seq in c["seqs"]:
                assert set(seq) <= set("ACGTN")
    for c in by["split"]:
        assert set(c["query"]) <= set("ACGTN") and set(c["target"]) <= set("ACGTN")
        assert max(c["qual"]) <= 70
