"""Host-side checks that need no GPU: the product library loads and exports every symbol the header declares,
fails loudly without a device, and the ctypes mirror matches the C struct sizes."""
import ctypes as C
import os
import re
import subprocess

import pytest

import manta_b200
from manta_b200 import _cabi as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(manta_b200.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(manta_b200.library_path())


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "manta_b200.h")).read()
    return sorted(set(re.findall(r"MB200_API\s+[\w\s\*]*?(mb200_\w+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "libmanta_b200.so does not export %s" % s


def test_struct_sizes_match_header(tmp_path):
    prog = tmp_path / "sizes.c"
    names = ["mb200_scores", "mb200_seqview", "mb200_aln_result", "mb200_jump_result", "mb200_sr_result",
             "mb200_sr_job", "mb200_qscore_tables", "mb200_asm_opts", "mb200_contig", "mb200_asm_locus",
             "mb200_config", "mb200_timing", "mb200_candidate_opts", "mb200_candidate_timing", "mb200_ref_cuts",
             "mb200_trim_result", "mb200_qc_opts", "mb200_contig_qc", "mb200_small_sv", "mb200_jump_orient", "mb200_jump_sv"]
    prog.write_text('#include <stdio.h>\n#include "manta_b200.h"\nint main(){' +
                    "".join('printf("%%zu\\n", sizeof(%s));' % n for n in names) + "return 0;}")
    exe = tmp_path / "sizes"
    # the header is plain C: strict C99 must accept it
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(cb.Scores), cb.SEQVIEW_DT.itemsize, cb.ALN_RESULT_DT.itemsize, cb.JUMP_RESULT_DT.itemsize,
            cb.SR_RESULT_DT.itemsize, cb.SR_JOB_DT.itemsize, cb.QSCORE_TABLES_DT.itemsize, cb.ASM_OPTS_DT.itemsize,
            cb.CONTIG_DT.itemsize, cb.ASM_LOCUS_DT.itemsize, cb.CONFIG_DT.itemsize, cb.TIMING_DT.itemsize,
            cb.CAND_OPTS_DT.itemsize, cb.CAND_TIMING_DT.itemsize, cb.REF_CUTS_DT.itemsize, cb.TRIM_RESULT_DT.itemsize,
            cb.QC_OPTS_DT.itemsize, cb.CONTIG_QC_DT.itemsize, cb.SMALL_SV_DT.itemsize, cb.JUMP_ORIENT_DT.itemsize,
            cb.JUMP_SV_DT.itemsize]
    assert got == want


def test_no_cpu_fallback(lib):
    """without a CUDA device the product refuses to run instead of silently computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cb.Mb200Error) as e:
        manta_b200.Context(0)
    assert e.value.status == cb.MB200_ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    """the oracle is test infrastructure: nothing under manta_b200/ may import, link or load it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "manta_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libmanta_oracle" not in text and "libmanta_ref" not in text and "manta_oracle.h" not in text, f
    out = subprocess.check_output(["ldd", manta_b200.library_path()]).decode()
    assert "oracle" not in out and "manta_ref" not in out
