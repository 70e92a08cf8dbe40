"""manta_b200 -- B200-native implementation of Illumina/manta's GenerateSVCandidates hot path.

The product is ``libmanta_b200.so`` (hand-written sm_100a CUDA behind the C ABI of include/manta_b200.h).
This package is the thin host-side mirror of the reference interfaces on that path, used by the parity
tests and the bench:

    GlobalJumpAligner(scores, jumpScore).align(query, ref1, ref2)   lib/alignment/GlobalJumpAligner.hpp:45-53
    GlobalLargeIndelAligner(scores, largeIndelScore).align(q, ref)  lib/alignment/GlobalLargeIndelAligner.hpp:48-54
    GlobalAligner(scores).align(q, ref)                             lib/alignment/GlobalAligner.hpp
    splitReadAligner(...)                                           GSC/SplitReadAlignment.hpp:63-70
    runIterativeAssembler(opt, reads) -> (contigs, reads, readInfo) lib/assembly/IterativeAssembler.hpp:43-47

There is NO CPU fallback: importing works anywhere, but creating a Context raises unless the CUDA library is
built and a compute-capability-10.x device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from ._cabi import (AlignmentResult, AssembledContig, AssemblyReadInfo, Binding, JumpAlignmentResult, Mb200Error,
                    SRAlignmentInfo, Scores, default_asm_opts)

__all__ = ["Context", "AlignmentScores", "GlobalAligner", "GlobalJumpAligner", "GlobalLargeIndelAligner",
           "IterativeAssemblerOptions", "runIterativeAssembler", "splitReadAligner", "qscore_snp", "Mb200Error",
           "library_path"]

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    # MANTA_B200_LIBRARY: an instrumented build of the same sources (tools/asm_profile.py), never a different backend
    return os.environ.get("MANTA_B200_LIBRARY") or os.path.join(_HERE, "libmanta_b200.so")


def _load_library() -> C.CDLL:
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("manta_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.mb200_create.restype = C.c_int
    lib.mb200_last_error.restype = C.c_char_p
    lib.mb200_version.restype = C.c_char_p
    return lib


class Context(Binding):
    """One device context (mb200_ctx): owns the stream and all device buffers."""

    def __init__(self, device: int = 0, ctas_per_sm: int = 0, scratch_bytes: int = 0, force_int32: bool = False,
                 max_contig_len: int = 0, no_tag_kernel: bool = False, assembler: str = "auto"):
        lib = _load_library()
        cfg = np.zeros(1, dtype=_cabi.CONFIG_DT)
        cfg["ctasPerSm"] = ctas_per_sm
        cfg["maxContigLen"] = max_contig_len  # assembler workspace cap per contig (0 = library default 2048)
        cfg["scratchBytes"] = scratch_bytes
        # bit 0: always use the int32 one-alignment-per-warp DP kernel; bit 1: int16x2 pairs on the predicate/SEL
        # kernel, never the tagged-score kernel (the default where its score range allows)
        # bits 2/3: assembler kernel -- "global" (window-indexed tables in HBM, 32 loci per SM: throughput), "onchip" (dense
        # tables in shared memory, 3 loci per SM: latency); "auto" takes the on-chip kernel for batches of up to 3 loci per SM
        if assembler not in ("auto", "global", "onchip"):
            raise ValueError("assembler must be 'auto', 'global' or 'onchip'")
        cfg["reserved"][0][0] = ((1 if force_int32 else 0) | (2 if no_tag_kernel else 0) |
                                 (4 if assembler == "global" else 0) | (8 if assembler == "onchip" else 0))
        handle = C.c_void_p()
        st = lib.mb200_create(C.c_int(device), cfg.ctypes.data_as(C.c_void_p), C.byref(handle))
        if st != _cabi.MB200_OK:
            raise Mb200Error(st, "mb200_create(device=%d) failed with status %d: no usable sm_100 CUDA device "
                                 "(manta_b200 has no CPU fallback)" % (device, st))
        super().__init__(lib, "mb200_", handle)
        self.device = device

    def close(self):
        if self.ctx is not None:
            self.lib.mb200_destroy.restype = None
            self.lib.mb200_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def timing(self) -> dict:
        t = np.zeros(1, dtype=_cabi.TIMING_DT)
        self.lib.mb200_get_timing(self.ctx, t.ctypes.data_as(C.c_void_p))
        return {k: (float(t[k][0]) if t[k].dtype.kind == "f" else int(t[k][0])) for k in t.dtype.names}

    def version(self) -> str:
        return self.lib.mb200_version().decode()


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---- reference-shaped front ends ---------------------------------------------------------------
def AlignmentScores(match, mismatch, open, extend, offEdge, isAllowEdgeInsertion=False) -> Scores:
    """AlignmentScores<int>  lib/alignment/AlignmentScores.hpp:23-57"""
    return Scores(match, mismatch, open, extend, offEdge, 1 if isAllowEdgeInsertion else 0)


class GlobalJumpAligner:
    def __init__(self, scores: Scores, jumpScore: int, ctx: Optional[Context] = None):
        if scores.isAllowEdgeInsertion:
            raise ValueError("unsupported option: isAllowEdgeInsertion (GlobalJumpAligner.hpp:41)")
        self.scores, self.jumpScore, self.ctx = scores, jumpScore, ctx

    def align(self, query, ref1, ref2) -> JumpAlignmentResult:
        return self.align_batch([(query, ref1, ref2)])[0]

    def align_batch(self, triples: Sequence[Tuple]) -> List[JumpAlignmentResult]:
        return (self.ctx or default_context()).jump_align(self.scores, self.jumpScore, triples)


class GlobalLargeIndelAligner:
    def __init__(self, scores: Scores, largeIndelScore: int, ctx: Optional[Context] = None):
        self.scores, self.largeIndelScore, self.ctx = scores, largeIndelScore, ctx

    def align(self, query, ref) -> AlignmentResult:
        return self.align_batch([(query, ref)])[0]

    def align_batch(self, pairs: Sequence[Tuple]) -> List[AlignmentResult]:
        return (self.ctx or default_context()).single_align("indel", self.scores, self.largeIndelScore, pairs)


class GlobalAligner:
    def __init__(self, scores: Scores, ctx: Optional[Context] = None):
        self.scores, self.ctx = scores, ctx

    def align(self, query, ref) -> AlignmentResult:
        return self.align_batch([(query, ref)])[0]

    def align_batch(self, pairs: Sequence[Tuple]) -> List[AlignmentResult]:
        return (self.ctx or default_context()).single_align("global", self.scores, 0, pairs)


def IterativeAssemblerOptions(**kw) -> np.ndarray:
    """IterativeAssemblerOptions  lib/options/IterativeAssemblerOptions.hpp:26-59 (defaults identical)"""
    return default_asm_opts(**kw)


def runIterativeAssembler(opt: np.ndarray, reads: Sequence[str], ctx: Optional[Context] = None):
    """returns (contigs, reads incl. appended pseudo reads, readInfo) -- the three in/out arguments of the
    reference call (lib/assembly/IterativeAssembler.cpp:844-931)"""
    contigs, pseudo, info = (ctx or default_context()).assemble(opt, [list(reads)])[0]
    return contigs, list(reads) + pseudo, info


class qscore_snp:
    """qscore_snp(snp_prob)  lib/blt_util/qscore_snp.cpp:29-41 -- the two per-phred log tables (host libm)"""

    def __init__(self, snp_prob: float):
        import math
        self.snp_prob = snp_prob
        comp_snp3 = 1.0 - (snp_prob / 3.0)
        self.lnComp, self.lnErr = [], []
        for i in range(_cabi.MAX_QSCORE + 1):
            qerr = math.pow(10.0, -float(i) / 10.0)
            q2p = (qerr * comp_snp3) + ((1 - qerr) * snp_prob)
            x = -q2p
            # C semantics: log(0) == -inf (phred 0 with snp_prob 0)
            self.lnComp.append(math.log1p(x) if abs(x) < 0.01 else (math.log(1 + x) if 1 + x > 0 else -math.inf))
            self.lnErr.append(math.log(q2p) if q2p > 0 else -math.inf)


#: The reference builds its reference-allele table from CallOptionsShared::snpPrior, a FLOAT (options/CallOptionsShared.hpp:30,
#: SVScorer.hpp:47): qscore_snp(REFERENCE_SNP_PRIOR), not qscore_snp(0.001), reproduces GenerateSVCandidates' split-read
#: likelihoods to the bit (tests/test_demo_calls.py); the alt-allele table is qscore_snp(0).
REFERENCE_SNP_PRIOR = float(np.float32(1e-3))


def qscore_tables(t0: "qscore_snp", t1: "qscore_snp") -> np.ndarray:
    t = np.zeros(1, dtype=_cabi.QSCORE_TABLES_DT)
    t["lnComp"][0][0], t["lnComp"][0][1] = t0.lnComp, t1.lnComp
    t["lnErr"][0][0], t["lnErr"][0][1] = t0.lnErr, t1.lnErr
    return t


def splitReadAligner(flankScoreSize: int, querySeq: str, qualConvert: "qscore_snp", queryQual: Sequence[int],
                     targetSeq: str, targetBpOffsetRange: Tuple[int, int],
                     ctx: Optional[Context] = None) -> SRAlignmentInfo:
    """splitReadAligner  GSC/SplitReadAlignment.hpp:63-70 (batch of one; raises like the reference on bad input)"""
    sp = _cabi.SeqPool()
    jobs = np.zeros(1, dtype=_cabi.SR_JOB_DT)
    jobs["q_off"][0], jobs["q_len"][0] = sp.add(querySeq)
    jobs["qualOff"][0], _ = sp.add(bytes(queryQual))
    jobs["t_off"][0], jobs["t_len"][0] = sp.add(targetSeq)
    jobs["bpBegin"][0], jobs["bpEnd"][0] = targetBpOffsetRange
    res = (ctx or default_context()).split_align_raw(qscore_tables(qualConvert, qualConvert), (0, 0),
                                                     flankScoreSize, sp.finish(), jobs)
    if int(res["status"][0]) != _cabi.MB200_OK:
        st = int(res["status"][0])
        raise Mb200Error(st, _cabi.ERROR_MESSAGES.get(st, "status %d" % st))
    return _cabi.decode_sr(res[0])
