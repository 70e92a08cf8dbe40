"""ctypes mirror of include/manta_b200.h plus batch marshalling helpers.

The helpers are parametrised by (library handle, symbol prefix, optional ctx) so the very same
marshalling drives the product library (prefix ``mb200_``, with a ctx) and -- from tests/ and
bench.py only -- the reference / oracle checker libraries (prefixes ``ref_`` / ``orc_``, no ctx).
Nothing in this module loads a library by itself.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# ---- status codes (include/manta_b200.h) ---------------------------------------------------
MB200_OK = 0
MB200_ERR_CUDA = 1
MB200_ERR_NO_DEVICE = 2
MB200_ERR_EMPTY_QUERY = 3
MB200_ERR_EMPTY_REF = 4
MB200_ERR_EMPTY_REF1 = 5
MB200_ERR_EMPTY_REF2 = 6
MB200_ERR_SPLIT_INPUT = 7
MB200_ERR_SPLIT_SCAN = 8
MB200_ERR_SPLIT_OUTCOME = 9
MB200_ERR_BAD_QUAL = 10
MB200_ERR_POOL_OVERFLOW = 11
MB200_ERR_BAD_ARG = 12
MB200_ERR_BAD_BASE = 13
MB200_ERR_ASM_OVERFLOW = 14
MB200_ERR_UNSUPPORTED = 15
MB200_ERR_SV_RANGE = 16

# messages the reference throws for the same conditions (GlobalJumpAlignerImpl.hpp:50-58 etc.)
ERROR_MESSAGES = {
    MB200_ERR_EMPTY_QUERY: "Unexpected empty query sequence",
    MB200_ERR_EMPTY_REF: "Unexpected empty reference sequence",
    MB200_ERR_EMPTY_REF1: "Unexpected empty reference1 sequence",
    MB200_ERR_EMPTY_REF2: "Unexpected empty reference2 sequence",
    MB200_ERR_SPLIT_INPUT: "Unexpected split read alignment input.",
    MB200_ERR_SPLIT_SCAN: "Unexpected split read alignment input condition: scanEnd < scanStart.",
    MB200_ERR_SPLIT_OUTCOME: "Unexpected split read alignment outcome.",
    MB200_ERR_BAD_QUAL: "basecall quality out of range",
}

# ALIGNPATH::align_t -> CIGAR character (lib/blt_util/align_path.hpp:36-61)
SEG_CODES = "?MIDNSHP=X"

MAX_QSCORE = 70


class Scores(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("open", C.c_int32), ("extend", C.c_int32),
                ("offEdge", C.c_int32), ("isAllowEdgeInsertion", C.c_int32)]


SEQVIEW_DT = np.dtype([("off", "<u8"), ("len", "<u4"), ("reserved", "<u4")], align=True)
ALN_RESULT_DT = np.dtype([("score", "<i4"), ("beginPos", "<i4"), ("isJumped", "<u4"), ("cigarLen", "<u4"),
                          ("cigarOff", "<u8")], align=True)
JUMP_RESULT_DT = np.dtype([("score", "<i4"), ("jumpInsertSize", "<u4"), ("jumpRange", "<u4"),
                           ("beginPos1", "<i4"), ("beginPos2", "<i4"), ("cigarLen1", "<u4"),
                           ("cigarLen2", "<u4"), ("reserved", "<u4"), ("cigarOff1", "<u8"),
                           ("cigarOff2", "<u8")], align=True)
SR_RESULT_DT = np.dtype([("alignPos", "<u4"), ("leftSize", "<u4"), ("homSize", "<u4"), ("rightSize", "<u4"),
                         ("leftMismatches", "<u4"), ("homMismatches", "<u4"), ("rightMismatches", "<u4"),
                         ("alignScore", "<u4"), ("alignLnLhood", "<f4"), ("isEvidence", "<u4"),
                         ("isTier2Evidence", "<u4"), ("evidence", "<f4"), ("status", "<i4"),
                         ("reserved", "<u4")], align=True)
SR_JOB_DT = np.dtype([("q_off", "<u8"), ("q_len", "<u4"), ("q_res", "<u4"), ("qualOff", "<u8"),
                      ("t_off", "<u8"), ("t_len", "<u4"), ("t_res", "<u4"), ("bpBegin", "<i4"),
                      ("bpEnd", "<i4"), ("table", "<u4"), ("reserved", "<u4")], align=True)
SPLIT_TARGET_DT = np.dtype([("off", "<u8"), ("len", "<u4"), ("res", "<u4"), ("bpBegin", "<i4"), ("bpEnd", "<i4")], align=True)
SPLIT_CAND_DT = np.dtype([("altBp1", SPLIT_TARGET_DT), ("altBp2", SPLIT_TARGET_DT), ("refBp1", SPLIT_TARGET_DT),
                          ("refBp2", SPLIT_TARGET_DT)], align=True)
SPLIT_READ_DT = np.dtype([("off", "<u8"), ("len", "<u4"), ("res", "<u4"), ("qualOff", "<u8"), ("fragment", "<u4"),
                          ("isRead1", "u1"), ("mapq", "u1"), ("reserved", "u1", (2,))], align=True)
SPLIT_EVIDENCE_DT = np.dtype([("status", "<i4"), ("isAnchored", "u1"), ("isTier2Anchored", "u1"), ("countedRef", "u1"),
                              ("countedAlt", "u1"), ("splitLnLhood", "<f4", (4,)), ("splitEvidence", "<f4", (4,)),
                              ("isSplitSupport", "u1", (4,)), ("isTier2SplitSupport", "u1", (4,)), ("refEvidence", "<f4"),
                              ("altEvidence", "<f4")], align=True)
SPLIT_ALLELE_DT = np.dtype([("splitReadCount", "<u4"), ("splitReadEvidence", "<f4"), ("splitReadMapQ", "<f4"),
                            ("reserved", "<u4")], align=True)
SPLIT_COUNTS_DT = np.dtype([("ref", SPLIT_ALLELE_DT), ("alt", SPLIT_ALLELE_DT)], align=True)
assert SPLIT_TARGET_DT.itemsize == 24 and SPLIT_CAND_DT.itemsize == 96 and SPLIT_READ_DT.itemsize == 32
assert SPLIT_EVIDENCE_DT.itemsize == 56 and SPLIT_COUNTS_DT.itemsize == 32
MB200_SPLIT_DUPLICATE = 1
ASM_OPTS_DT = np.dtype([("minWordLength", "<u4"), ("maxWordLength", "<u4"), ("wordStepSize", "<u4"),
                        ("minCoverage", "<u4"), ("minConservativeCoverage", "<u4"),
                        ("maxAssemblyCount", "<u4"), ("minUnusedReads", "<u4"), ("minSupportReads", "<u4")],
                       align=True)
CONTIG_DT = np.dtype([("seqOff", "<u8"), ("seqLen", "<u4"), ("seedReadCount", "<u4"),
                      ("conservativeBegin", "<i4"), ("conservativeEnd", "<i4"), ("supportOff", "<u8"),
                      ("rejectOff", "<u8")], align=True)
ASM_LOCUS_DT = np.dtype([("firstContig", "<u4"), ("nContigs", "<u4"), ("firstPseudo", "<u4"),
                         ("nPseudo", "<u4"), ("setWords", "<u4"), ("lastWordLength", "<u4"),
                         ("status", "<i4"), ("reserved", "<u4")], align=True)
QSCORE_TABLES_DT = np.dtype([("lnComp", "<f8", (2, MAX_QSCORE + 1)), ("lnErr", "<f8", (2, MAX_QSCORE + 1))])
TIMING_DT = np.dtype([("h2dMs", "<f4"), ("kernelMs", "<f4"), ("d2hMs", "<f4"), ("launches", "<u4"),
                      ("cells", "<u8"), ("algoBytes", "<u8")], align=True)
CONFIG_DT = np.dtype([("ctasPerSm", "<u4"), ("maxContigLen", "<u4"), ("scratchBytes", "<u8"),
                      ("reserved", "<u4", (4,))], align=True)

REF_CUTS_DT = np.dtype([("leadingCut", "<i4"), ("trailingCut", "<i4"), ("maxLeadingCut", "<i4"),
                        ("maxTrailingCut", "<i4")], align=True)
TRIM_RESULT_DT = np.dtype([("adjLeadingCut", "<i4"), ("adjTrailingCut", "<i4")], align=True)
QC_OPTS_DT = np.dtype([("filterScores", "<i4", (6,)), ("minCandidateVariantSize", "<u4"), ("reserved", "<u4")],
                      align=True)
CONTIG_QC_DT = np.dtype([("isCandidate", "<u4"), ("isLowQuality", "<u4"), ("segOrdinal", "<u4"), ("nSegments", "<u4"),
                         ("pathFirst", "<u4"), ("pathLast", "<u4"), ("variantSize", "<u4"), ("reserved", "<u4")],
                        align=True)
JUMP_ORIENT_DT = np.dtype([("isBp2AlignedFirst", "u1"), ("isBp1Reversed", "u1"), ("isBp2Reversed", "u1"), ("bp1RightOpen", "u1"),
                           ("bp2RightOpen", "u1"), ("reserved", "u1", (3,))])
JUMP_SV_DT = np.dtype([("valid", "<i4"), ("contig", "<u4"), ("bp1Begin", "<i4"), ("bp1End", "<i4"), ("bp2Begin", "<i4"),
                       ("bp2End", "<i4"), ("insertBegin", "<i4"), ("insertEnd", "<i4")], align=True)
assert JUMP_ORIENT_DT.itemsize == 8 and JUMP_SV_DT.itemsize == 32
SMALL_SV_DT = np.dtype([("locus", "<u4"), ("contig", "<u4"), ("segment", "<u4"), ("status", "<i4"), ("pathFirst", "<u4"),
                        ("pathLast", "<u4"), ("refBegin", "<i4"), ("refEnd", "<i4"), ("readBegin", "<i4"), ("readEnd", "<i4"),
                        ("ciposBegin", "<i4"), ("ciposEnd", "<i4"), ("bp1Begin", "<i4"), ("bp1End", "<i4"), ("bp2Begin", "<i4"),
                        ("bp2End", "<i4")], align=True)
assert SMALL_SV_DT.itemsize == 64
LARGE_INSERT_INFO_DT = np.dtype([("isLeftCandidate", "u1"), ("isRightCandidate", "u1"), ("reserved", "u1", (2,)),
                                 ("contigOffset", "<u4"), ("refOffset", "<u4"), ("score", "<i4")], align=True)
LARGE_INSERTION_DT = np.dtype([("status", "<i4"), ("leftContig", "<u4"), ("rightContig", "<u4"), ("breakDist", "<i4"),
                               ("breakScore", "<i4"), ("score", "<i4"), ("beginPos", "<i4"), ("cigarLen", "<u4"),
                               ("cigarOff", "<u8"), ("segFirst", "<u4"), ("segLast", "<u4"), ("insertTrimBegin", "<i4"),
                               ("insertTrimEnd", "<i4"), ("refBegin", "<i4"), ("refEnd", "<i4"), ("readBegin", "<i4"),
                               ("readEnd", "<i4"), ("ciposBegin", "<i4"), ("ciposEnd", "<i4"), ("bp1Begin", "<i4"),
                               ("bp1End", "<i4"), ("bp2Begin", "<i4"), ("bp2End", "<i4"), ("leftSeqBegin", "<u4"),
                               ("rightSeqLen", "<u4")], align=True)
assert LARGE_INSERT_INFO_DT.itemsize == 16 and LARGE_INSERTION_DT.itemsize == 104
LI_NONE, LI_SEGMENT_QC, LI_FLANK_QC, LI_TRIM_QC, LI_RANGE_ERROR, LI_DUPLICATE, LI_ACCEPTED = range(7)
CAND_OPTS_DT = np.dtype([("asmOpts", ASM_OPTS_DT), ("indelScores", "<i4", (6,)), ("largeIndelScore", "<i4"),
                         ("jumpScores", "<i4", (6,)), ("jumpScore", "<i4"), ("qcOpts", QC_OPTS_DT)], align=True)
CAND_TIMING_DT = np.dtype([("h2dMs", "<f4"), ("asmMs", "<f4"), ("indelMs", "<f4"), ("jumpMs", "<f4"), ("d2hMs", "<f4"),
                           ("totalMs", "<f4"), ("launches", "<u4"), ("nContigs", "<u4"), ("nIndelTasks", "<u4"),
                           ("nJumpTasks", "<u4"), ("asmWindows", "<u8"), ("indelCells", "<u8"), ("jumpCells", "<u8"),
                           ("asmBytes", "<u8"), ("indelBytes", "<u8"), ("jumpBytes", "<u8"), ("h2dBytes", "<u8"),
                           ("d2hBytes", "<u8"), ("trimMs", "<f4"), ("qcMs", "<f4")], align=True)
assert CAND_OPTS_DT.itemsize == 120 and CAND_TIMING_DT.itemsize == 112
assert REF_CUTS_DT.itemsize == 16 and TRIM_RESULT_DT.itemsize == 8 and QC_OPTS_DT.itemsize == 32
assert CONTIG_QC_DT.itemsize == 32

assert SEQVIEW_DT.itemsize == 16 and ALN_RESULT_DT.itemsize == 24 and JUMP_RESULT_DT.itemsize == 48
assert SR_RESULT_DT.itemsize == 56 and SR_JOB_DT.itemsize == 56 and CONTIG_DT.itemsize == 40
assert ASM_LOCUS_DT.itemsize == 32 and TIMING_DT.itemsize == 32 and CONFIG_DT.itemsize == 32


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---- python-side result objects (field names follow the reference structs) -----------------
@dataclass
class Alignment:  # lib/alignment/Alignment.hpp:28-41
    beginPos: int = 0
    cigar: str = ""  # apath_to_cigar(apath)

    def isAligned(self) -> bool:
        return self.cigar != ""


@dataclass
class AlignmentResult:  # lib/alignment/SingleRefAlignerShared.hpp:34-47
    score: int = 0
    isJumped: bool = False
    align: Alignment = field(default_factory=Alignment)


@dataclass
class JumpAlignmentResult:  # lib/alignment/JumpAlignerBase.hpp:37-56
    score: int = 0
    jumpInsertSize: int = 0
    jumpRange: int = 0
    align1: Alignment = field(default_factory=Alignment)
    align2: Alignment = field(default_factory=Alignment)


@dataclass
class SRAlignmentInfo:  # GSC/SplitReadAlignment.hpp:35-49
    alignPos: int = 0
    leftSize: int = 0
    homSize: int = 0
    rightSize: int = 0
    leftMismatches: int = 0
    homMismatches: int = 0
    rightMismatches: int = 0
    alignScore: int = 0
    alignLnLhood: float = 0.0
    isEvidence: bool = False
    isTier2Evidence: bool = False
    evidence: float = 0.0


@dataclass
class AssembledContig:  # lib/assembly/AssembledContig.hpp:38-50
    seq: str = ""
    seedReadCount: int = 0
    supportReads: Tuple[int, ...] = ()
    rejectReads: Tuple[int, ...] = ()
    conservativeRange: Tuple[int, int] = (0, 0)


@dataclass
class AssemblyReadInfo:  # lib/assembly/AssemblyReadInfo.hpp:29-41
    isUsed: bool = False
    isFiltered: bool = False
    isPseudo: bool = False
    contigIds: List[int] = field(default_factory=list)


def cigar_from_segments(segs: np.ndarray) -> str:
    return "".join("%d%s" % (int(s) >> 4, SEG_CODES[int(s) & 0xF]) for s in segs)


class Mb200Error(RuntimeError):
    """Raised for a non-zero status; mirrors illumina::common::GeneralException for the input checks."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


# ---- sequence pool builder ------------------------------------------------------------------
class SeqPool:
    """Concatenates byte strings into one pool with 16-byte aligned starts (TMA bulk-copy friendly)."""

    ALIGN = 16

    def __init__(self):
        self._chunks: List[bytes] = []
        self._size = 0

    def add(self, s) -> Tuple[int, int]:
        b = s.encode("latin-1") if isinstance(s, str) else bytes(s)
        off = self._size
        padded = (len(b) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self._chunks.append(b + b"\0" * (padded - len(b)))
        self._size += padded
        return off, len(b)

    def views(self, items: Sequence) -> np.ndarray:
        v = np.zeros(len(items), dtype=SEQVIEW_DT)
        for i, s in enumerate(items):
            v["off"][i], v["len"][i] = self.add(s)
        return v

    def finish(self) -> np.ndarray:
        data = b"".join(self._chunks) + b"\0" * 64
        return np.frombuffer(data, dtype=np.uint8).copy()


class Binding:
    """Function table for one implementation of the ABI (product, reference shim or C oracle)."""

    def __init__(self, lib: C.CDLL, prefix: str, ctx: Optional[C.c_void_p] = None):
        self.lib = lib
        self.prefix = prefix
        self.ctx = ctx

    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def _call(self, name, *args):
        f = self._fn(name)
        if self.ctx is not None:
            st = f(self.ctx, *args)
        else:
            st = f(*args)
        return st

    def _raise(self, st):
        msg = ERROR_MESSAGES.get(st)
        if self.ctx is not None:
            self.lib.mb200_last_error.restype = C.c_char_p
            detail = self.lib.mb200_last_error(self.ctx)
            if detail:
                msg = detail.decode()
        raise Mb200Error(st, msg or ("status %d" % st))

    # -- aligners -----------------------------------------------------------------------------
    def _with_cigar_pool(self, n, call, guess):
        cap = max(64, guess)
        for _ in range(8):
            cig = np.zeros(cap, dtype=np.uint32)
            used = C.c_uint64(0)
            st = call(cig, cap, used)
            if st == MB200_ERR_POOL_OVERFLOW:
                cap = max(int(used.value), cap * 2)
                continue
            if st != MB200_OK:
                self._raise(st)
            return cig, int(used.value)
        raise Mb200Error(MB200_ERR_POOL_OVERFLOW, "cigar pool overflow")

    def jump_align_raw(self, scores: Scores, jumpScore: int, pool: np.ndarray, q: np.ndarray, r1: np.ndarray,
                       r2: np.ndarray):
        n = len(q)
        res = np.zeros(n, dtype=JUMP_RESULT_DT)

        def call(cig, cap, used):
            args = [C.byref(scores), C.c_int32(jumpScore), C.c_uint32(n), _ptr(pool)]
            if self.ctx is not None:
                args.append(C.c_uint64(pool.size))
            args += [_ptr(q), _ptr(r1), _ptr(r2), _ptr(res), _ptr(cig), C.c_uint64(cap), C.byref(used)]
            return self._call("jump_align_batch", *args)

        cig, used = self._with_cigar_pool(n, call, 32 * n)
        return res, cig[:used]

    def single_align_raw(self, kind: str, scores: Scores, largeIndelScore: int, pool: np.ndarray,
                         q: np.ndarray, r: np.ndarray):
        n = len(q)
        res = np.zeros(n, dtype=ALN_RESULT_DT)

        def call(cig, cap, used):
            args = [C.byref(scores)]
            if kind == "indel":
                args.append(C.c_int32(largeIndelScore))
            args += [C.c_uint32(n), _ptr(pool)]
            if self.ctx is not None:
                args.append(C.c_uint64(pool.size))
            args += [_ptr(q), _ptr(r), _ptr(res), _ptr(cig), C.c_uint64(cap), C.byref(used)]
            return self._call("indel_align_batch" if kind == "indel" else "global_align_batch", *args)

        cig, used = self._with_cigar_pool(n, call, 32 * n)
        return res, cig[:used]

    def jump_align(self, scores: Scores, jumpScore: int, triples: Sequence[Tuple]) -> List[JumpAlignmentResult]:
        sp = SeqPool()
        q = sp.views([t[0] for t in triples])
        r1 = sp.views([t[1] for t in triples])
        r2 = sp.views([t[2] for t in triples])
        res, cig = self.jump_align_raw(scores, jumpScore, sp.finish(), q, r1, r2)
        return [decode_jump(res[i], cig) for i in range(len(triples))]

    def single_align(self, kind: str, scores: Scores, largeIndelScore: int,
                     pairs: Sequence[Tuple]) -> List[AlignmentResult]:
        sp = SeqPool()
        q = sp.views([t[0] for t in pairs])
        r = sp.views([t[1] for t in pairs])
        res, cig = self.single_align_raw(kind, scores, largeIndelScore, sp.finish(), q, r)
        return [decode_aln(res[i], cig) for i in range(len(pairs))]

    # -- split read ---------------------------------------------------------------------------
    def split_align_raw(self, tables: np.ndarray, priors: Tuple[float, float], flank: int, pool: np.ndarray,
                        jobs: np.ndarray) -> np.ndarray:
        n = len(jobs)
        res = np.zeros(n, dtype=SR_RESULT_DT)
        if self.prefix == "ref_":
            st = self._call("split_align_batch", C.c_double(priors[0]), C.c_double(priors[1]),
                            C.c_uint32(flank), C.c_uint32(n), _ptr(pool), _ptr(jobs), _ptr(res))
        else:
            args = [_ptr(tables), C.c_uint32(flank), C.c_uint32(n), _ptr(pool)]
            if self.ctx is not None:
                args.append(C.c_uint64(pool.size))
            args += [_ptr(jobs), _ptr(res)]
            st = self._call("split_align_batch", *args)
        if st != MB200_OK:
            self._raise(st)
        return res

    def split_score_raw(self, tables: np.ndarray, flank: int, minMapQ: int, minTier2MapQ: int, cands: np.ndarray,
                        readFirst: np.ndarray, reads: np.ndarray, pool: np.ndarray):
        """getReadSplitScore for every read of every candidate (row a11); returns (sr[4 per read], evidence, counts)"""
        nC, nR = len(cands), len(reads)
        sr = np.zeros(max(1, 4 * nR), dtype=SR_RESULT_DT)
        ev = np.zeros(max(1, nR), dtype=SPLIT_EVIDENCE_DT)
        counts = np.zeros(max(1, nC), dtype=SPLIT_COUNTS_DT)
        args = [_ptr(tables), C.c_uint32(flank), C.c_uint32(minMapQ), C.c_uint32(minTier2MapQ), C.c_uint32(nC), _ptr(cands),
                _ptr(readFirst), _ptr(reads), _ptr(pool)]
        if self.ctx is not None:
            args.append(C.c_uint64(pool.size))
        st = self._call("split_score_batch", *args, _ptr(sr), _ptr(ev), _ptr(counts))
        if st != MB200_OK:
            self._raise(st)
        return sr[:4 * nR], ev[:nR], counts[:nC]

    # -- assembler ----------------------------------------------------------------------------
    def assemble_raw(self, opts: np.ndarray, readOff: np.ndarray, views: np.ndarray, pool: np.ndarray):
        nLoci = len(readOff) - 1
        maxAsm = int(opts["maxAssemblyCount"][0])
        contigCap = nLoci * maxAsm
        pseudoCap = nLoci * 2 * maxAsm
        nReadsMax = int(np.max(np.diff(readOff))) if nLoci else 0
        totalBases = int(views["len"].sum()) if len(views) else 0
        poolCap = max(1 << 16, 4 * totalBases)
        setCap = max(1024, contigCap * 2 * ((nReadsMax + 2 * maxAsm + 31) // 32))
        for _ in range(6):
            loci = np.zeros(nLoci, dtype=ASM_LOCUS_DT)
            contigs = np.zeros(contigCap, dtype=CONTIG_DT)
            pseudo = np.zeros(pseudoCap, dtype=SEQVIEW_DT)
            cpool = np.zeros(poolCap, dtype=np.uint8)
            spool = np.zeros(setCap, dtype=np.uint32)
            args = [_ptr(opts), C.c_uint32(nLoci), _ptr(readOff), _ptr(views), _ptr(pool)]
            if self.ctx is not None:
                args.append(C.c_uint64(pool.size))
            args += [_ptr(loci), _ptr(contigs), C.c_uint64(contigCap), _ptr(pseudo), C.c_uint64(pseudoCap),
                     _ptr(cpool), C.c_uint64(poolCap), _ptr(spool), C.c_uint64(setCap)]
            st = self._call("assemble_batch", *args)
            if st == MB200_ERR_POOL_OVERFLOW:
                poolCap *= 4
                setCap *= 2
                continue
            if st != MB200_OK:
                self._raise(st)
            return loci, contigs, pseudo, cpool, spool
        raise Mb200Error(MB200_ERR_POOL_OVERFLOW, "assembler output pool overflow")

    # -- reference-window trim / contig QC + selection ----------------------------------------
    def ref_trim_raw(self, pool: np.ndarray, contig: np.ndarray, ref: np.ndarray, cuts: np.ndarray) -> np.ndarray:
        n = len(contig)
        out = np.zeros(n, dtype=TRIM_RESULT_DT)
        args = [C.c_uint32(n), _ptr(pool)]
        if self.ctx is not None:
            args.append(C.c_uint64(pool.size))
        st = self._call("ref_trim_batch", *args, _ptr(contig), _ptr(ref), _ptr(cuts), _ptr(out))
        if st != MB200_OK:
            self._raise(st)
        return out

    def select_contigs_raw(self, opts: np.ndarray, isJump: np.ndarray, slotFirst: np.ndarray, nContigs: np.ndarray,
                           aln: np.ndarray, jump: np.ndarray, cigar: np.ndarray, pool: np.ndarray,
                           contigSeq: np.ndarray, ref1: np.ndarray, support: np.ndarray):
        nLoci, nSlots = len(isJump), len(contigSeq)
        qc = np.zeros(nSlots, dtype=CONTIG_QC_DT)
        best = np.zeros(nLoci, dtype=np.int32)
        args = [_ptr(opts), C.c_uint32(nLoci), _ptr(isJump), _ptr(slotFirst), _ptr(nContigs), C.c_uint32(nSlots),
                _ptr(aln), _ptr(jump), _ptr(cigar)]
        if self.ctx is not None:
            args.append(C.c_uint64(cigar.size))
        args.append(_ptr(pool))
        if self.ctx is not None:
            args.append(C.c_uint64(pool.size))
        st = self._call("select_contigs_batch", *args, _ptr(contigSeq), _ptr(ref1), _ptr(support), _ptr(qc), _ptr(best))
        if st != MB200_OK:
            self._raise(st)
        return qc, best

    def small_sv_raw(self, minSize: int, isJump: np.ndarray, slotFirst: np.ndarray, best: np.ndarray, qc: np.ndarray,
                     aln: np.ndarray, cigar: np.ndarray, pool: np.ndarray, contigSeq: np.ndarray, ref1: np.ndarray):
        """refined small SVs of the selected contigs; returns (svFirst[nLoci+1], records)"""
        nLoci, nSlots = len(isJump), len(contigSeq)
        svFirst = np.zeros(nLoci + 1, dtype=np.uint32)
        cap = max(16, nLoci)
        for _ in range(4):
            svs = np.zeros(cap, dtype=SMALL_SV_DT)
            used = C.c_uint64(0)
            args = [C.c_uint32(minSize), C.c_uint32(nLoci), _ptr(isJump), _ptr(slotFirst), C.c_uint32(nSlots), _ptr(best),
                    _ptr(qc), _ptr(aln), _ptr(cigar)]
            if self.ctx is not None:
                args.append(C.c_uint64(cigar.size))
            args.append(_ptr(pool))
            if self.ctx is not None:
                args.append(C.c_uint64(pool.size))
            st = self._call("small_sv_batch", *args, _ptr(contigSeq), _ptr(ref1), _ptr(svFirst), _ptr(svs), C.c_uint64(cap),
                            C.byref(used))
            if st == MB200_ERR_POOL_OVERFLOW:
                cap = int(used.value) + 16
                continue
            if st != MB200_OK:
                self._raise(st)
            return svFirst, svs[:int(used.value)]
        raise Mb200Error(MB200_ERR_POOL_OVERFLOW, "small SV record overflow")

    def large_insertion_raw(self, edgeScores, completeScores, isJump: np.ndarray, slotFirst: np.ndarray, nContigs: np.ndarray,
                            aln: np.ndarray, cigar: np.ndarray, pool: np.ndarray, contigSeq: np.ndarray,
                            consRange: np.ndarray, ref1: np.ndarray, cuts: Optional[np.ndarray] = None,
                            svFirst: Optional[np.ndarray] = None, svs: Optional[np.ndarray] = None):
        """large insertions with an incompletely assembled insert (GSC/SVCandidateAssemblyRefiner.cpp:2079-2112, :833-1007);
        returns (info[nSlots], records[nLoci], cigar pool of the joined-contig alignments)"""
        nLoci, nSlots = len(isJump), len(contigSeq)
        info = np.zeros(max(1, nSlots), dtype=LARGE_INSERT_INFO_DT)
        out = np.zeros(max(1, nLoci), dtype=LARGE_INSERTION_DT)
        es, cs = Scores(*edgeScores), Scores(*completeScores)
        consRange = np.ascontiguousarray(consRange, dtype=np.int32)
        cap = 256
        for _ in range(4):
            oc = np.zeros(cap, dtype=np.uint32)
            used = C.c_uint64(0)
            args = [C.byref(es), C.byref(cs), C.c_uint32(nLoci), _ptr(isJump), _ptr(slotFirst), _ptr(nContigs), C.c_uint32(nSlots),
                    _ptr(aln), _ptr(cigar)]
            if self.ctx is not None:
                args.append(C.c_uint64(cigar.size))
            args.append(_ptr(pool))
            if self.ctx is not None:
                args.append(C.c_uint64(pool.size))
            st = self._call("large_insertion_batch", *args, _ptr(contigSeq), _ptr(consRange), _ptr(ref1),
                            _ptr(cuts) if cuts is not None else None, _ptr(svFirst) if svFirst is not None else None,
                            _ptr(svs) if svs is not None and len(svs) else None, _ptr(info), _ptr(out), _ptr(oc),
                            C.c_uint64(cap), C.byref(used))
            if st == MB200_ERR_POOL_OVERFLOW:
                cap = int(used.value) + 64
                continue
            if st != MB200_OK:
                self._raise(st)
            return info[:nSlots], out[:nLoci], oc[:int(used.value)]
        raise Mb200Error(MB200_ERR_POOL_OVERFLOW, "large insertion cigar overflow")

    def jump_sv_raw(self, isJump: np.ndarray, slotFirst: np.ndarray, nSlots: int, best: np.ndarray, jump: np.ndarray,
                    cigar: np.ndarray, ref1: np.ndarray, ref2: np.ndarray, orient: np.ndarray) -> np.ndarray:
        """refined breakends of the selected spanning contigs (one record per candidate)"""
        nLoci = len(isJump)
        out = np.zeros(nLoci, dtype=JUMP_SV_DT)
        args = [C.c_uint32(nLoci), _ptr(isJump), _ptr(slotFirst), C.c_uint32(nSlots), _ptr(best), _ptr(jump), _ptr(cigar)]
        if self.ctx is not None:
            args.append(C.c_uint64(cigar.size))
        st = self._call("jump_sv_batch", *args, _ptr(ref1), _ptr(ref2), _ptr(orient), _ptr(out))
        if st != MB200_OK:
            self._raise(st)
        return out

    def candidates_jump_svs(self, orient: np.ndarray) -> np.ndarray:
        """refined spanning breakends of the last fused run (product library only)"""
        out = np.zeros(self._cand_shape[0], dtype=JUMP_SV_DT)
        st = self._call("candidates_jump_svs", _ptr(orient), _ptr(out))
        if st != MB200_OK:
            self._raise(st)
        return out

    def candidates_small_svs(self):
        """refined small SVs of the last fused run (product library only)"""
        nLoci = self._cand_shape[0]
        svFirst = np.zeros(nLoci + 1, dtype=np.uint32)
        cap = max(16, nLoci)
        for _ in range(4):
            svs = np.zeros(cap, dtype=SMALL_SV_DT)
            used = C.c_uint64(0)
            st = self._call("candidates_small_svs", _ptr(svFirst), _ptr(svs), C.c_uint64(cap), C.byref(used))
            if st == MB200_ERR_POOL_OVERFLOW:
                cap = int(used.value) + 16
                continue
            if st != MB200_OK:
                self._raise(st)
            return svFirst, svs[:int(used.value)]
        raise Mb200Error(MB200_ERR_POOL_OVERFLOW, "small SV record overflow")

    # -- fused candidate path (product library only) -----------------------------------------
    def candidates_upload(self, batch, cuts: Optional[np.ndarray] = None) -> None:
        st = self._call("candidates_upload", C.c_uint32(batch.nLoci), _ptr(batch.readOff), _ptr(batch.readViews),
                        _ptr(batch.isJump), _ptr(batch.ref1), _ptr(batch.ref2), _ptr(batch.pool),
                        C.c_uint64(batch.pool.size), _ptr(cuts) if cuts is not None else None)
        if st != MB200_OK:
            self._raise(st)
        # the copy is asynchronous: keep the host arrays alive until the consuming run has returned
        self._cand_keep = getattr(self, "_cand_keep", [])[-2:] + [(batch, cuts)]
        # staged batches are consumed first-in first-out: the shape of the NEWEST upload is not the shape of the next run
        self._cand_fifo = getattr(self, "_cand_fifo", []) + [(batch.nLoci, int(batch.readViews["len"].sum()), batch.readsPerLocus)]

    def candidates_run(self, opts: np.ndarray, out: Optional[dict] = None, select: bool = False, zeros=np.zeros) -> dict:
        """run on the staged batch; `out` (from a previous call) is reused to avoid re-allocating host buffers.
        select=True also returns the per-contig trim / QC records and the per-locus bestAlignmentIndex.
        `zeros(n, dtype=...)` allocates the host result buffers (pass a pinned-memory allocator for faster copies)."""
        # the batch this run consumes: the oldest staged one, else the one consumed last (repeat on resident inputs)
        if getattr(self, "_cand_fifo", None):
            self._cand_shape = self._cand_fifo.pop(0)
        nLoci, totalBases, rpl = self._cand_shape
        pend = C.c_uint32(0)
        if self._call("candidates_pending", C.byref(pend), None, None) != MB200_OK or pend.value != nLoci:
            raise Mb200Error(MB200_ERR_BAD_ARG, "staged-batch bookkeeping out of step with the library")
        A = int(opts["asmOpts"]["maxAssemblyCount"][0])
        if out is not None and (len(out["loci"]) < nLoci or len(out["contigs"]) < nLoci * A or
                                ("qc" in out and len(out["qc"]) < nLoci * A)):
            out = None  # buffers of a smaller batch: the library writes nLoci*A records, never reuse them
        if out is None:
            out = dict(loci=zeros(nLoci, dtype=ASM_LOCUS_DT), contigs=zeros(nLoci * A, dtype=CONTIG_DT),
                       pseudo=zeros(nLoci * 2 * A, dtype=SEQVIEW_DT),
                       contigPool=zeros(max(1 << 16, totalBases // 2), dtype=np.uint8),
                       setPool=zeros(max(1024, nLoci * A * 2 * ((rpl + 2 * A + 31) // 32)), dtype=np.uint32),
                       aln=zeros(nLoci * A, dtype=ALN_RESULT_DT), jump=zeros(nLoci * A, dtype=JUMP_RESULT_DT),
                       cigar=zeros(max(4096, nLoci * A * 24), dtype=np.uint32))
        if select and "qc" not in out:
            out.update(trim=zeros(nLoci * A, dtype=TRIM_RESULT_DT), qc=zeros(nLoci * A, dtype=CONTIG_QC_DT),
                       best=zeros(nLoci, dtype=np.int32))
        sel = [_ptr(out["trim"]), _ptr(out["qc"]), _ptr(out["best"])] if select else [None, None, None]
        for _ in range(6):
            used = [C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)]
            st = self._call("candidates_run", _ptr(opts), _ptr(out["loci"]), _ptr(out["contigs"]), _ptr(out["pseudo"]),
                            _ptr(out["contigPool"]), C.c_uint64(out["contigPool"].size), C.byref(used[0]),
                            _ptr(out["setPool"]), C.c_uint64(out["setPool"].size), C.byref(used[1]), _ptr(out["aln"]),
                            _ptr(out["jump"]), _ptr(out["cigar"]), C.c_uint64(out["cigar"].size), C.byref(used[2]), *sel)
            if st == MB200_ERR_POOL_OVERFLOW:
                if used[0].value > out["contigPool"].size:
                    out["contigPool"] = zeros(int(used[0].value * 1.25) + 64, dtype=np.uint8)
                if used[1].value > out["setPool"].size:
                    out["setPool"] = zeros(int(used[1].value * 1.25) + 64, dtype=np.uint32)
                if used[2].value > out["cigar"].size:
                    out["cigar"] = zeros(int(used[2].value * 1.25) + 64, dtype=np.uint32)
                continue
            if st != MB200_OK:
                self._raise(st)
            out["used"] = tuple(int(u.value) for u in used)
            return out
        raise Mb200Error(MB200_ERR_POOL_OVERFLOW, "candidate output pools overflow")

    def candidates_mt(self, opts: np.ndarray, batch, n_loci: int, threads: int, cuts: Optional[np.ndarray] = None,
                      select: bool = False):
        """checker libraries only (ref_/orc_): the multi-threaded CPU composition; returns
        (nContigs, cells, scoreSum, qcSum)"""
        nc, cells, ssum, qsum = C.c_uint64(), C.c_uint64(), C.c_int64(), C.c_int64()
        st = self._call("candidates_mt", _ptr(opts), C.c_uint32(n_loci), _ptr(batch.readOff), _ptr(batch.readViews),
                        _ptr(batch.isJump), _ptr(batch.ref1), _ptr(batch.ref2), _ptr(batch.pool),
                        _ptr(cuts) if cuts is not None else None, C.c_int(1 if select else 0), C.c_int(threads),
                        C.byref(nc), C.byref(cells), C.byref(ssum), C.byref(qsum))
        if st != MB200_OK:
            self._raise(st)
        return int(nc.value), int(cells.value), int(ssum.value), int(qsum.value)

    # -- the one collective of the path (product library only) --------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self.lib.mb200_nccl_unique_id.restype = C.c_int
        st = self.lib.mb200_nccl_unique_id(buf)
        if st != MB200_OK:
            raise Mb200Error(st, "mb200_nccl_unique_id failed (status %d): libnccl.so.2 not available?" % st)
        return bytes(buf)

    def nccl_comm_init(self, n_ranks: int, unique_id: bytes, rank: int) -> C.c_void_p:
        comm = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        st = self._call("nccl_comm_init", C.c_int(n_ranks), idbuf, C.c_int(rank), C.byref(comm))
        if st != MB200_OK:
            self._raise(st)
        return comm

    def nccl_comm_destroy(self, comm) -> None:
        self.lib.mb200_nccl_comm_destroy(comm)

    def gather_results(self, comm, root: int, local_dev_ptr: int, local_bytes: int, root_dev_ptr: int, root_cap: int,
                       n_ranks: int):
        """mb200_gather_results on device pointers; returns (status, bytes per rank)"""
        sizes = (C.c_uint64 * n_ranks)()
        st = self._call("gather_results", comm, C.c_int(root), C.c_void_p(local_dev_ptr), C.c_uint64(local_bytes),
                        C.c_void_p(root_dev_ptr), C.c_uint64(root_cap), sizes)
        return st, [int(x) for x in sizes]

    def candidates_timing(self) -> dict:
        t = np.zeros(1, dtype=CAND_TIMING_DT)
        self._call("candidates_timing", _ptr(t))
        return {k: (float(t[k][0]) if t[k].dtype.kind == "f" else int(t[k][0])) for k in t.dtype.names}

    def assemble(self, opts: np.ndarray, loci_reads: Sequence[Sequence]):
        """returns per locus (contigs, pseudoReads, readInfo) like runIterativeAssembler's outputs"""
        sp = SeqPool()
        readOff = np.zeros(len(loci_reads) + 1, dtype=np.uint64)
        allReads = []
        for i, reads in enumerate(loci_reads):
            allReads.extend(reads)
            readOff[i + 1] = len(allReads)
        views = sp.views(allReads)
        loci, contigs, pseudo, cpool, spool = self.assemble_raw(opts, readOff, views, sp.finish())
        out = []
        for i, reads in enumerate(loci_reads):
            out.append(decode_locus(loci[i], contigs, pseudo, cpool, spool, len(reads)))
        return out


def default_candidate_opts(**asm_kw) -> np.ndarray:
    """SVRefinerOptions defaults (lib/options/SVRefinerOptions.hpp:40-45) + IterativeAssemblerOptions defaults"""
    o = np.zeros(1, dtype=CAND_OPTS_DT)
    o["asmOpts"] = default_asm_opts(**asm_kw)[0]
    o["indelScores"] = (2, -8, -24, -1, -1, 0)
    o["largeIndelScore"] = -100
    o["jumpScores"] = (2, -8, -12, -1, -1, 0)
    o["jumpScore"] = -100
    o["qcOpts"] = default_qc_opts()[0]
    return o


def default_qc_opts() -> np.ndarray:
    """contigFilterScores (lib/options/SVRefinerOptions.hpp:50) + minCandidateVariantSize (ReadScannerOptions.hpp:68)"""
    o = np.zeros(1, dtype=QC_OPTS_DT)
    o["filterScores"] = (2, -8, -18, 0, -1, 0)
    o["minCandidateVariantSize"] = 10
    return o


def default_ref_cuts(n: int) -> np.ndarray:
    """the normal-case cuts of a complex candidate's window (GSC/SVCandidateAssemblyRefiner.cpp:1884-1918)"""
    c = np.zeros(n, dtype=REF_CUTS_DT)
    c["leadingCut"] = 100
    c["trailingCut"] = 100
    c["maxLeadingCut"] = 800
    c["maxTrailingCut"] = 800
    return c


def qc_checksum(isJump, nContigs, A, trim, qc, best, with_trim: bool) -> int:
    """the QC checksum the CPU checker libraries report from candidates_mt, rebuilt from flat outputs
    (slot = locus*A + contig)"""
    total = 0
    nLoci = len(isJump)
    slot = (np.arange(nLoci, dtype=np.int64)[:, None] * A + np.arange(A, dtype=np.int64)[None, :])
    live = np.arange(A)[None, :] < np.asarray(nContigs, dtype=np.int64)[:, None]
    s = slot[live]
    q = qc[s]
    total += int((q["isCandidate"].astype(np.int64) + 2 * q["isLowQuality"].astype(np.int64) +
                  4 * q["variantSize"].astype(np.int64) + 8 * q["nSegments"].astype(np.int64) +
                  64 * q["pathFirst"].astype(np.int64) + 128 * q["pathLast"].astype(np.int64)).sum())
    if with_trim:
        cplx = (np.asarray(isJump) == 0)[:, None] & live
        t = trim[slot[cplx]]
        total += int((t["adjLeadingCut"].astype(np.int64) * 16 + t["adjTrailingCut"].astype(np.int64) * 32).sum())
    total += int(((np.asarray(best, dtype=np.int64) + 1) * 1000003).sum())
    return total


def default_asm_opts(**kw) -> np.ndarray:
    o = np.zeros(1, dtype=ASM_OPTS_DT)
    o["minWordLength"] = 41
    o["maxWordLength"] = 76
    o["wordStepSize"] = 5
    o["minCoverage"] = 1
    o["minConservativeCoverage"] = 2
    o["maxAssemblyCount"] = 10
    o["minUnusedReads"] = 3
    o["minSupportReads"] = 2
    for k, v in kw.items():
        o[k] = v
    return o


def _segs(cig, off, ln):
    return cigar_from_segments(cig[int(off):int(off) + int(ln)])


def decode_aln(r, cig) -> AlignmentResult:
    return AlignmentResult(int(r["score"]), bool(r["isJumped"]),
                           Alignment(int(r["beginPos"]), _segs(cig, r["cigarOff"], r["cigarLen"])))


def decode_jump(r, cig) -> JumpAlignmentResult:
    return JumpAlignmentResult(int(r["score"]), int(r["jumpInsertSize"]), int(r["jumpRange"]),
                               Alignment(int(r["beginPos1"]), _segs(cig, r["cigarOff1"], r["cigarLen1"])),
                               Alignment(int(r["beginPos2"]), _segs(cig, r["cigarOff2"], r["cigarLen2"])))


def decode_sr(r) -> SRAlignmentInfo:
    return SRAlignmentInfo(int(r["alignPos"]), int(r["leftSize"]), int(r["homSize"]), int(r["rightSize"]),
                           int(r["leftMismatches"]), int(r["homMismatches"]), int(r["rightMismatches"]),
                           int(r["alignScore"]), float(r["alignLnLhood"]), bool(r["isEvidence"]),
                           bool(r["isTier2Evidence"]), float(r["evidence"]))


def _set_from_words(words: np.ndarray) -> Tuple[int, ...]:
    out = []
    for w, v in enumerate(words):
        v = int(v)
        while v:
            b = (v & -v).bit_length() - 1
            out.append(w * 32 + b)
            v &= v - 1
    return tuple(out)


def decode_locus(L, contigs, pseudo, cpool, spool, nNormalReads: int):
    sw = int(L["setWords"])
    cs = []
    for ci in range(int(L["firstContig"]), int(L["firstContig"]) + int(L["nContigs"])):
        c = contigs[ci]
        seq = bytes(cpool[int(c["seqOff"]):int(c["seqOff"]) + int(c["seqLen"])]).decode("latin-1")
        sup = _set_from_words(spool[int(c["supportOff"]):int(c["supportOff"]) + sw])
        rej = _set_from_words(spool[int(c["rejectOff"]):int(c["rejectOff"]) + sw])
        cs.append(AssembledContig(seq, int(c["seedReadCount"]), sup, rej,
                                  (int(c["conservativeBegin"]), int(c["conservativeEnd"]))))
    ps = []
    for pi in range(int(L["firstPseudo"]), int(L["firstPseudo"]) + int(L["nPseudo"])):
        v = pseudo[pi]
        ps.append(bytes(cpool[int(v["off"]):int(v["off"]) + int(v["len"])]).decode("latin-1"))
    # AssemblyReadOutput as selectContigs leaves it (IterativeAssembler.cpp:822-831)
    info = [AssemblyReadInfo() for _ in range(nNormalReads + len(ps))]
    for r in range(nNormalReads, len(info)):
        info[r].isPseudo = True
    for cid, c in enumerate(cs):
        for rd in c.supportReads:
            # When the LAST word length fails, the reference replaces the pseudo reads after the contigs were
            # built (IterativeAssembler.cpp:882-910), so a contig can name a pseudo read index past the end of
            # the final read list (the reference indexes its vector out of bounds there, :765,:826).  Such
            # indices are >= the normal read count and are treated as pseudo reads.
            while rd >= len(info):
                info.append(AssemblyReadInfo(isPseudo=True))
            info[rd].isUsed = True
            info[rd].contigIds.append(cid)
    return cs, ps, info
