"""Seeded synthetic workloads shared by tests/, tests/golden/make_golden.py and bench.py.

All generators are deterministic functions of their seed.  Nothing here touches a GPU or any library.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    COMP[_a] = _b


# ---------------------------------------------------------------------------------------------
# small python-level generators for parity tests / golden vectors
def _rseq(rng: random.Random, n: int, alpha: str = "ACGT") -> str:
    return "".join(rng.choice(alpha) for _ in range(n))


def _mutate(rng: random.Random, s: str, rate: float, alpha: str = "ACGT") -> str:
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(rng.choice(alpha))
            out.append(ch)
            continue
        if r < rate:
            out.append(rng.choice(alpha))
            continue
        out.append(ch)
    return "".join(out)


def random_alignment_cases(seed: int, n: int, kind: str, maxq: int = 260, maxr: int = 700) -> List[Tuple[str, ...]]:
    """contig-vs-reference shaped cases: deletions, insertions, jumps, soft-clips, N's, tiny inputs"""
    rng = random.Random(seed)
    cases = []
    for i in range(n):
        alpha = "ACGTN" if rng.random() < 0.25 else ("AC" if rng.random() < 0.1 else "ACGT")
        r1 = _rseq(rng, rng.randint(1, maxr), alpha)
        r2 = _rseq(rng, rng.randint(1, maxr), alpha)
        mode = rng.random()
        a = rng.randint(0, len(r1) - 1)
        left = r1[a:a + rng.randint(1, maxq // 2)]
        if kind == "jump":
            b = rng.randint(0, len(r2) - 1)
            right = r2[b:b + rng.randint(1, maxq // 2)]
            ins = _rseq(rng, rng.randint(0, 12)) if mode < 0.3 else ""
            q = left + ins + right
            if mode > 0.85:
                q = left  # aligns to one reference only
            elif mode > 0.8:
                q = _rseq(rng, 8) + q + _rseq(rng, 8)  # off-edge soft clips
            q = _mutate(rng, q, 0.04, alpha) or "A"
            cases.append((q[:maxq], r1, r2))
        else:
            # large deletion / insertion inside one reference
            gap = rng.randint(0, max(0, len(r1) - a - len(left) - 1))
            b = min(len(r1) - 1, a + len(left) + gap)
            right = r1[b:b + rng.randint(1, maxq // 2)]
            ins = _rseq(rng, rng.randint(20, 120)) if mode < 0.25 else ""
            q = left + ins + right
            if mode > 0.9:
                q = _rseq(rng, 6) + q + _rseq(rng, 6)
            q = _mutate(rng, q, 0.04, alpha) or "A"
            cases.append((q[:maxq], r1))
    # degenerate sizes
    if kind == "jump":
        cases += [("A", "A", "A"), ("ACGT", "T", "G"), ("A" * 40, "A" * 3, "A" * 50), ("ACGTACGTAC", "ACGTA", "CGTAC")]
    else:
        cases += [("A", "A"), ("ACGT", "T"), ("A" * 40, "A" * 7), ("T", "ACGTACGT" * 12)]
    return cases


def random_assembly_reads(rng: random.Random, nreads: int, rl: int, with_repeat: bool, n_rate: float = 0.02,
                          err: float = 0.005) -> List[str]:
    reflen = 1200
    refw = _rseq(rng, reflen)
    bp = reflen // 2
    kind = rng.random()
    if kind < 0.4:
        alt = refw[:bp] + refw[bp + rng.randint(50, 400):]
    elif kind < 0.6:
        alt = refw[:bp] + _rseq(rng, rng.randint(50, 300)) + refw[bp:]
    else:
        alt = refw[:bp] + _rseq(rng, 600)
    if with_repeat:
        per = rng.randint(1, 30)
        unit = _rseq(rng, per)
        rep = (unit * (200 // per + 2))[:rng.randint(60, 200)]
        alt = alt[:bp] + rep + alt[bp:]
    reads = []
    for i in range(nreads):
        src = alt if i < nreads * 3 // 4 else refw
        st = rng.randint(max(0, bp - rl + 10), max(0, min(len(src) - rl, bp - 10))) if len(src) > rl else 0
        r = list(src[st:st + rl])
        for j in range(len(r)):
            x = rng.random()
            if x < err:
                r[j] = rng.choice("ACGT")
            elif x < err + n_rate:
                r[j] = "N"
        reads.append("".join(r))
    return reads


def random_assembly_cases(seed: int, n: int) -> List[Tuple[Dict[str, int], List[str]]]:
    rng = random.Random(seed)
    out = []
    for i in range(n):
        opts: Dict[str, int] = {}
        m = rng.random()
        if m < 0.35:
            opts = dict(minWordLength=rng.choice([8, 12, 21]), maxWordLength=rng.choice([21, 31, 41]),
                        wordStepSize=rng.choice([3, 5, 10]))
        if rng.random() < 0.3:
            opts["minCoverage"] = rng.choice([2, 3])
        reads = random_assembly_reads(rng, rng.choice([6, 20, 64, 90]), rng.choice([50, 100, 150]),
                                      with_repeat=rng.random() < 0.4)
        if i == 0:
            reads = reads[:3] + ["", "ACGT", "N" * 60] + reads[3:]  # empty / short / all-N reads
        out.append((opts, reads))
    return out


def random_split_cases(seed: int, n: int) -> List[dict]:
    rng = random.Random(seed)
    out = []
    for i in range(n):
        T = rng.randint(160, 1300)
        tgt = _rseq(rng, T, "ACGTN" if rng.random() < 0.2 else "ACGT")
        Q = rng.randint(20, min(150, T - 1))
        bp = rng.randint(0, T - 1)
        hom = rng.randint(0, 10)
        st = max(0, min(T - Q, bp - rng.randint(0, Q)))
        q = list(tgt[st:st + Q])
        for j in range(Q):
            if rng.random() < 0.03:
                q[j] = rng.choice("ACGTN")
        if rng.random() < 0.3:
            q[Q // 2:] = list(_rseq(rng, Q - Q // 2))
        qual = [rng.choice([2, 2, 5, 20, 30, 37, 40, 41, 0, 70]) for _ in range(Q)]
        bpr = [bp, bp + hom]
        if rng.random() < 0.03:
            bpr = [T + 5, T + 6]
        out.append(dict(test="seeded-random-%d" % i, query="".join(q), qual=qual, target=tgt, bp=bpr,
                        table=rng.randint(0, 1)))
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[1]: "synthetic 50k candidate loci x 64 reads x 150 bp, assembly + 5 kb DP"
@dataclass
class CandidateBatch:
    """One batch of candidate loci laid out for the C ABI (all numpy, host memory)."""
    nLoci: int
    readsPerLocus: int
    readLen: int
    pool: np.ndarray         # uint8: all reads, then all reference windows (16-byte aligned starts)
    readOff: np.ndarray      # uint64 [nLoci+1] into readViews
    readViews: np.ndarray    # SEQVIEW_DT [nLoci*readsPerLocus]
    isJump: np.ndarray       # uint8 [nLoci]: 1 -> spanning candidate (GlobalJumpAligner), 0 -> complex (large-indel aligner)
    ref1: np.ndarray         # SEQVIEW_DT [nLoci]: reference window (complex) or ref1 (spanning)
    ref2: np.ndarray         # SEQVIEW_DT [nLoci]: ref2 for spanning candidates (len 0 otherwise)


def make_candidate_batch(n_loci: int, seed: int = 20260923, reads_per_locus: int = 64, read_len: int = 150,
                         window: int = 5000, seqview_dtype=None, locus_ids=None) -> CandidateBatch:
    """SURVEY.md 8(d) config 2.  Per locus: a `window`-bp iid reference; an SV at a breakpoint
    (40 % deletion 50-2000 bp, 20 % novel insertion 50-300 bp, 20 % inversion, 20 % translocation to a second
    window); 48 reads from the alt haplotype tiling the breakpoint (start in [bp-140, bp-10]) + 16 from the
    reference; 0.5 % substitutions; 2 % of bases low quality -> 'N' (minQval 5, SVCandidateAssembler.cpp:102-136);
    10 % of loci carry a tandem repeat (period 2-30, >= 100 bp) at the breakpoint.
    Deletion/insertion loci are 'complex' candidates aligned with the large-indel aligner against the whole
    window; inversion/translocation loci are 'spanning' candidates aligned with the jump aligner against two
    window/2 references (R1 + R2 = window).

    `locus_ids` (strong-scaling bench): generate exactly the loci with these GLOBAL indices of one fixed candidate list,
    each from its own stream default_rng([seed, id]) -- any rank can build any subset of the same list (n_loci is then
    len(locus_ids))."""
    from ._cabi import SEQVIEW_DT
    if locus_ids is not None:
        n_loci = len(locus_ids)
    rng = np.random.default_rng(seed)
    n_alt = reads_per_locus * 3 // 4
    read_stride = read_len
    reads_bytes = n_loci * reads_per_locus * read_stride
    ref_stride = (window + 15) // 16 * 16
    half = window // 2
    half_stride = (half + 15) // 16 * 16
    reads_bytes_al = (reads_bytes + 15) // 16 * 16
    pool = np.zeros(reads_bytes_al + n_loci * (ref_stride + 2 * half_stride) + 64, dtype=np.uint8)

    readViews = np.zeros(n_loci * reads_per_locus, dtype=SEQVIEW_DT)
    readViews["off"] = np.arange(n_loci * reads_per_locus, dtype=np.uint64) * read_stride
    readViews["len"] = read_len
    readOff = np.arange(n_loci + 1, dtype=np.uint64) * reads_per_locus
    isJump = np.zeros(n_loci, dtype=np.uint8)
    ref1 = np.zeros(n_loci, dtype=SEQVIEW_DT)
    ref2 = np.zeros(n_loci, dtype=SEQVIEW_DT)

    col = np.arange(read_len)
    ref_base = reads_bytes_al
    for li in range(n_loci):
        if locus_ids is not None:
            rng = np.random.default_rng([seed, int(locus_ids[li])])
        w1 = ACGT[rng.integers(0, 4, window)]
        kind = rng.random()
        bp = int(rng.integers(window // 5, window // 2))
        if kind < 0.4:  # deletion
            dl = int(rng.integers(50, max(51, min(2000, window - bp - 200)) + 1))
            alt = np.concatenate([w1[:bp], w1[bp + dl:]])
        elif kind < 0.6:  # novel insertion
            il = int(rng.integers(50, 301))
            alt = np.concatenate([w1[:bp], ACGT[rng.integers(0, 4, il)], w1[bp:]])
        else:
            w2 = ACGT[rng.integers(0, 4, window)]
            bp2 = int(rng.integers(window // 5, window // 2))
            if kind < 0.8:  # inversion: continue on the reverse complement of a second segment
                seg = COMP[w2[::-1]]
            else:  # translocation
                seg = w2
            alt = np.concatenate([w1[:bp], seg[bp2:]])
            isJump[li] = 1
        if rng.random() < 0.10:
            per = int(rng.integers(2, 31))
            ln = int(rng.integers(100, 181))
            rep = np.tile(ACGT[rng.integers(0, 4, per)], ln // per + 1)[:ln]
            alt = np.concatenate([alt[:bp], rep, alt[bp:]])
        # reads
        st_alt = rng.integers(bp - 140, bp - 9, n_alt)
        st_ref = rng.integers(bp - 140, bp - 9, reads_per_locus - n_alt)
        rd = np.empty((reads_per_locus, read_len), dtype=np.uint8)
        rd[:n_alt] = alt[st_alt[:, None] + col[None, :]]
        rd[n_alt:] = w1[st_ref[:, None] + col[None, :]]
        u = rng.random((reads_per_locus, read_len))
        sub = u < 0.005
        rd[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
        rd[(u >= 0.005) & (u < 0.025)] = ord("N")
        o = li * reads_per_locus * read_stride
        pool[o:o + reads_per_locus * read_len] = rd.reshape(-1)
        # references
        rb = ref_base + li * (ref_stride + 2 * half_stride)
        if isJump[li]:
            a0 = max(0, min(window - half, bp - half // 2))
            b0 = max(0, min(window - half, bp2 - half // 2))
            pool[rb + ref_stride:rb + ref_stride + half] = w1[a0:a0 + half]
            pool[rb + ref_stride + half_stride:rb + ref_stride + half_stride + half] = seg[b0:b0 + half]
            ref1["off"][li], ref1["len"][li] = rb + ref_stride, half
            ref2["off"][li], ref2["len"][li] = rb + ref_stride + half_stride, half
        else:
            pool[rb:rb + window] = w1
            ref1["off"][li], ref1["len"][li] = rb, window
    return CandidateBatch(n_loci, reads_per_locus, read_len, pool, readOff, readViews, isJump, ref1, ref2)


def make_large_insertion_batch(n_loci: int, seed: int = 7, reads_per_locus: int = 64, read_len: int = 150, window: int = 3000,
                               min_insert: int = 400, max_insert: int = 900) -> CandidateBatch:
    """Complex candidates whose novel insertion is too long to assemble through (GSC/SVCandidateAssemblyRefiner.cpp
    processLargeInsertion :833): reads cover the left edge (reference flank + start of the insert), the right edge (end
    of the insert + reference flank) and the reference allele, so the assembler returns a left and a right contig.
    A fifth of the loci carry a short (assemblable) insertion instead, a tenth a second copy of the left edge shifted by a
    few bases (competing left candidates).  Same layout as make_candidate_batch."""
    from ._cabi import SEQVIEW_DT
    rng = np.random.default_rng(seed)
    read_stride = read_len
    reads_bytes_al = (n_loci * reads_per_locus * read_stride + 15) // 16 * 16
    ref_stride = (window + 15) // 16 * 16
    pool = np.zeros(reads_bytes_al + n_loci * ref_stride + 64, dtype=np.uint8)
    readViews = np.zeros(n_loci * reads_per_locus, dtype=SEQVIEW_DT)
    readViews["off"] = np.arange(n_loci * reads_per_locus, dtype=np.uint64) * read_stride
    readViews["len"] = read_len
    readOff = np.arange(n_loci + 1, dtype=np.uint64) * reads_per_locus
    isJump = np.zeros(n_loci, dtype=np.uint8)
    ref1 = np.zeros(n_loci, dtype=SEQVIEW_DT)
    ref2 = np.zeros(n_loci, dtype=SEQVIEW_DT)
    col = np.arange(read_len)
    third = reads_per_locus // 3
    for li in range(n_loci):
        w1 = ACGT[rng.integers(0, 4, window)]
        bp = int(rng.integers(window // 3, 2 * window // 3))
        il = int(rng.integers(60, 200)) if rng.random() < 0.2 else int(rng.integers(min_insert, max_insert + 1))
        alt = np.concatenate([w1[:bp], ACGT[rng.integers(0, 4, il)], w1[bp:]])
        st_left = rng.integers(bp - 140, bp - 9, third)
        st_right = rng.integers(bp + il - 140, bp + il - 9, third)
        if rng.random() < 0.1:
            st_left[::2] += int(rng.integers(3, 12))
        st_ref = rng.integers(bp - 140, bp - 9, reads_per_locus - 2 * third)
        rd = np.empty((reads_per_locus, read_len), dtype=np.uint8)
        rd[:third] = alt[st_left[:, None] + col[None, :]]
        rd[third:2 * third] = alt[st_right[:, None] + col[None, :]]
        rd[2 * third:] = w1[st_ref[:, None] + col[None, :]]
        u = rng.random((reads_per_locus, read_len))
        sub = u < 0.004
        rd[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
        rd[(u >= 0.004) & (u < 0.014)] = ord("N")
        o = li * reads_per_locus * read_stride
        pool[o:o + reads_per_locus * read_len] = rd.reshape(-1)
        rb = reads_bytes_al + li * ref_stride
        pool[rb:rb + window] = w1
        ref1["off"][li], ref1["len"][li] = rb, window
    return CandidateBatch(n_loci, reads_per_locus, read_len, pool, readOff, readViews, isJump, ref1, ref2)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: split-read realign microbench
def make_split_batch(n_reads: int, seed: int = 5, n_contigs: int = 1000, contig_len: int = 1200, read_len: int = 150):
    """1000 alt contigs of 1200 bp with a breakpoint at 600+-100 and 0-10 bp of homology, plus two 1200-bp reference
    windows per contig; `n_reads` reads of 150 bp, each assigned to one contig (70 % spanning the breakpoint at a
    uniform offset, 30 % flank only), 1 % substitutions, qualities uniform 20-40 except 2 % at Q2; every read is
    aligned to 4 targets (alt contig with the bp1 / bp2 breakpoint views, ref window 1, ref window 2) like
    getReadSplitScore (GSC/SVScorerSplit.cpp:190-226).  Returns (pool, jobs[4*n_reads])."""
    from ._cabi import SR_JOB_DT
    rng = np.random.default_rng(seed)
    stride = (contig_len + 15) // 16 * 16
    tgt_bytes = n_contigs * 3 * stride
    rstride = read_len
    pool = np.zeros(tgt_bytes + 2 * n_reads * rstride + 64, dtype=np.uint8)
    alt = ACGT[rng.integers(0, 4, (n_contigs, contig_len))]
    ref1 = ACGT[rng.integers(0, 4, (n_contigs, contig_len))]
    ref2 = ACGT[rng.integers(0, 4, (n_contigs, contig_len))]
    bp = rng.integers(500, 701, n_contigs)
    hom = rng.integers(0, 11, n_contigs)
    # the references share the alt sequence on "their" side of the breakpoint
    col = np.arange(contig_len)
    ref1 = np.where(col[None, :] < (bp + hom)[:, None], alt, ref1)
    ref2 = np.where(col[None, :] >= bp[:, None], alt, ref2)
    tv = pool[:tgt_bytes].reshape(n_contigs, 3, stride)
    tv[:, 0, :contig_len] = alt
    tv[:, 1, :contig_len] = ref1
    tv[:, 2, :contig_len] = ref2
    cid = rng.integers(0, n_contigs, n_reads)
    span = rng.random(n_reads) < 0.7
    start = np.where(span, bp[cid] - rng.integers(10, read_len - 10, n_reads),
                     np.where(rng.random(n_reads) < 0.5, bp[cid] - read_len - rng.integers(0, 200, n_reads),
                              bp[cid] + hom[cid] + rng.integers(0, 200, n_reads)))
    start = np.clip(start, 0, contig_len - read_len - 1)
    rd = alt[cid[:, None], start[:, None] + np.arange(read_len)[None, :]]
    u = rng.random((n_reads, read_len))
    sub = u < 0.01
    rd[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
    qual = rng.integers(20, 41, (n_reads, read_len)).astype(np.uint8)
    qual[rng.random((n_reads, read_len)) < 0.02] = 2
    rbase = tgt_bytes
    pool[rbase:rbase + n_reads * rstride] = rd.reshape(-1)
    qbase = rbase + n_reads * rstride
    pool[qbase:qbase + n_reads * rstride] = qual.reshape(-1)
    jobs = np.zeros(4 * n_reads, dtype=SR_JOB_DT)
    r = np.repeat(np.arange(n_reads, dtype=np.uint64), 4)
    which = np.tile(np.arange(4), n_reads)
    jobs["q_off"] = rbase + r * rstride
    jobs["q_len"] = read_len
    jobs["qualOff"] = qbase + r * rstride
    c4 = np.repeat(cid, 4).astype(np.uint64)
    tsel = np.array([0, 0, 1, 2], dtype=np.uint64)[which]
    jobs["t_off"] = (c4 * 3 + tsel) * stride
    jobs["t_len"] = contig_len
    b4, h4 = np.repeat(bp[cid], 4), np.repeat(hom[cid], 4)
    jobs["bpBegin"] = b4 - 1
    jobs["bpEnd"] = b4 - 1 + h4
    jobs["table"] = np.array([1, 1, 0, 0], dtype=np.uint32)[which]  # alt alleles: snpPrior 0, ref alleles: 1e-3
    return pool, jobs


def make_split_candidates(n_cand: int, reads_per_cand: int = 40, seed: int = 11, contig_len: int = 900, read_len: int = 150):
    """Row a11 inputs: per candidate the four targets of its SVAlignmentInfo (alt contig with the bp1 / bp2 offset ranges,
    reference windows at bp1 / bp2) and the reads the scorer would meet around the two breakends: reads from the alt
    allele across bp1 or bp2, reads from either reference window, flank-only reads; 1 % substitutions, a few N, qualities
    20-40 with 2 % at Q2, MAPQ 0-60; every eighth read is met a second time (same fragment and read number).
    Returns (pool, candidates, readFirst, reads)."""
    from ._cabi import SPLIT_CAND_DT, SPLIT_READ_DT
    rng = np.random.default_rng(seed)
    stride = (contig_len + 15) // 16 * 16
    n_reads = n_cand * reads_per_cand
    tbytes = n_cand * 3 * stride
    pool = np.zeros(tbytes + 2 * n_reads * read_len + 64, dtype=np.uint8)
    cands = np.zeros(n_cand, dtype=SPLIT_CAND_DT)
    reads = np.zeros(n_reads, dtype=SPLIT_READ_DT)
    readFirst = np.arange(n_cand + 1, dtype=np.uint32) * reads_per_cand
    rbase, qbase = tbytes, tbytes + n_reads * read_len
    for c in range(n_cand):
        alt = ACGT[rng.integers(0, 4, contig_len)]
        bp1 = int(rng.integers(300, 400))
        ins = int(rng.integers(0, 120))          # inserted sequence between the two breakpoints on the contig
        bp2 = bp1 + ins
        hom = int(rng.integers(0, 9))
        ref1 = ACGT[rng.integers(0, 4, contig_len)]
        ref2 = ACGT[rng.integers(0, 4, contig_len)]
        ref1[:bp1 + hom] = alt[:bp1 + hom]          # reference at bp1 agrees with the contig left of the breakpoint
        ref2[bp2:] = alt[bp2:]                      # reference at bp2 agrees right of it
        o = c * 3 * stride
        pool[o:o + contig_len] = alt
        pool[o + stride:o + stride + contig_len] = ref1
        pool[o + 2 * stride:o + 2 * stride + contig_len] = ref2
        for name, off, b in (("altBp1", o, bp1), ("altBp2", o, bp2), ("refBp1", o + stride, bp1), ("refBp2", o + 2 * stride, bp2)):
            cands[name]["off"][c], cands[name]["len"][c] = off, contig_len
            cands[name]["bpBegin"][c], cands[name]["bpEnd"][c] = b - 1, b - 1 + hom
        for k in range(reads_per_cand):
            r = c * reads_per_cand + k
            if k % 8 == 7:   # the same read met again from the other breakend's region
                reads[r] = reads[r - 3]
                continue
            kind = rng.random()
            src = alt if kind < 0.5 else (ref1 if kind < 0.7 else (ref2 if kind < 0.9 else alt))
            center = bp1 if (kind < 0.25 or 0.5 <= kind < 0.7) else bp2
            start = center - int(rng.integers(10, read_len - 10)) if kind < 0.9 else center + int(rng.integers(20, 200))
            start = int(np.clip(start, 0, contig_len - read_len - 1))
            rd = src[start:start + read_len].copy()
            u = rng.random(read_len)
            rd[u < 0.01] = ACGT[rng.integers(0, 4, int((u < 0.01).sum()))]
            rd[(u >= 0.01) & (u < 0.013)] = ord("N")
            q = rng.integers(20, 41, read_len).astype(np.uint8)
            q[rng.random(read_len) < 0.02] = 2
            pool[rbase + r * read_len:rbase + (r + 1) * read_len] = rd
            pool[qbase + r * read_len:qbase + (r + 1) * read_len] = q
            reads["off"][r], reads["len"][r], reads["qualOff"][r] = rbase + r * read_len, read_len, qbase + r * read_len
            reads["fragment"][r], reads["isRead1"][r], reads["mapq"][r] = k // 2, k % 2, int(rng.integers(0, 61))
    return pool, cands, readFirst, reads
