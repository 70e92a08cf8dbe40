"""Input builders for the reference-window trim and contig QC / selection tests (rows a5 / a9).

Two families:
  * pipeline cases: a synthetic candidate batch pushed through a checker's assembler and aligners, so the QC sees
    alignments of the kind the refiner produces;
  * fuzz cases: random alignment paths with sequences made consistent with them, tuned to hit the corners of the QC
    (short flanks, clipped ends, adjacent insert/delete runs, several candidate indels, repeated flanks that trigger
    the ambiguity filter, empty jump arms).
"""
import numpy as np

from manta_b200 import _cabi as cb

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
M, I, D, S, EQ, X = 1, 2, 3, 5, 8, 9


class QcCase:
    """flat arrays in the layout of mb200_select_contigs_batch"""

    def __init__(self, isJump, slotFirst, nContigs, aln, jump, cigar, pool, contigSeq, ref1, support):
        self.isJump, self.slotFirst, self.nContigs = isJump, slotFirst, nContigs
        self.aln, self.jump, self.cigar, self.pool = aln, jump, cigar, pool
        self.contigSeq, self.ref1, self.support = contigSeq, ref1, support

    def small_svs(self, binding, qc, best, minSize=10):
        return binding.small_sv_raw(minSize, self.isJump, self.slotFirst, best, qc, self.aln, self.cigar, self.pool,
                                    self.contigSeq, self.ref1)

    def run(self, binding, opts=None):
        opts = cb.default_qc_opts() if opts is None else opts
        return binding.select_contigs_raw(opts, self.isJump, self.slotFirst, self.nContigs, self.aln, self.jump,
                                          self.cigar, self.pool, self.contigSeq, self.ref1, self.support)


def pipeline_case(binding, batch, opts, cuts=None):
    """assemble + [trim] + align with `binding` (a checker), packed for the QC.  Returns (QcCase, trims)."""
    asm_opts = np.zeros(1, dtype=cb.ASM_OPTS_DT)
    asm_opts[0] = opts["asmOpts"][0]
    loci, contigs, pseudo, cpool, spool = binding.assemble_raw(asm_opts, batch.readOff, batch.readViews, batch.pool)
    nLoci = batch.nLoci
    nContigs = loci["nContigs"].astype(np.uint32)
    slotFirst = np.concatenate([[0], np.cumsum(nContigs)[:-1]]).astype(np.uint32)
    nSlots = int(nContigs.sum())
    base = batch.pool.size
    pool = np.concatenate([batch.pool, cpool])
    contigSeq = np.zeros(nSlots, dtype=cb.SEQVIEW_DT)
    support = np.zeros(nSlots, dtype=np.uint32)
    locusOf = np.zeros(nSlots, dtype=np.int64)
    for li in range(nLoci):
        for k in range(int(nContigs[li])):
            c = contigs[int(loci["firstContig"][li]) + k]
            s = int(slotFirst[li]) + k
            contigSeq[s] = (base + int(c["seqOff"]), int(c["seqLen"]), 0)
            w = spool[int(c["supportOff"]):int(c["supportOff"]) + int(loci["setWords"][li])]
            support[s] = sum(bin(int(x)).count("1") for x in w)
            locusOf[s] = li
    cplx = np.nonzero(batch.isJump[locusOf] == 0)[0]
    jmp = np.nonzero(batch.isJump[locusOf] != 0)[0]
    refv = batch.ref1[locusOf].copy()
    trims = np.zeros(nSlots, dtype=cb.TRIM_RESULT_DT)
    if cuts is not None and len(cplx):
        trims[cplx] = binding.ref_trim_raw(pool, contigSeq[cplx], refv[cplx], cuts[locusOf[cplx]])
        refv["off"][cplx] += trims["adjLeadingCut"][cplx].astype(np.uint64)
        refv["len"][cplx] -= (trims["adjLeadingCut"][cplx] + trims["adjTrailingCut"][cplx]).astype(np.uint32)
    aln = np.zeros(nSlots, dtype=cb.ALN_RESULT_DT)
    jump = np.zeros(nSlots, dtype=cb.JUMP_RESULT_DT)
    cigs, cigBase = [], 0
    if len(cplx):
        S_ = cb.Scores(*[int(x) for x in opts["indelScores"][0]])
        r, cig = binding.single_align_raw("indel", S_, int(opts["largeIndelScore"][0]), pool, contigSeq[cplx], refv[cplx])
        r["beginPos"] += trims["adjLeadingCut"][cplx]
        aln[cplx] = r
        cigs.append(cig)
        cigBase = len(cig)
    if len(jmp):
        S_ = cb.Scores(*[int(x) for x in opts["jumpScores"][0]])
        r, cig = binding.jump_align_raw(S_, int(opts["jumpScore"][0]), pool, contigSeq[jmp], batch.ref1[locusOf[jmp]],
                                        batch.ref2[locusOf[jmp]])
        r["cigarOff1"] += cigBase
        r["cigarOff2"] += cigBase
        jump[jmp] = r
        cigs.append(cig)
    cigar = np.concatenate(cigs) if cigs else np.zeros(1, dtype=np.uint32)
    case = QcCase(batch.isJump.copy(), slotFirst, nContigs, aln, jump, cigar, pool, contigSeq, batch.ref1.copy(), support)
    case.asm = (loci, contigs, pseudo, cpool, spool)  # the assembler's own outputs, for tests that compare them too
    return case, trims


def unit_case(t, scores):
    """one transcribed reference unit test (tests/golden/refiner_qc.json "unit") -> QcCase + options"""
    import re
    code = {c: i for i, c in enumerate(cb.SEG_CODES)}

    def segs(cigar):
        return [(int(n) << 4) | code[c] for n, c in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]

    n = len(t["contigs"])
    aln = np.zeros(n, dtype=cb.ALN_RESULT_DT)
    jump = np.zeros(n, dtype=cb.JUMP_RESULT_DT)
    cig = []
    contig = np.frombuffer(t.get("contig", "ACGTACGTACGT").encode(), dtype=np.uint8)
    refseq = np.frombuffer(t.get("ref", "ACGT").encode(), dtype=np.uint8)
    pool = np.concatenate([contig, refseq]).copy()
    views = np.zeros(n, dtype=cb.SEQVIEW_DT)
    views["len"] = len(contig)
    support = np.zeros(n, dtype=np.uint32)
    for i, c in enumerate(t["contigs"]):
        if t["kind"] == "jump":
            s1, s2 = segs(c["cigar1"]), segs(c["cigar2"])
            jump[i] = (c["score"], 0, 0, 0, 0, len(s1), len(s2), 0, len(cig), len(cig) + len(s1))
            cig += s1 + s2
        else:
            s1 = segs(c["cigar"])
            aln[i] = (c["score"], c["beginPos"], c["isJumped"], len(s1), len(cig))
            cig += s1
            support[i] = c["support"]
    case = QcCase(np.asarray([1 if t["kind"] == "jump" else 0], np.uint8), np.zeros(1, np.uint32),
                           np.asarray([n], np.uint32), aln, jump, np.asarray(cig + [0], dtype=np.uint32), pool, views,
                           np.asarray([(len(contig), len(refseq), 0)], dtype=cb.SEQVIEW_DT), support)
    o = cb.default_qc_opts()
    o["filterScores"] = tuple(scores)
    o["minCandidateVariantSize"] = t["minSize"]
    return case, o



# ---------------------------------------------------------------------------------------------------------------------
def _rand_path(rng, flavour):
    """list of (type, len); flavour steers the corner being exercised"""
    segs = []

    def match_run(lo, hi):
        total = int(rng.integers(lo, hi))
        while total > 0:
            ln = min(total, int(rng.integers(1, 60)))
            segs.append((EQ, ln))
            total -= ln
            if total > 0 and rng.random() < 0.5:
                segs.append((X, int(rng.integers(1, 4))))

    if rng.random() < 0.3:
        segs.append((S, int(rng.integers(1, 40))))
    n_ev = {"none": 0, "one": 1, "many": int(rng.integers(2, 6))}[flavour]
    match_run(5, 260)
    for _ in range(n_ev):
        for _ in range(int(rng.integers(1, 4))):  # adjacent insert/delete runs form one 'complex' segment
            segs.append((I if rng.random() < 0.5 else D, int(rng.integers(1, 80))))
        match_run(5, 260)
        if rng.random() < 0.3:  # a small indel that belongs to no candidate segment
            segs.append((I if rng.random() < 0.5 else D, int(rng.integers(1, 9))))
            match_run(5, 120)
    if rng.random() < 0.3:
        segs.append((S, int(rng.integers(1, 40))))
    # merge equal neighbours (the aligners never emit them)
    out = []
    for t, ln in segs:
        if out and out[-1][0] == t:
            out[-1] = (t, out[-1][1] + ln)
        else:
            out.append((t, ln))
    return out


def _apply_path(rng, path, lead):
    """contig and reference consistent with the path (reference has `lead` random bases before beginPos)"""
    contig, ref = [], [ACGT[rng.integers(0, 4, lead)]]
    for t, ln in path:
        if t == EQ:
            b = ACGT[rng.integers(0, 4, ln)]
            contig.append(b)
            ref.append(b)
        elif t == X:
            b = rng.integers(0, 4, ln)
            contig.append(ACGT[b])
            ref.append(ACGT[(b + 1 + rng.integers(0, 3, ln)) % 4])
        elif t in (I, S):
            contig.append(ACGT[rng.integers(0, 4, ln)])
        elif t == D:
            ref.append(ACGT[rng.integers(0, 4, ln)])
    ref.append(ACGT[rng.integers(0, 4, int(rng.integers(0, 300)))])
    return np.concatenate(contig), np.concatenate(ref)


def fuzz_case(seed, n_loci=200, exotic=False):
    """exotic=True sprinkles hard clips at the path ends and reference skips ('N') inside jump-arm paths -- segment
    types the aligners never emit but the path helpers handle (CPU-only comparisons use it)"""
    rng = np.random.default_rng(seed)
    isJump = (rng.random(n_loci) < 0.4).astype(np.uint8)
    nContigs = rng.integers(0, 6, n_loci).astype(np.uint32)
    slotFirst = np.concatenate([[0], np.cumsum(nContigs)[:-1]]).astype(np.uint32)
    nSlots = int(nContigs.sum())
    aln = np.zeros(nSlots, dtype=cb.ALN_RESULT_DT)
    jump = np.zeros(nSlots, dtype=cb.JUMP_RESULT_DT)
    contigSeq = np.zeros(nSlots, dtype=cb.SEQVIEW_DT)
    ref1 = np.zeros(n_loci, dtype=cb.SEQVIEW_DT)
    support = rng.integers(0, 40, nSlots).astype(np.uint32)
    chunks, cig, off = [], [], 0

    def put(arr):
        nonlocal off
        chunks.append(arr)
        o = off
        off += len(arr)
        return o

    def put_path(path):
        o = len(cig)
        cig.extend((ln << 4) | t for t, ln in path)
        return o, len(path)

    for li in range(n_loci):
        s0 = int(slotFirst[li])
        if isJump[li]:
            ref1[li] = (put(ACGT[rng.integers(0, 4, 50)]), 50, 0)
            for k in range(int(nContigs[li])):
                s = s0 + k
                p1 = _rand_path(rng, "none") if rng.random() < 0.9 else []
                p2 = _rand_path(rng, "none") if rng.random() < 0.9 else []
                if exotic:
                    for p in (p1, p2):
                        if p and rng.random() < 0.5:
                            p.append((6, int(rng.integers(1, 20))))        # trailing hard clip
                        if p and rng.random() < 0.5:
                            p.insert(0, (6, int(rng.integers(1, 20))))     # leading hard clip
                        if len(p) > 2 and rng.random() < 0.4:
                            p.insert(len(p) // 2, (4, int(rng.integers(1, 300))))  # skip
                if rng.random() < 0.3:  # short arm
                    p1 = [(EQ, int(rng.integers(1, 40)))]
                o1, n1 = put_path(p1)
                o2, n2 = put_path(p2)
                jump[s] = (int(rng.integers(-50, 600)), 0, 0, 0, 0, n1, n2, 0, o1, o2)
                contigSeq[s] = (put(ACGT[rng.integers(0, 4, 12)]), 12, 0)
            continue
        # complex locus: all contigs share one reference; each contig's alignment sits at its own beginPos inside it
        refs = []
        for k in range(int(nContigs[li])):
            s = s0 + k
            path = _rand_path(rng, rng.choice(["none", "one", "one", "many"]))
            lead = int(rng.integers(0, 400))
            contig, ref = _apply_path(rng, path, lead)
            if rng.random() < 0.25 and lead > 0:
                # plant a second copy of the left flank upstream of the alignment end -> ambiguity filter
                left = 0
                for t, ln in path:
                    if t in (I, D):
                        break
                    if t != D:
                        left += ln
                if 0 < left <= lead:
                    ref[lead - left:lead] = contig[:left]
            o, n = put_path(path)
            aln[s] = (int(rng.integers(-50, 600)), 0, int(rng.random() < 0.3), n, o)
            contigSeq[s] = (put(contig), len(contig), 0)
            refs.append((s, lead, ref))
        # concatenate the per-contig references into the locus window and shift each beginPos
        pos, parts = 0, []
        for s, lead, ref in refs:
            aln["beginPos"][s] = pos + lead
            parts.append(ref)
            pos += len(ref)
        window = np.concatenate(parts) if parts else ACGT[rng.integers(0, 4, 40)]
        ref1[li] = (put(window), len(window), 0)
    pool = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.uint8)
    cigar = np.asarray(cig if cig else [0], dtype=np.uint32)
    return QcCase(isJump, slotFirst, nContigs, aln, jump, cigar, pool, contigSeq, ref1, support)


def trim_fuzz(seed, n=300):
    """(pool, contig views, ref views, cuts): windows of assorted sizes, contigs cut out of them (so 10-mers are
    shared), some with no shared 10-mer at all, some windows shorter than the cuts"""
    rng = np.random.default_rng(seed)
    chunks, off = [], 0
    contig = np.zeros(n, dtype=cb.SEQVIEW_DT)
    ref = np.zeros(n, dtype=cb.SEQVIEW_DT)
    cuts = np.zeros(n, dtype=cb.REF_CUTS_DT)
    for i in range(n):
        rl = int(rng.choice([15, 40, 300, 900, 2000, 3000]))
        r = ACGT[rng.integers(0, 4, rl)]
        cl = int(rng.integers(10, 400))
        if rng.random() < 0.8 and rl > cl + 2:
            st = int(rng.integers(0, rl - cl))
            c = r[st:st + cl].copy()
            if rng.random() < 0.5:  # a few substitutions
                idx = rng.integers(0, cl, max(1, cl // 30))
                c[idx] = ACGT[rng.integers(0, 4, len(idx))]
            if rng.random() < 0.2:
                c[rng.integers(0, cl)] = ord("N")
        else:
            c = ACGT[rng.integers(0, 4, cl)]
        if rng.random() < 0.1:
            r = r.copy()
            r[rng.integers(0, rl, 3)] = ord("N")
        lead = int(rng.choice([0, 20, 100]))
        trail = int(rng.choice([0, 20, 100]))
        cuts[i] = (lead, trail, lead + int(rng.choice([0, 200, 700])), trail + int(rng.choice([0, 200, 700])))
        contig[i] = (off, cl, 0)
        chunks.append(c)
        off += cl
        ref[i] = (off, rl, 0)
        chunks.append(r)
        off += rl
    return np.concatenate(chunks), contig, ref, cuts


def jump_sv_inputs(case, seed):
    """orientation flags, window sizes and alignment offsets for the spanning candidates of a QcCase (mutates the
    jump records: beginPos1/2, jumpRange, jumpInsertSize).  Returns (ref1 views, ref2 views, orient)."""
    rng = np.random.default_rng(seed)
    n = len(case.isJump)
    ref1 = np.zeros(n, dtype=cb.SEQVIEW_DT)
    ref2 = np.zeros(n, dtype=cb.SEQVIEW_DT)
    ref1["len"] = rng.integers(1500, 6000, n)
    ref2["len"] = rng.integers(1500, 6000, n)
    orient = np.zeros(n, dtype=cb.JUMP_ORIENT_DT)
    for k in ("isBp2AlignedFirst", "isBp1Reversed", "isBp2Reversed", "bp1RightOpen", "bp2RightOpen"):
        orient[k] = rng.integers(0, 2, n)
    m = len(case.jump)
    case.jump["beginPos1"] = rng.integers(0, 1000, m)
    case.jump["beginPos2"] = rng.integers(0, 1000, m)
    case.jump["jumpRange"] = rng.integers(0, 30, m) * (rng.random(m) < 0.5)
    case.jump["jumpInsertSize"] = rng.integers(0, 40, m) * (rng.random(m) < 0.3)
    return ref1, ref2, orient
