"""Fused candidate path (assemble -> align every contig, contigs resident on the device) against the oracle
composition of the same reference calls."""
import numpy as np
import pytest

from manta_b200 import _cabi as cb
from manta_b200 import synth

pytestmark = pytest.mark.gpu


def oracle_candidates(orc, batch, opts):
    """runIterativeAssembler, then GlobalLargeIndelAligner / GlobalJumpAligner on every contig (oracle, CPU)"""
    asm_opts = np.zeros(1, dtype=cb.ASM_OPTS_DT)
    asm_opts[0] = opts["asmOpts"][0]
    loci, contigs, pseudo, cpool, spool = orc.assemble_raw(asm_opts, batch.readOff, batch.readViews, batch.pool)
    pool = batch.pool
    out = []
    for li in range(batch.nLoci):
        cs, ps, info = cb.decode_locus(loci[li], contigs, pseudo, cpool, spool, batch.readsPerLocus)
        r1 = bytes(pool[int(batch.ref1["off"][li]):int(batch.ref1["off"][li]) + int(batch.ref1["len"][li])])
        alns = []
        if batch.isJump[li]:
            r2 = bytes(pool[int(batch.ref2["off"][li]):int(batch.ref2["off"][li]) + int(batch.ref2["len"][li])])
            if cs:
                alns = orc.jump_align(cb.Scores(*[int(x) for x in opts["jumpScores"][0]]), int(opts["jumpScore"][0]),
                                      [(c.seq, r1, r2) for c in cs])
        elif cs:
            alns = orc.single_align("indel", cb.Scores(*[int(x) for x in opts["indelScores"][0]]),
                                    int(opts["largeIndelScore"][0]), [(c.seq, r1) for c in cs])
        out.append((cs, ps, alns))
    return out


def decode_candidates(batch, opts, out):
    A = int(opts["asmOpts"]["maxAssemblyCount"][0])
    res = []
    for li in range(batch.nLoci):
        cs, ps, info = cb.decode_locus(out["loci"][li], out["contigs"], out["pseudo"], out["contigPool"], out["setPool"],
                                       batch.readsPerLocus)
        alns = []
        for k in range(len(cs)):
            slot = li * A + k
            if batch.isJump[li]:
                alns.append(cb.decode_jump(out["jump"][slot], out["cigar"]))
            else:
                alns.append(cb.decode_aln(out["aln"][slot], out["cigar"]))
        res.append((cs, ps, alns))
    return res


def test_candidates_vs_oracle(gpu, orc):
    batch = synth.make_candidate_batch(48, seed=7, window=1500)
    assert 0 < int(batch.isJump.sum()) < batch.nLoci
    opts = cb.default_candidate_opts()
    gpu.candidates_upload(batch)
    out = gpu.candidates_run(opts)
    got = decode_candidates(batch, opts, out)
    want = oracle_candidates(orc, batch, opts)
    for li in range(batch.nLoci):
        assert got[li] == want[li], li
    t = gpu.candidates_timing()
    assert t["nContigs"] == sum(len(w[0]) for w in want) and t["launches"] >= 2
    # second run on the resident batch (reusing host buffers) gives identical results
    out2 = gpu.candidates_run(opts, out)
    assert decode_candidates(batch, opts, out2) == want


def test_candidates_full_window_properties(gpu):
    """BASELINE config-2 sized windows (5 kb): size-independent properties of every alignment"""
    batch = synth.make_candidate_batch(64, seed=20260923)
    opts = cb.default_candidate_opts()
    gpu.candidates_upload(batch)
    out = gpu.candidates_run(opts)
    A = int(opts["asmOpts"]["maxAssemblyCount"][0])
    n = 0
    for li in range(batch.nLoci):
        for k in range(int(out["loci"]["nContigs"][li])):
            slot = li * A + k
            qlen = int(out["contigs"]["seqLen"][slot])
            if batch.isJump[li]:
                r = cb.decode_jump(out["jump"][slot], out["cigar"])
                cig = r.align1.cigar + r.align2.cigar
                extra = r.jumpInsertSize
            else:
                r = cb.decode_aln(out["aln"][slot], out["cigar"])
                cig, extra = r.align.cigar, 0
            import re
            read_len = sum(int(a) for a, op in re.findall(r"(\d+)([=XIS])", cig)) + extra
            assert read_len == qlen, (li, k, cig)       # every query base is accounted for exactly once
            assert r.score <= 2 * qlen                   # cannot beat a perfect match
            n += 1
    assert n > 64


def test_candidates_checksum_full_window(gpu, orc):
    """BASELINE config-2 sized loci: the sum of all alignment scores and the contig count must equal the CPU
    composition (oracle, multi-threaded driver) -- a checksum of checksums over the whole batch"""
    batch = synth.make_candidate_batch(160, seed=99)
    opts = cb.default_candidate_opts()
    gpu.candidates_upload(batch)
    out = gpu.candidates_run(opts)
    A = int(opts["asmOpts"]["maxAssemblyCount"][0])
    total, ncontig = 0, 0
    for li in range(batch.nLoci):
        for k in range(int(out["loci"]["nContigs"][li])):
            total += int(out["jump"]["score"][li * A + k]) if batch.isJump[li] else int(out["aln"]["score"][li * A + k])
            ncontig += 1
    nc, cells, ssum, _ = orc.candidates_mt(opts, batch, batch.nLoci, 8)
    assert (ncontig, total) == (nc, ssum)
    t = gpu.candidates_timing()
    assert t["indelCells"] + t["jumpCells"] == cells


def test_candidates_other_word_length(gpu, orc):
    """a run whose first word length differs from the one the staged batch was scanned for (re-scan path), followed
    by a second upload that is scanned for the new word length directly"""
    batch = synth.make_candidate_batch(24, seed=5, window=1500)
    opts = cb.default_candidate_opts(minWordLength=46, maxWordLength=66)
    gpu.candidates_upload(batch)
    want = oracle_candidates(orc, batch, opts)
    assert decode_candidates(batch, opts, gpu.candidates_run(opts)) == want
    gpu.candidates_upload(batch)
    assert decode_candidates(batch, opts, gpu.candidates_run(opts)) == want
    opts = cb.default_candidate_opts()
    assert decode_candidates(batch, opts, gpu.candidates_run(opts)) == oracle_candidates(orc, batch, opts)


def test_candidates_staging_queue(gpu, orc):
    """two staged batches are consumed first-in first-out; with nothing staged a run repeats on the last batch; a
    third upload before any run is refused"""
    opts = cb.default_candidate_opts()
    b1 = synth.make_candidate_batch(16, seed=11, window=1500)
    b2 = synth.make_candidate_batch(20, seed=12, window=1500)
    w1, w2 = oracle_candidates(orc, b1, opts), oracle_candidates(orc, b2, opts)
    gpu.candidates_upload(b1)
    gpu.candidates_upload(b2)
    with pytest.raises(cb.Mb200Error):
        gpu.candidates_upload(b1)
    # the binding sizes the host arrays from the batch the run CONSUMES (b1: 16 loci), not from the newest upload
    # (b2: 20 loci), and a reused `out` of a smaller batch is replaced instead of being overrun
    out1 = gpu.candidates_run(opts)
    assert len(out1["loci"]) == b1.nLoci
    assert decode_candidates(b1, opts, out1) == w1
    gpu.candidates_upload(b1)  # overwrites the slot of the batch just consumed while b2 is still staged
    out2 = gpu.candidates_run(opts, out=out1)
    assert len(out2["loci"]) == b2.nLoci
    assert decode_candidates(b2, opts, out2) == w2
    assert decode_candidates(b1, opts, gpu.candidates_run(opts)) == w1
    assert decode_candidates(b1, opts, gpu.candidates_run(opts)) == w1  # nothing staged: repeat on the resident batch
