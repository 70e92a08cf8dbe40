"""Full-record parity of the fused candidate path AT THE MEASURED SIZE (BASELINE configs[1]: 64 reads x 150 bp, 5 kb
windows, the refiner's default window cuts): every field the library returns -- contigs (sequence, conservative range,
support / reject read sets), pseudo reads, window trims, large-indel and jump alignment records with their CIGARs, the
QC records, bestAlignmentIndex, the refined small-SV records and the refined spanning breakends -- against the oracle's
composition of the same reference steps on the same seeded loci.  The oracle side runs in 64-locus chunks on host threads
(ctypes releases the GIL)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import qc_cases
from manta_b200 import _cabi as cb
from manta_b200 import synth

pytestmark = pytest.mark.gpu

N_LOCI = int(os.environ.get("MB200_SCALE_LOCI", "2048"))
CHUNK = 64


def _sub(batch, a, b):
    """loci [a, b) of a batch (views keep pointing into the shared pool)"""
    r0, r1 = int(batch.readOff[a]), int(batch.readOff[b])
    return synth.CandidateBatch(nLoci=b - a, readsPerLocus=batch.readsPerLocus, readLen=batch.readLen, pool=batch.pool,
                                readOff=(batch.readOff[a:b + 1] - batch.readOff[a]).astype(np.uint64),
                                readViews=batch.readViews[r0:r1].copy(), isJump=batch.isJump[a:b].copy(),
                                ref1=batch.ref1[a:b].copy(), ref2=batch.ref2[a:b].copy())


def _cig(pool, off, n):
    return pool[int(off):int(off) + int(n)].tolist()


def test_full_records_at_config_size(gpuasm, orc):
    gpu = gpuasm  # once per assembler kernel (2 048 loci: the default would take the global-memory kernel)
    batch = synth.make_candidate_batch(N_LOCI, seed=20260923)   # the bench's generator and seed, 5 kb windows
    opts = cb.default_candidate_opts()
    cuts = cb.default_ref_cuts(batch.nLoci)
    A = int(opts["asmOpts"]["maxAssemblyCount"][0])
    rng = np.random.default_rng(8)
    orient = np.zeros(batch.nLoci, dtype=cb.JUMP_ORIENT_DT)
    for k in ("isBp2AlignedFirst", "isBp1Reversed", "isBp2Reversed", "bp1RightOpen", "bp2RightOpen"):
        orient[k] = rng.integers(0, 2, batch.nLoci)

    gpu.candidates_upload(batch, cuts)
    out = gpu.candidates_run(opts, select=True)
    svFirst, svs = gpu.candidates_small_svs()
    jsv = gpu.candidates_jump_svs(orient)

    def oracle_chunk(a):
        b = min(a + CHUNK, batch.nLoci)
        sub = _sub(batch, a, b)
        case, trims = qc_cases.pipeline_case(orc, sub, opts, cuts[a:b])
        qc, best = case.run(orc)
        f, s = case.small_svs(orc, qc, best)
        j = orc.jump_sv_raw(case.isJump, case.slotFirst, len(case.jump), best, case.jump, case.cigar, sub.ref1, sub.ref2,
                            orient[a:b])
        return a, b, case, trims, qc, best, f, s, j

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        chunks = list(ex.map(oracle_chunk, range(0, batch.nLoci, CHUNK)))

    nAln = nJump = nSv = nJsv = nSel = 0
    for a, b, case, trims, qc, best, f, s, j in chunks:
        loci, contigs, pseudo, cpool, spool = case.asm
        assert out["loci"]["nContigs"][a:b].tolist() == case.nContigs.tolist(), a
        assert out["best"][a:b].tolist() == best.tolist(), a
        nSel += int((best >= 0).sum())
        for li in range(a, b):
            # the assembler's outputs: contig strings, seed counts, read sets, conservative ranges, pseudo reads
            got = cb.decode_locus(out["loci"][li], out["contigs"], out["pseudo"], out["contigPool"], out["setPool"],
                                  batch.readsPerLocus)
            want = cb.decode_locus(loci[li - a], contigs, pseudo, cpool, spool, batch.readsPerLocus)
            assert got == want, li
            assert int(out["loci"]["lastWordLength"][li]) == int(loci["lastWordLength"][li - a]), li
            n = int(case.nContigs[li - a])
            s0 = int(case.slotFirst[li - a])
            for k in range(n):
                g, o = li * A + k, s0 + k
                assert out["qc"][g].tobytes() == qc[o].tobytes(), (li, k)
                assert out["trim"][g].tobytes() == trims[o].tobytes(), (li, k)
                if batch.isJump[li]:
                    ga, oa = out["jump"][g], case.jump[o]
                    for fld in ("score", "jumpInsertSize", "jumpRange", "beginPos1", "beginPos2", "cigarLen1", "cigarLen2"):
                        assert int(ga[fld]) == int(oa[fld]), (li, k, fld)
                    assert _cig(out["cigar"], ga["cigarOff1"], ga["cigarLen1"]) == _cig(case.cigar, oa["cigarOff1"], oa["cigarLen1"]), (li, k)
                    assert _cig(out["cigar"], ga["cigarOff2"], ga["cigarLen2"]) == _cig(case.cigar, oa["cigarOff2"], oa["cigarLen2"]), (li, k)
                    nJump += 1
                else:
                    ga, oa = out["aln"][g], case.aln[o]
                    for fld in ("score", "beginPos", "isJumped", "cigarLen"):
                        assert int(ga[fld]) == int(oa[fld]), (li, k, fld)
                    assert _cig(out["cigar"], ga["cigarOff"], ga["cigarLen"]) == _cig(case.cigar, oa["cigarOff"], oa["cigarLen"]), (li, k)
                    nAln += 1
        # refined small SVs (candidate order) and refined spanning breakends of the chunk
        assert (svFirst[a:b + 1] - svFirst[a]).tolist() == f.tolist(), a
        s = s.copy()
        s["locus"] += a  # the chunk's candidate indices are relative to the chunk
        assert svs[int(svFirst[a]):int(svFirst[b])].tobytes() == s.tobytes(), a
        assert jsv[a:b].tobytes() == j.tobytes(), a
        nSv += len(s)
        nJsv += int(j["valid"].sum())
    # the batch exercises every record type in numbers
    assert nAln > N_LOCI and nJump > N_LOCI // 2 and nSel > N_LOCI // 2 and nSv > N_LOCI // 8 and nJsv > N_LOCI // 16


def test_pair_kernels_sentinel_and_uneven_pairs(gpu, gpusel, orc):
    """the int16x2 pair kernels on (i) a batch whose longest query puts it in the regime where the -10000 sentinel takes
    part in the arithmetic (no tagged kernel, NOBAD=false) and (ii) pairs whose two alignments differ maximally in query
    and reference length (the shorter one idles through most of the pair's steps)"""
    import random
    rng = random.Random(3)

    def rs(n):
        return "".join(rng.choice("ACGT") for _ in range(n))
    ref = rs(5000)
    S = cb.Scores(2, -8, -24, -1, -1, 0)
    long_q = ref[200:900] + rs(60) + ref[1500:2100]          # 1360 rows: L*(Q+1) > 10000 -> sentinel regime
    pairs = [(long_q, ref), (ref[100:140], ref[50:400]), (rs(1300), rs(200)), (ref[3000:3030], ref), (long_q[:1290], ref[:700])]
    want = orc.single_align("indel", S, -100, pairs)
    for dev in (gpu, gpusel):
        assert dev.single_align("indel", S, -100, pairs) == want
    # same class (strip of 16 rows per lane), maximally different members: Q 385 vs 512, R 60 vs 5000
    q_small, q_big = ref[1000:1385], ref[2000:2300] + rs(12) + ref[3800:4000]
    pairs = [(q_big, ref), (q_small, ref[980:1040]), (q_big[:400], ref[1900:2500]), (q_small, ref)]
    want = orc.single_align("indel", S, -100, pairs)
    for dev in (gpu, gpusel):
        assert dev.single_align("indel", S, -100, pairs) == want
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    triples = [(q_big, ref[:2500], ref[2500:]), (q_small, ref[:2500], ref[2500:2560]), (long_q, ref[:2500], ref[2500:]),
               (ref[2400:2440], ref[:2500], ref[2500:])]
    want = orc.jump_align(S, -100, triples)
    for dev in (gpu, gpusel):
        assert dev.jump_align(S, -100, triples) == want
