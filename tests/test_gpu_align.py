"""CUDA aligners (through the C ABI) against the oracle and the golden vectors."""
import json
import os
import random

import pytest

from manta_b200 import _cabi as cb
from manta_b200 import synth
from test_oracle_golden import load, lookup, run_alignment

pytestmark = pytest.mark.gpu

SCORE_SETS = [((2, -4, -5, -1, -1, 0), -3), ((2, -8, -12, -1, -1, 0), -100), ((2, -8, -24, -1, -1, 0), -100),
              ((2, -8, -18, 0, -1, 1), -50), ((1, -4, -5, -1, -2, 1), -10)]


@pytest.fixture(params=["tag16", "pair16", "int32"])
def dev(request, gpu, gpusel, gpu32):
    return {"tag16": gpu, "pair16": gpusel, "int32": gpu32}[request.param]


def test_alignment_golden_gpu(dev):
    gpu = dev
    for c in load("alignment.json")["cases"]:
        got = run_alignment(gpu, c)
        for k, v in c["expected"].items():
            assert lookup(got, k) == v, (c["test"], k)
        assert got == c["reference"], c["test"]


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_aligners_vs_oracle(dev, orc, seed):
    gpu = dev
    for si, (sc, extra) in enumerate(SCORE_SETS):
        S = cb.Scores(*sc)
        if not sc[5]:
            cases = synth.random_alignment_cases(seed * 100 + si, 40, "jump")
            assert gpu.jump_align(S, extra, cases) == orc.jump_align(S, extra, cases)
        cases = synth.random_alignment_cases(seed * 100 + si + 50, 40, "indel")
        assert gpu.single_align("indel", S, extra, cases) == orc.single_align("indel", S, extra, cases)
        assert gpu.single_align("global", S, 0, cases) == orc.single_align("global", S, 0, cases)


def test_long_sequences_multi_tile_multi_pass(dev, orc):
    gpu = dev
    """reference windows spanning several 2 KiB staging tiles, queries spanning several 512-row passes,
    and the q > 1250 regime where the -10000 sentinel takes part in the arithmetic (SURVEY.md A.4)"""
    rng = random.Random(5)

    def rs(n):
        return "".join(rng.choice("ACGT") for _ in range(n))
    ref = rs(5000)
    ref2 = rs(4500)
    contig = ref[1000:1200] + ref[2900:3100]           # 1.7 kb deletion
    jump = ref[2000:2250] + "ACGTTGCA" + ref2[3000:3200]
    longq = ref[100:900] + rs(40) + ref[1500:2400]       # 1740 rows: 4 passes of 512
    S = cb.Scores(2, -8, -24, -1, -1, 0)
    pairs = [(contig, ref), (longq, ref), (rs(1600), rs(300)), (contig, ref[:2047]), (contig, ref[:2048]), (contig, ref[:2049])]
    assert gpu.single_align("indel", S, -100, pairs) == orc.single_align("indel", S, -100, pairs)
    assert gpu.single_align("global", cb.Scores(2, -8, -12, -1, -1, 1), 0, pairs) == \
        orc.single_align("global", cb.Scores(2, -8, -12, -1, -1, 1), 0, pairs)
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    triples = [(jump, ref, ref2), (longq, ref[:2100], ref2[:2300]), (rs(1600), rs(300), rs(200)),
               (jump, ref[:2041], ref2), (jump, ref[1:2049], ref2[3:])]
    assert gpu.jump_align(S, -100, triples) == orc.jump_align(S, -100, triples)


def test_unaligned_views_and_shared_pool(dev, orc):
    gpu = dev
    """views at arbitrary byte offsets of one shared pool (no 16-byte alignment of the caller's data)"""
    import numpy as np
    rng = random.Random(9)
    blob = "".join(rng.choice("ACGT") for _ in range(6000))
    pool = np.frombuffer((blob + "\0" * 8).encode(), dtype=np.uint8).copy()
    n = 24
    q = np.zeros(n, dtype=cb.SEQVIEW_DT)
    r1 = np.zeros(n, dtype=cb.SEQVIEW_DT)
    r2 = np.zeros(n, dtype=cb.SEQVIEW_DT)
    for i in range(n):
        q["off"][i], q["len"][i] = rng.randint(0, 3000), rng.randint(1, 300)
        r1["off"][i], r1["len"][i] = rng.randint(0, 3000), rng.randint(1, 2500)
        r2["off"][i], r2["len"][i] = rng.randint(0, 3000), rng.randint(1, 2500)
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    a, ca = gpu.jump_align_raw(S, -100, pool, q, r1, r2)
    b, cbb = orc.jump_align_raw(S, -100, pool, q, r1, r2)
    assert [cb.decode_jump(a[i], ca) for i in range(n)] == [cb.decode_jump(b[i], cbb) for i in range(n)]
    a, ca = gpu.single_align_raw("indel", cb.Scores(2, -8, -24, -1, -1, 0), -100, pool, q, r1)
    b, cbb = orc.single_align_raw("indel", cb.Scores(2, -8, -24, -1, -1, 0), -100, pool, q, r1)
    assert [cb.decode_aln(a[i], ca) for i in range(n)] == [cb.decode_aln(b[i], cbb) for i in range(n)]


def test_empty_input_errors_gpu(gpu):
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    for triple, st in [(("", "A", "A"), cb.MB200_ERR_EMPTY_QUERY), (("A", "", "A"), cb.MB200_ERR_EMPTY_REF1),
                       (("A", "A", ""), cb.MB200_ERR_EMPTY_REF2)]:
        with pytest.raises(cb.Mb200Error) as e:
            gpu.jump_align(S, -100, [triple])
        assert e.value.status == st
        assert str(e.value) == cb.ERROR_MESSAGES[st]
    with pytest.raises(cb.Mb200Error) as e:
        gpu.single_align("global", S, 0, [("A", "")])
    assert e.value.status == cb.MB200_ERR_EMPTY_REF


def test_pair_kernel_batches(gpu, orc):
    """many tasks of mixed sizes in one call: exercises pairing (odd counts, jump pairs that must share R1,
    pairs with different Q / R, all four row classes)"""
    rng = random.Random(17)

    def rs(n):
        return "".join(rng.choice("ACGT") for _ in range(n))
    refs = [rs(rng.choice([700, 1500, 2500, 2500, 3100])) for _ in range(6)]
    S = cb.Scores(2, -8, -24, -1, -1, 0)
    pairs = []
    for i in range(41):
        r = rng.choice(refs)
        a = rng.randint(0, len(r) - 400)
        q = r[a:a + rng.randint(20, 200)] + r[a + 250:a + 250 + rng.randint(10, 230)]
        pairs.append((q, r))
    assert gpu.single_align("indel", S, -100, pairs) == orc.single_align("indel", S, -100, pairs)
    assert gpu.single_align("global", S, 0, pairs) == orc.single_align("global", S, 0, pairs)
    S = cb.Scores(2, -8, -12, -1, -1, 0)
    triples = []
    for i in range(37):
        r1, r2 = rng.choice(refs), rng.choice(refs)
        a, b = rng.randint(0, len(r1) - 300), rng.randint(0, len(r2) - 300)
        triples.append((r1[a:a + rng.randint(30, 250)] + r2[b:b + rng.randint(30, 250)], r1, r2))
    assert gpu.jump_align(S, -100, triples) == orc.jump_align(S, -100, triples)
