"""CUDA iterative assembler (through the C ABI) against the golden vectors and the oracle -- every test once per
assembler kernel (fixture `gpuasm`: global-memory tables / on-chip tables)."""
import random

import numpy as np
import pytest

from manta_b200 import _cabi as cb
from manta_b200 import synth
from test_oracle_golden import asm_to_json, load

pytestmark = pytest.mark.gpu


def acgtn_only(reads):
    return all(set(r) <= set("ACGTN") for r in reads)


def test_assembly_golden_gpu(gpuasm):
    n = 0
    for c in load("assembly.json")["cases"]:
        opts = cb.default_asm_opts(**c["opts"])
        if not acgtn_only(c["reads"]):
            with pytest.raises(cb.Mb200Error) as e:
                gpuasm.assemble(opts, [c["reads"]])
            assert e.value.status == cb.MB200_ERR_BAD_BASE
            continue
        got = asm_to_json(gpuasm.assemble(opts, [c["reads"]])[0])
        assert got == c["reference"], c["test"]
        n += 1
    assert n >= 25


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_assembler_vs_oracle(gpuasm, orc, seed):
    cases = synth.random_assembly_cases(seed, 40)
    # batched by identical options
    groups = {}
    for opts, reads in cases:
        groups.setdefault(tuple(sorted(opts.items())), []).append(reads)
    for key, loci in groups.items():
        o = cb.default_asm_opts(**dict(key))
        assert gpuasm.assemble(o, loci) == orc.assemble(o, loci), key


def test_assembler_repeat_heavy(gpuasm, orc):
    """tandem repeats of every period 1..30: homopolymers, small and large circles, the order-dependent DFS"""
    rng = random.Random(77)
    loci = []
    for per in range(1, 31):
        loci.append(synth.random_assembly_reads(rng, 48, 120, with_repeat=True))
    for kw in (dict(), dict(minWordLength=12, maxWordLength=32, wordStepSize=4), dict(minWordLength=5, maxWordLength=9, wordStepSize=2)):
        o = cb.default_asm_opts(**kw)
        assert gpuasm.assemble(o, loci) == orc.assemble(o, loci), kw


def test_assembler_edge_inputs(gpuasm, orc):
    o = cb.default_asm_opts()
    loci = [[], ["ACGT"], ["N" * 100] * 5, ["A" * 150] * 10, ["ACGT" * 40] * 6,
            ["".join(random.Random(3).choice("ACGT") for _ in range(150))] * 70]
    assert gpuasm.assemble(o, loci) == orc.assemble(o, loci)
    # many reads in one locus (several 32-bit words of read-set bits), k up to 96
    rng = random.Random(4)
    big = synth.random_assembly_reads(rng, 300, 150, with_repeat=False)
    o2 = cb.default_asm_opts(minWordLength=81, maxWordLength=96, wordStepSize=5)
    assert gpuasm.assemble(o, [big]) == orc.assemble(o, [big])
    assert gpuasm.assemble(o2, [big[:80]]) == orc.assemble(o2, [big[:80]])


def test_assembler_config2_shape(gpuasm, orc):
    """BASELINE.json configs[1] shape: 64 reads x 150 bp per locus (small batch of the bench generator)"""
    b = synth.make_candidate_batch(96, seed=20260923)
    o = cb.default_asm_opts()
    g = gpuasm.assemble_raw(o, b.readOff, b.readViews, b.pool)
    r = orc.assemble_raw(o, b.readOff, b.readViews, b.pool)
    for i in range(b.nLoci):
        assert cb.decode_locus(g[0][i], *g[1:], b.readsPerLocus) == cb.decode_locus(r[0][i], *r[1:], b.readsPerLocus), i
