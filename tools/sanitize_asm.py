"""compute-sanitizer driver for the two assembler kernels (repeat-heavy, degenerate and bench-shaped loci):
    compute-sanitizer --tool memcheck  python tools/sanitize_asm.py onchip|global
    compute-sanitizer --tool racecheck python tools/sanitize_asm.py onchip
"""
import random, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import manta_b200
from manta_b200 import _cabi as cb, synth
import ctypes as C, os
mode = sys.argv[1]
ctx = manta_b200.Context(0, assembler=mode)
rng = random.Random(77)
loci = [synth.random_assembly_reads(rng, 48, 120, with_repeat=True) for _ in range(6)]
loci += [[], ["ACGT"], ["N" * 100] * 5, ["A" * 150] * 10, ["ACGT" * 40] * 6]
o = cb.default_asm_opts()
r = ctx.assemble(o, loci)
b = synth.make_candidate_batch(12, seed=20260923)
g = ctx.assemble_raw(o, b.readOff, b.readViews, b.pool)
print("ok", mode, len(r), sum(len(x["contigs"]) if isinstance(x, dict) and "contigs" in x else 0 for x in r))
ctx.close()
