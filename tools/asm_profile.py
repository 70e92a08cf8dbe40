#!/usr/bin/env python
"""Per-phase clock totals of assembleKernel on N config-2 loci (instrumented build, -DMB200_ASM_PROFILE).

    python tools/asm_profile.py --build            # here (no GPU): compiles build/prof/libmanta_b200_prof.so
    python tools/asm_profile.py --loci 6000        # on the GPU box
"""
import argparse
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "tools", "_prof", "libmanta_b200_prof.so")  # build/ does not travel to the GPU box
ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--loci", type=int, default=6000)
ap.add_argument("--seed", type=int, default=20260923)
args = ap.parse_args()
if args.build:
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    os.makedirs(os.path.dirname(PROF), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(g.CSRC, "*.cu")))
    objs = []
    for s in srcs:
        o = os.path.join(ROOT, "build", "obj", os.path.basename(s)[:-3] + ".o")
        if os.path.basename(s) == "assemble.cu":
            o = os.path.join(ROOT, "tools", "_prof", "assemble_prof.o")
            subprocess.check_call(["nvcc"] + [f for f in g.NVCC_FLAGS if f != "-shared"] + ["-DMB200_ASM_PROFILE", "-c", s, "-o", o])
        objs.append(o)
    subprocess.check_call(["nvcc", "-shared", "-o", PROF] + objs)
    print("built", PROF)
    sys.exit(0)

os.environ["MANTA_B200_LIBRARY"] = PROF
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import manta_b200  # noqa: E402
from manta_b200 import _cabi as cb, synth  # noqa: E402

batch = synth.make_candidate_batch(args.loci, seed=args.seed)
opts = cb.default_candidate_opts()
cuts = cb.default_ref_cuts(batch.nLoci)
NAMES = ["pack reads", "windows+init", "key extract", "insert", "succ/pred", "kahn", "cycle replay", "seed select", "walk",
         "pseudo reads", "(loop exit)", "select+output", "replay: order", "replay: DFS", "pred+unused"]
with manta_b200.Context(0) as ctx:
    lib = ctx.lib
    buf = (C.c_ulonglong * 32)()
    ctx.candidates_upload(batch, cuts)
    ctx.candidates_run(opts, select=True)
    lib.mb200_debug_asm_profile(buf, 1)
    ctx.candidates_upload(batch, cuts)
    ctx.candidates_run(opts, select=True)
    t = ctx.candidates_timing()
    lib.mb200_debug_asm_profile(buf, 0)
v = np.array(list(buf), dtype=np.float64)
tot = v[:15].sum()
print("asm_ms", t["asmMs"], "loci", int(v[19]))
for i, n in enumerate(NAMES):
    print("%-14s %6.2f %%   %10.0f cycles/locus" % (n, 100 * v[i] / tot, v[i] / v[19]))
print("k-iterations/locus %.2f  contigs/locus %.2f  walk steps/locus %.0f  words/iter %.0f  windows/iter %.0f  cyclic iters %.3f" %
      (v[16] / v[19], v[17] / v[19], v[18] / v[19], v[20] / v[16], v[21] / v[16], v[22] / v[16]))
