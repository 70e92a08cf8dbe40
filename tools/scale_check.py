#!/usr/bin/env python
"""One-off parity check at scale (too slow for the test suite): the fused GPU path on N config-2 loci against the
multi-threaded CPU composition of the oracle (and of the reference build when present) through whole-batch checksums:
contig count, sum of alignment scores, DP cell count, QC/selection checksum.

    python tools/scale_check.py --loci 6000 --seed 12345 --threads 16
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manta_b200  # noqa: E402
from manta_b200 import _cabi as cb, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--loci", type=int, default=6000)
ap.add_argument("--seed", type=int, default=12345)
ap.add_argument("--threads", type=int, default=16)
args = ap.parse_args()

batch = synth.make_candidate_batch(args.loci, seed=args.seed)
opts = cb.default_candidate_opts()
cuts = cb.default_ref_cuts(batch.nLoci)
A = int(opts["asmOpts"]["maxAssemblyCount"][0])
with manta_b200.Context(0) as ctx:
    ctx.candidates_upload(batch, cuts)
    out = ctx.candidates_run(opts, select=True)
    t = ctx.candidates_timing()
nC = out["loci"]["nContigs"]
live = np.arange(A)[None, :] < nC[:, None]
score = np.where((batch.isJump != 0)[:, None], out["jump"]["score"].reshape(-1, A), out["aln"]["score"].reshape(-1, A))
gpu = (int(live.sum()), int(t["indelCells"] + t["jumpCells"]), int(score[live].sum()),
       cb.qc_checksum(batch.isJump, nC, A, out["trim"], out["qc"], out["best"], True))
print("gpu      :", gpu, flush=True)
ok = True
libs = [("oracle", os.path.join(ROOT, "oracle", "_build", "libmanta_oracle.so"), "orc_"),
        ("reference", os.path.join(ROOT, "oracle", "_ref", "libmanta_ref.so"), "ref_")]
for name, path, prefix in libs:
    if not os.path.exists(path):
        continue
    b = cb.Binding(C.CDLL(path), prefix)
    t0 = time.time()
    got = b.candidates_mt(opts, batch, batch.nLoci, args.threads, cuts, True)
    print("%-9s: %s  (%.1f s)" % (name, got, time.time() - t0), flush=True)
    ok &= (got == gpu)
print("PARITY OK" if ok else "PARITY MISMATCH")
sys.exit(0 if ok else 1)
